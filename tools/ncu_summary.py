"""Summarise an .ncu-rep (raw metrics + per-opcode / per-line stall samples).  python tools/ncu_summary.py file.ncu-rep [topN]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct", "launch__occupancy_limit", "smsp__inst_executed.sum", "sm__throughput.avg.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts",
        "sm__cycles_active.avg", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg"]
for ki, V in enumerate(rows[2:]):
    name = V[H.index("Kernel Name")] if "Kernel Name" in H else ""
    print(f"=== launch {ki}: {name[:100]}")
    for h, u, v in zip(H, rows[1], V):
        if any(h.startswith(w) for w in want):
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
bsel = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for blk in blocks[bsel:bsel + 1]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    H = rows[1]; idx = {h: i for i, h in enumerate(H)}
    data = [r for r in rows[2:] if len(r) == len(H)]
    def f(r, k):
        try: return float(r[idx[k]])
        except Exception: return 0.0
    tot = sum(f(r, "# Samples") for r in data) or 1
    stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    print("--- stall reasons (all samples)")
    for s, v in sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda kv: -kv[1])[:8]:
        print(f"  {s:26s} {100*v/tot:5.1f}%")
    agg = collections.defaultdict(lambda: [0, 0])
    for r in data:
        t = r[idx["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")).split(".")[0]
        agg[op][0] += f(r, "# Samples"); agg[op][1] += f(r, "Instructions Executed")
    print("--- by opcode: samples%, warp-instructions")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"  {k:10s} {100*v[0]/tot:5.1f}%  {v[1]:14.0f}")
    print("  total warp-instructions", sum(v[1] for v in agg.values()))
    print(f"--- top {topn} instructions")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
        st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
        print(f"  {100*f(r,'# Samples')/tot:4.1f}%  {r[idx['Source']].strip()[:80]:80s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
