"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.  python tools/ncu_launch_table.py file.csv [topN]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    u = row["Metric Unit"]
    ms = v / 1e6 if u.startswith("ns") else (v / 1e3 if u.startswith("us") else v)
    name = re.sub(r"f5b::", "", re.sub(r"\(.*", "", row["Kernel Name"]))
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"| `{k[:90]}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0] * 1e3:.1f} |")
print(f"\ntotal {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised ncu replay; shares, not absolutes, are comparable)")
