"""Generates tests/golden/align_golden.pt from the REFERENCE's own code (run in the build container, where /root/reference exists):
  python tests/golden/make_align_golden.py
alignment_utils.py imports viphoneme / phonemizer at module level (absent here), so the two pure-torch functions are compiled from
the reference file's syntax tree at generation time -- nothing of the reference's source is stored in this repository.
DurationPredictor is imported from its file directly."""
import ast
import importlib.util
import os

import torch

REF = "/root/reference/src/f5_tts/model"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "align_golden.pt")


def ref_functions(names):
    src = open(os.path.join(REF, "alignment_utils.py"), encoding="utf-8").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "alignment_utils.py", "exec"), ns)
    return [ns[n] for n in names]


def main():
    vit, win, prog = ref_functions(["viterbi_vectorized_alignment", "windowed_monotonic_alignment", "progressive_monotonic_alignment"])
    g = torch.Generator().manual_seed(2024)
    cases = []
    for (b, nt, T, kind) in [(2, 5, 23, "randn"), (3, 12, 64, "randn"), (1, 1, 9, "randn"), (2, 7, 7, "randn"), (1, 9, 40, "neg"),
                             (2, 6, 50, "diag"), (1, 8, 3, "randn"), (2, 10, 33, "pos")]:
        sim = torch.randn(b, nt, T, generator=g)
        if kind == "neg":
            sim = -sim.abs() - 0.1          # every gradient negative: boundaries collapse to 0
        elif kind == "pos":
            sim = sim.abs() + 0.05
        elif kind == "diag":                 # similarity concentrated on a monotone band, like a trained aligner's
            n_idx = torch.arange(nt)[:, None].float() / nt
            t_idx = torch.arange(T)[None, :].float() / T
            sim = 3.0 * torch.exp(-((n_idx - t_idx) ** 2) * 60.0)[None].repeat(b, 1, 1) + 0.3 * sim
        c = dict(sim=sim, viterbi=vit(sim.clone()), progressive=prog(sim.clone()))
        if b > 1:  # the batch-total bookkeeping of the progressive sweep only shows with b > 1; a b = 1 twin shows the plain greedy
            c["progressive_item0"] = prog(sim[:1].clone())
        try:
            c["window"] = win(sim.clone())
        except Exception as e:  # the reference raises on an empty window
            c["window"] = None
            c["window_error"] = type(e).__name__
        cases.append(c)
    spec = importlib.util.spec_from_file_location("ref_duration_predictor", os.path.join(REF, "duration_predictor.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(7)
    dp = mod.DurationPredictor(40, 64, 16, 3, 0.5).eval()
    with torch.no_grad():
        for p in dp.parameters():
            p.add_(0.05 * torch.randn_like(p))
    ids = torch.randint(0, 40, (3, 21), generator=g)
    lens = torch.tensor([21, 13, 1])
    mask = (torch.arange(21)[None] < lens[:, None]).int()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, -1))
    with torch.no_grad():
        out = dp(ids, mask)
        pout = dp.phoneme_forward(ids + 1, mask)
    torch.save(dict(cases=cases, dp=dict(state_dict=dp.state_dict(), args=(40, 64, 16, 3, 0.5), ids=ids, mask=mask, out=out,
                                         phoneme_out=pout)), OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [(tuple(c["sim"].shape), c.get("window_error")) for c in cases])


if __name__ == "__main__":
    main()
