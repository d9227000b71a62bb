"""Monotonic alignment search at a training-batch shape: GPU kernels vs the CPU restatement of the reference loop.
    python tests/scripts/align_bench.py [B] [nt] [T]
The reference (model/alignment_utils.py:154-212) runs nt x T torch ops per call in a Python double loop; the oracle is the same
recurrence in numpy, vectorised over the batch exactly like the reference, so its time is a LOWER bound on the reference's."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200.model import alignment_utils as U  # noqa: E402
from oracle import align_oracle as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 200
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
g = torch.Generator().manual_seed(0)
n_idx = torch.arange(nt)[:, None].float() / nt
t_idx = torch.arange(T)[None, :].float() / T
sim = 3.0 * torch.exp(-((n_idx - t_idx) ** 2) * 200.0)[None] + 0.3 * torch.randn(B, nt, T, generator=g)
d = sim.cuda()
for name, fn in (("viterbi", U.viterbi_vectorized_alignment), ("window", U.windowed_monotonic_alignment)):
    for _ in range(3):
        fn(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = fn(d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: GPU {ms:.3f} ms per call  ({B} x {nt} x {T}; {B * nt * T * 4 / ms / 1e6:.1f} GB/s of similarity read)")
t0 = time.perf_counter()
ref, _ = A.viterbi_alignment(sim.numpy())
cpu = time.perf_counter() - t0
print(f"viterbi: CPU oracle (numpy, batch-vectorised like the reference loop) {cpu * 1e3:.0f} ms;  bit-exact vs GPU:",
      bool(np.array_equal(ref, U.viterbi_vectorized_alignment(d).cpu().numpy())))
