"""North-star acceptance tests at BASELINE.json's own configurations: the product `CFM.sample` — F5TTS_Base
depth 22 and the pruned F5TTS_Small, NFE 32, sway -1, CFG 2, ragged batches at the benchmark sequence lengths — against the fp32
oracle running on the same GPU (oracle/acceptance.py), with north_star's tolerances taken as ABSOLUTE numbers:
velocity max-abs <= 2e-2 per DiT.forward (bf16 path), final log-mel mean-abs <= 1e-2 over the generated frames.
The measured numbers are written to gpurun_out/parity_r02.json (copied to profiles/ after a GPU run)."""
import json
import os

import pytest
import torch

from oracle import acceptance as A
from oracle import f5_oracle as O

from helpers import build_cfm

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(tag, res):
    path = os.path.join(ROOT, "gpurun_out", "parity_r02.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    for src in (os.path.join(ROOT, "profiles", "r02_parity_acceptance.json"), path):  # start from the committed record
        if os.path.exists(src):
            try:
                data.update(json.load(open(src)))
            except Exception:  # noqa: BLE001
                pass
    data[tag] = res
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(f"[parity] {tag}: velocity max-abs {res['velocity_max_abs']:.3e} (ref max {max(v['ref_abs_max'] for v in res['velocity'].values()):.2f}), "
          f"mel mean-abs (generated) {res['mel_mean_abs_generated']:.3e}, max {res['mel_max_abs_generated']:.3e}", flush=True)


@pytest.mark.parametrize("tag,cfg,ref_frames,totals", [
    # cfg-2's shape: F5TTS_Base (1024 / 22 / 16), 20 s utterances = 1875 frames, ragged, 6 s reference
    ("base_d22_n1875", O.DiTConfig(), 563, [1875, 1610, 1333]),
    # cfg-3's shape: F5TTS_Small (768 / 12 heads) depth-pruned to 12 blocks, 1376-frame chunks, 8 s reference
    ("small_pruned12_n1376", O.DiTConfig(dim=768, heads=12, depth=12), 750, [1376, 1290, 1201, 1100]),
])
def test_sample_nfe32_vs_fp32_oracle_on_gpu(tag, cfg, ref_frames, totals):
    model, sd = build_cfm(cfg, 0)
    res = A.sample_parity(model, sd, cfg, ref_frames, totals, steps=32, cfg_strength=2.0, sway=-1.0, seed=0)
    if tag.startswith("base"):
        res["reference_eager_bf16_velocity_error"] = A.eager_bf16_velocity_error(sd, cfg, ref_frames, totals)
    _record(tag, res)
    assert res["velocity_max_abs"] <= A.VEL_TOL_BF16, res["velocity"]
    assert res["mel_mean_abs_generated"] <= A.MEL_MEAN_TOL, res


def test_sample_b1_cfg1_shape_vs_fp32_oracle_on_gpu():
    """cfg-1's shape (the reference wrapper's serial B = 1 chunk: no key mask, CUDA-graph step path): 940 frames, depth 22, NFE 32"""
    cfg = O.DiTConfig()
    model, sd = build_cfm(cfg, 0)
    res = A.sample_parity(model, sd, cfg, 376, [940], steps=32, cfg_strength=2.0, sway=-1.0, seed=0)
    _record("base_d22_b1_n940", res)
    assert res["velocity_max_abs"] <= A.VEL_TOL_BF16, res["velocity"]
    assert res["mel_mean_abs_generated"] <= A.MEL_MEAN_TOL, res
