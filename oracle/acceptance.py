"""North-star acceptance check (TEST / BENCH-CHECKER INFRASTRUCTURE, never on the product path): the product `CFM` against the
fp32 oracle running on the SAME GPU, at BASELINE.json's own configurations (F5TTS_Base depth 22 / pruned F5TTS_Small, NFE 32,
sway sampling, CFG), with the tolerances north_star states, taken as ABSOLUTE numbers:

  * DiT velocity field, bf16 tensor-core path: max |v - v_ref| <= 2e-2 (per DiT.forward, cond and uncond branch, at the first
    and at a middle point of the reference trajectory);
  * final log-mel after the whole ODE: mean |mel - mel_ref| <= 1e-2 over the GENERATED frames (the reference frames are copied
    from the prompt and would only dilute the mean).

Used by tests/test_gpu_acceptance.py (asserting) and by bench.py's `parity` leg (reporting, after the timed region)."""
from __future__ import annotations

import contextlib

import torch

from . import f5_oracle as O
from .weights import synthetic_inputs

VEL_TOL_BF16 = 2e-2       # north_star: "2e-2 max absolute error in bf16"
VEL_RTOL_FP32 = 1e-3      # north_star: "1e-3 relative error on the DiT velocity field in fp32"
MEL_MEAN_TOL = 1e-2       # north_star: "<= 1e-2 mean-absolute log-mel error on the final mel"


@contextlib.contextmanager
def strict_fp32():
    """the oracle on a GPU must be real fp32: no TF32 in cuBLAS / cuDNN (cuDNN's default would use it for conv_pos_embed)"""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


def _noise(duration, mel_dim, seed):
    y0 = []
    for dur in duration.tolist():
        torch.manual_seed(seed)
        y0.append(torch.randn(int(dur), mel_dim))
    return torch.nn.utils.rnn.pad_sequence(y0, padding_value=0, batch_first=True)


@torch.no_grad()
def sample_parity(model, sd, cfg: O.DiTConfig, ref_frames: int, totals, steps=32, cfg_strength=2.0, sway=-1.0, seed=0,
                  input_seed=1234, device="cuda", mid_step=None, **sample_kw) -> dict:
    """Runs product `model.sample` and `oracle.cfm_sample` (fp32, on `device`) on the same synthetic ragged batch and returns
    the measured errors (plain floats).  `sd` is the oracle's CPU state dict the product model was loaded from."""
    dev = torch.device(device)
    B = len(totals)
    cond, text, duration, lens = synthetic_inputs(cfg, B, ref_frames, list(totals), seed=input_seed)
    noise = _noise(duration, cfg.mel_dim, seed)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    with strict_fp32():
        ref_out, ref_traj = O.cfm_sample(sd_dev, cfg, cond.to(dev), text.to(dev), duration.to(dev), lens=lens.to(dev), steps=steps,
                                         cfg_strength=cfg_strength, sway_sampling_coef=sway, seed=seed)
    out, traj = model.sample(cond=cond.to(dev), text=text.to(dev), duration=duration.to(dev), lens=lens.to(dev), steps=steps,
                             cfg_strength=cfg_strength, sway_sampling_coef=sway, seed=seed, noise=noise, **sample_kw)
    torch.cuda.synchronize()
    n = int(duration.max())
    valid = (torch.arange(n)[None, :] < duration[:, None]).to(dev)                      # frames inside each utterance
    gen = valid & ~(torch.arange(n)[None, :] < lens[:, None]).to(dev)                   # ... that were generated
    res = dict(B=B, n=n, depth=cfg.depth, dim=cfg.dim, heads=cfg.heads, steps=steps, ref_frames=ref_frames, totals=list(totals))
    err = (out.float() - ref_out).abs()
    res["mel_mean_abs_generated"] = float(err[gen].mean())
    res["mel_max_abs_generated"] = float(err[gen].max())
    res["mel_mean_abs_all_valid"] = float(err[valid].mean())
    res["mel_ref_abs_mean"] = float(ref_out[gen].abs().mean())
    # trajectory drift, per step: mean |y_i - y_i_ref| over valid frames
    drift = [(traj[i].float() - ref_traj[i]).abs()[valid].mean() for i in (1, steps // 4, steps // 2, 3 * steps // 4, steps)]
    res["traj_mean_abs_at_1_q1_q2_q3_end"] = [float(d) for d in drift]

    # velocity field: one DiT.forward per branch at identical inputs (the reference trajectory's states), product vs oracle
    t_grid = O.sway_time_grid(steps, sway, device=dev)
    mask = valid if B > 1 else None
    vel = {}
    for name, i in (("first", 0), ("mid", steps // 2 if mid_step is None else mid_step)):
        x = ref_traj[i].contiguous()
        n_cond = torch.nn.functional.pad(cond, (0, 0, 0, n - cond.shape[1])).to(dev)
        cm = (torch.arange(n)[None, :] < lens[:, None]).to(dev).unsqueeze(-1)
        step_cond = torch.where(cm, n_cond, torch.zeros_like(n_cond))
        for branch, (da, dt) in (("cond", (False, False)), ("uncond", (True, True))):
            with strict_fp32():
                ref = O.dit_forward(sd_dev, cfg, x, step_cond, text.to(dev), t_grid[i], da, dt, mask)
            got = model.transformer(x=x, cond=step_cond, text=text.to(dev), time=t_grid[i], drop_audio_cond=da, drop_text=dt, mask=mask)
            torch.cuda.synchronize()
            d = (got.float() - ref).abs()
            vel[f"{name}_{branch}"] = dict(max_abs=float(d[valid].max()), mean_abs=float(d[valid].mean()),
                                           max_abs_all_rows=float(d.max()), ref_abs_max=float(ref[valid].abs().max()),
                                           ref_rms=float(ref[valid].pow(2).mean().sqrt()),
                                           rel_fro=float((got.float() - ref)[valid].norm() / ref[valid].norm()))
    res["velocity"] = vel
    res["velocity_max_abs"] = max(v["max_abs"] for v in vel.values())
    res["velocity_rel_fro"] = max(v["rel_fro"] for v in vel.values())
    return res


@torch.no_grad()
def eager_bf16_velocity_error(sd, cfg: O.DiTConfig, ref_frames: int, totals, input_seed=1234, seed=0, device="cuda") -> dict:
    """Context for the velocity numbers: the error of the REFERENCE'S OWN bf16 GPU path (the oracle network cast to bf16 and run
    as stock PyTorch eager — SURVEY 8c adjustment 4) against the same fp32 oracle, at the first trajectory point."""
    dev = torch.device(device)
    B = len(totals)
    cond, text, duration, lens = synthetic_inputs(cfg, B, ref_frames, list(totals), seed=input_seed)
    n = int(duration.max())
    x = _noise(duration, cfg.mel_dim, seed).to(dev)
    valid = (torch.arange(n)[None, :] < duration[:, None]).to(dev)
    mask = valid if B > 1 else None
    n_cond = torch.nn.functional.pad(cond, (0, 0, 0, n - cond.shape[1])).to(dev)
    cm = (torch.arange(n)[None, :] < lens[:, None]).to(dev).unsqueeze(-1)
    step_cond = torch.where(cm, n_cond, torch.zeros_like(n_cond))
    t0 = torch.zeros((), device=dev)
    sd32 = {k: v.to(dev) for k, v in sd.items()}
    sd16 = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in sd32.items()}
    out = {}
    for branch, (da, dt) in (("cond", (False, False)), ("uncond", (True, True))):
        with strict_fp32():
            ref = O.dit_forward(sd32, cfg, x, step_cond, text.to(dev), t0, da, dt, mask)
        got = O.dit_forward(sd16, cfg, x.to(torch.bfloat16), step_cond.to(torch.bfloat16), text.to(dev), t0.to(torch.bfloat16), da, dt, mask)
        d = (got.float() - ref).abs()
        out[branch] = dict(max_abs=float(d[valid].max()), mean_abs=float(d[valid].mean()))
    return out
