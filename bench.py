#!/usr/bin/env python
"""bench.py — headline benchmark of the F5-TTS hot path (BASELINE.json: mel-frames/s and RTF, F5TTS_Base, NFE=32, CFG).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg3d18|cfg4|cfg5|cfg5b] [--ragged]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one full pass of the hot path over one batch of synthetic utterances: CFM.sample (text embedding, 32 Euler steps x
{cond, uncond} DiT forwards fused as one 2B batch, CFG + Euler updates) + Vocos decode of the generated frames.
  value : total mel frames (ref + gen, what the DiT processes) per second, inputs already resident in HBM (CUDA events).
  e2e   : the same metric through the public API with HOST buffers (pinned reference waveform + token ids -> device, MelSpec,
          sample, vocoder, audio back to the host), host<->device copies inside the timed region.
  roofline     : dominant kernel class, algorithmic FLOPs / CUDA-event time measured live over the timed steps.
  cpu_baseline : the reference's own modules (byte-compiled into oracle/_ref/ by oracle/make_ref.py; the oracle port if that directory is
                 absent) timed on this box's host cores on a bounded sample.
--impl reference times the same CPU implementation of the path with all host threads on the same workload / metric (cfg-1 in full).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (arch kwargs, batch, ref frames, total frames, description)  — SURVEY.md §8d
    "cfg2": (dict(dim=1024, depth=22, heads=16), 16, 563, 1875, "F5TTS_Base bf16 batch=16 x 20 s utterances (ref 6 s + gen 14 s), NFE=32 sway -1 CFG 2"),
    "cfg1": (dict(dim=1024, depth=22, heads=16), 1, 376, 940, "F5TTS_Base one ~10 s utterance (ref 376 + gen 564 frames), NFE=32 Euler CFG 2"),
    "cfg3": (dict(dim=768, depth=12, heads=12), 32, 750, 1376, "F5TTS_Small pruned to 12 blocks, batch=32 streaming chunks (ref 8 s + 626 gen frames)"),
    # a fixed job of 256 cfg-2 utterances = 16 batches of 16, rank-strided over the GPUs (strong scaling)
    "cfg4": (dict(dim=1024, depth=22, heads=16), 16, 563, 1875, "256 utterances of cfg-2's shape = 16 batches of 16 per step, rank-strided over the GPUs "
             "(strong scaling, no collective on the sampling path)"),
    "cfg3d18": (dict(dim=768, depth=18, heads=12), 32, 750, 1376, "F5TTS_Small un-pruned (18 blocks), batch=32 streaming chunks (ref 8 s + 626 gen frames)"),
    # training step: ref_frames unused
    "cfg5": (dict(dim=1024, depth=22, heads=16), 32, 0, 1200, "F5TTS_Base one optimizer step on 32 x 1200 frames per GPU: CFM.forward + backward + "
             "gradient all-reduce + clip + AdamW + EMA, bf16 operands / fp32 master, dropout 0"),
    "cfg5b": (dict(dim=1024, depth=22, heads=16), 64, 0, 600, "F5TTS_Base one optimizer step on 64 x 600 frames per GPU: CFM.forward + backward + "
              "gradient all-reduce + clip + AdamW + EMA, bf16 operands / fp32 master, dropout 0"),
}
NFE, CFG, SWAY = 32, 2.0, -1.0


def dit_flops_per_forward(cfg, B, n):
    """algorithmic FLOPs of one DiT.forward over B x n tokens: 2MNK per GEMM, 4 n D per token attention."""
    D, F = cfg.dim, cfg.ff_mult * cfg.dim
    per_tok_block = 2 * D * 3 * D + 2 * D * D + 2 * 2 * D * F + 4 * n * D
    cpg = D // 16
    outside = 2 * (2 * cfg.mel_dim + cfg.text_dim) * D + 2 * 2 * D * cpg * 31 + 2 * D * cfg.mel_dim
    return B * n * (cfg.depth * per_tok_block + outside)


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons during the timed region (NVML; nvidia-smi CLI as a fallback)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit and nm != "gpu_idle":
                            self.reasons.add(nm)
                else:
                    import subprocess
                    o = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.sw_power_cap,"
                                        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(int(o[0]))
                    self.max_mhz = int(o[1])
                    for nm, v in zip(("sw_power_cap", "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"), o[2:]):
                        if "Active" in v and "Not" not in v:
                            self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


class _StdoutToStderr:
    """NCCL prints its version banner on stdout when the first communicator is created; the contract is ONE JSON line on stdout.
    Redirect fd 1 to fd 2 at the OS level around process-group creation + the first collective."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def init_nccl(dev):
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    with _StdoutToStderr():
        dist.init_process_group("nccl", device_id=dev)
        t = torch.zeros(1, device=dev)
        dist.all_reduce(t)  # forces communicator creation (and the banner) now
        torch.cuda.synchronize()


class Arch:
    """model.arch constants of the workload (configs/F5TTS_Base.yaml:25-34 / F5TTS_Small.yaml)"""
    ff_mult, mel_dim, text_dim, conv_layers, text_num_embeds, pe_attn_head, text_mask_padding = 2, 100, 512, 4, 2545, 1, False

    def __init__(self, dim, depth, heads):
        self.dim, self.depth, self.heads = dim, depth, heads


def make_inputs(arch, B, ref_frames, total, seed):
    """SURVEY.md §8d synthetic workload: log-mel-like reference mel / 0.1*randn reference audio, random token ids (-1 padded)."""
    g = torch.Generator().manual_seed(seed)
    cond = (torch.randn(B, ref_frames, arch.mel_dim, generator=g) * 2.0 - 1.5).clamp(-11.5, 5.0)
    nt = max(2, int(0.16 * total))
    text = torch.randint(0, arch.text_num_embeds, (B, nt), generator=g)
    for b in range(B):
        text[b, max(1, int(nt * (0.6 + 0.4 * (b + 1) / B))):] = -1
    duration = torch.full((B,), total, dtype=torch.long)
    lens = torch.full((B,), ref_frames, dtype=torch.long)
    wav = 0.1 * torch.randn(B, 256 * (ref_frames - 1), generator=g)  # MelSpec gives 1 + L // 256 = ref_frames frames
    return cond, text, duration, lens, wav


def build_product_models(arch, dev):
    """random-init F5TTS DiT + Vocos exactly as the reference constructs them (torch default init), with the zero-initialised
    AdaLN / proj_out tensors re-drawn N(0, 0.02) so the network output is not identically zero (checkpoints are not available)."""
    from eraxvif5tts_b200.model import CFM, DiT
    from eraxvif5tts_b200.vocoder import Vocos
    torch.manual_seed(0)
    tr = DiT(dim=arch.dim, depth=arch.depth, heads=arch.heads, ff_mult=arch.ff_mult, mel_dim=arch.mel_dim, text_num_embeds=arch.text_num_embeds,
             text_dim=arch.text_dim, text_mask_padding=arch.text_mask_padding, conv_layers=arch.conv_layers, pe_attn_head=arch.pe_attn_head)
    model = CFM(transformer=tr, mel_spec_kwargs=dict(n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=arch.mel_dim,
                                                     target_sample_rate=24000, mel_spec_type="vocos"), odeint_kwargs=dict(method="euler"))
    with torch.no_grad():
        for p_ in model.parameters():
            if float(p_.abs().max()) == 0.0:
                p_.normal_(0, 0.02)
    voc = Vocos()
    with torch.no_grad():
        voc.head.out.weight.mul_(0.5)  # keep exp(mag) in a log-mel-like range
    return model.to(dev).eval(), voc.to(dev).eval()


def cpu_reference_measure(arch, ref_frames, total, steps, warmup, full=False):
    """The CPU arm.  When the reference's own modules are loadable (oracle/ref_shim: /root/reference in the build container, the
    byte-compiled oracle/_ref/ on the GPU box) this times the REFERENCE ITSELF: its `CFM` / `DiT` classes, fp32, all host threads
    (`kind: "reference"`; only the un-vendored third-party pieces — torchdiffeq's fixed-grid solver, x_transformers' rotary
    embedding, the vocos vocoder — are the restatements of oracle/).  Otherwise the oracle port (`kind: "port"`).
      full=False: bounded sample = ONE utterance of the workload x ONE Euler step (cond + uncond DiT.forward) per timed step, + one
                  text-embedding pair + one Vocos decode; frames/s extrapolated to the NFE identical steps;
      full=True : the complete CFM.sample (NFE steps, CFG, sway) + Vocos decode of one utterance per timed step (cfg-1, as BASELINE.md
                  section 5.2 promises); at most one timed sample so the run stays within minutes."""
    from oracle import f5_oracle as O
    from oracle import ref_shim
    from oracle.weights import make_dit_state_dict, make_vocos_state_dict, synthetic_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.DiTConfig(dim=arch.dim, depth=arch.depth, heads=arch.heads)
    sd = make_dit_state_dict(cfg, 0)
    vc = O.VocosConfig()
    vsd = make_vocos_state_dict(vc, 1)
    cond, text, duration, lens = synthetic_inputs(cfg, 1, ref_frames, total, seed=1234)
    n = total
    kind = "reference" if ref_shim.available() else "port"
    ref_model = ref_shim.build_reference_cfm(cfg, sd) if kind == "reference" else None
    how = ("the reference's own CFM / DiT modules (fp32 torch CPU), loaded from " +
           ("/root/reference" if ref_shim.source_available() else "the byte-compiled oracle/_ref/")) if kind == "reference" else "oracle port (fp32 torch CPU)"
    with torch.no_grad():
        if full:
            def one_sample(nsteps):
                if ref_model is not None:
                    out, _ = ref_model.sample(cond=cond, text=text, duration=duration, lens=lens, steps=nsteps, cfg_strength=CFG,
                                              sway_sampling_coef=SWAY, seed=0)
                else:
                    out, _ = O.cfm_sample(sd, cfg, cond, text, duration, lens=lens, steps=nsteps, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=0)
                return O.vocos_decode(vsd, vc, out[:, ref_frames:].permute(0, 2, 1))
            if warmup > 0:
                one_sample(1)  # page-in / thread-pool warm-up: a 1-step sample
            times = []
            for _ in range(max(1, min(steps, 1))):
                t0 = time.perf_counter()
                one_sample(NFE)
                times.append(time.perf_counter() - t0)
            per_utt = sum(times) / len(times)
            return dict(value=total / per_utt, pair_s=per_utt / NFE, per_utterance_s=per_utt, cores=torch.get_num_threads(), kind=kind,
                        sample=f"{how}: the COMPLETE CFM.sample of 1 utterance ({total} frames, NFE {NFE}, CFG, sway) + Vocos decode, "
                               f"{len(times)} timed run(s) after a 1-step warm-up sample")
        step_cond = torch.nn.functional.pad(cond, (0, 0, 0, n - ref_frames))
        y = torch.randn(1, n, cfg.mel_dim)
        t0 = time.perf_counter()
        if ref_model is None:
            te_c = O.text_embedding(sd, cfg, text, n, False)
            te_u = O.text_embedding(sd, cfg, text, n, True)
        else:  # DiT caches both on the first cond / uncond forward (dit.py:202-210); time the two TextEmbedding calls themselves
            tr = ref_model.transformer
            tr.clear_cache()
            tr.text_cond = tr.text_embed(text, n, drop_text=False)
            tr.text_uncond = tr.text_embed(text, n, drop_text=True)
        t_embed = time.perf_counter() - t0
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            t = torch.tensor(0.3)
            if ref_model is None:
                pred = O.dit_forward(sd, cfg, y, step_cond, text, t, False, False, None, text_embed=te_c)
                null = O.dit_forward(sd, cfg, y, step_cond, text, t, True, True, None, text_embed=te_u)
            else:
                pred = tr(x=y, cond=step_cond, text=text, time=t, drop_audio_cond=False, drop_text=False, cache=True)
                null = tr(x=y, cond=step_cond, text=text, time=t, drop_audio_cond=True, drop_text=True, cache=True)
            y = y + 0.03 * (pred + (pred - null) * CFG)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        t0 = time.perf_counter()
        O.vocos_decode(vsd, vc, y[:, ref_frames:].permute(0, 2, 1))
        t_voc = time.perf_counter() - t0
    pair = sum(times) / len(times)
    per_utt = NFE * pair + t_embed + t_voc
    return dict(value=total / per_utt, pair_s=pair, embed_s=t_embed, vocos_s=t_voc, per_utterance_s=per_utt,
                cores=torch.get_num_threads(), kind=kind,
                sample=f"{how}: 1 utterance ({total} frames) x 1 of {NFE} Euler steps (cond+uncond DiT.forward) per timed step, mean of "
                       f"{len(times)}; +1 text-embedding pair +1 Vocos decode; extrapolated x{NFE} steps")


def torch_eager_gpu_measure(arch, B, ref_frames, total, dev, steps=3, warmup=2):
    """Second comparator of SURVEY.md §8d (opt-in, --torch-eager-gpu): the oracle's DiT forward run as stock PyTorch eager in bf16
    on the same B200 — cuBLAS linears, F.scaled_dot_product_attention (flash / cuDNN; no key mask, the favourable case for the
    library path), cuDNN grouped conv.  Bounded sample = the full batch x ONE Euler step (cond + uncond forward) per timed step,
    extrapolated to the 32 identical steps.  It is a baseline leg: nothing here is on the product path."""
    import torch.nn.functional as F
    from oracle import f5_oracle as O
    from oracle.weights import make_dit_state_dict, synthetic_inputs
    cfg = O.DiTConfig(dim=arch.dim, depth=arch.depth, heads=arch.heads)
    sd_cpu = make_dit_state_dict(cfg, 0)
    cond, text, duration, lens = synthetic_inputs(cfg, B, ref_frames, total, seed=1234)
    with torch.no_grad():
        te_c = O.text_embedding(sd_cpu, cfg, text[:1], total, False).expand(B, -1, -1).to(dev, torch.bfloat16)
        te_u = O.text_embedding(sd_cpu, cfg, text[:1], total, True).expand(B, -1, -1).to(dev, torch.bfloat16)
    sd = {k: v.to(dev, torch.bfloat16) for k, v in sd_cpu.items()}
    rope_cpu, attn_cpu = O.rotary_freqs, O.attention

    def attention_sdpa(sd_, cfg_, p, x, mask, rope, drop=None, attn_drop=None):
        b, n, _ = x.shape
        H, d = cfg_.heads, cfg_.dim_head
        q = F.linear(x, sd_[p + "to_q.weight"], sd_[p + "to_q.bias"]).view(b, n, H, d).transpose(1, 2)
        k = F.linear(x, sd_[p + "to_k.weight"], sd_[p + "to_k.bias"]).view(b, n, H, d).transpose(1, 2)
        v = F.linear(x, sd_[p + "to_v.weight"], sd_[p + "to_v.bias"]).view(b, n, H, d).transpose(1, 2)
        pn = cfg_.pe_attn_head if cfg_.pe_attn_head is not None else H
        q = torch.cat((O.apply_rotary(q[:, :pn], rope), q[:, pn:]), dim=1)
        k = torch.cat((O.apply_rotary(k[:, :pn], rope), k[:, pn:]), dim=1)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, H * d)
        return F.linear(o, sd_[p + "to_out.0.weight"], sd_[p + "to_out.0.bias"])

    O.rotary_freqs = lambda n, d=64, theta=10000.0: rope_cpu(n, d, theta).to(dev)
    O.attention = attention_sdpa
    try:
        step_cond = F.pad(cond, (0, 0, 0, total - ref_frames)).to(dev, torch.bfloat16)
        y = torch.randn(B, total, cfg.mel_dim, device=dev, dtype=torch.bfloat16)
        t = torch.tensor(0.3, device=dev, dtype=torch.bfloat16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            for i in range(warmup + steps):
                if i == warmup:
                    torch.cuda.synchronize()
                    e0.record()
                pred = O.dit_forward(sd, cfg, y, step_cond, text, t, False, False, None, text_embed=te_c)
                null = O.dit_forward(sd, cfg, y, step_cond, text, t, True, True, None, text_embed=te_u)
                y = y + 0.03 * (pred + (pred - null) * CFG)
            e1.record()
            torch.cuda.synchronize()
    finally:
        O.rotary_freqs, O.attention = rope_cpu, attn_cpu
    pair_ms = e0.elapsed_time(e1) / steps
    return {"value": B * total / (NFE * pair_ms * 1e-3), "unit": "mel-frames/s", "ms_per_euler_step": pair_ms, "dtype": "bf16",
            "sample": f"full batch ({B} x {total} frames) x 1 of {NFE} Euler steps (cond+uncond DiT forward) per timed step, mean of {steps}; "
                      f"x{NFE} extrapolated; MelSpec / text embedding / Vocos excluded (they favour this arm)",
            "kernels": "torch eager: cuBLAS linears, SDPA without key mask, cuDNN grouped conv"}

def torch_eager_gpu_train_measure(arch, B, n, dev, steps=3, warmup=2):
    """Second comparator for the training step (opt-in, --torch-eager-gpu): the oracle's CFM.forward under torch.autocast(bf16) with
    torch autograd, F.scaled_dot_product_attention, clip_grad_norm_ and fused torch.optim.AdamW on the same B200 -- what the
    reference's Trainer runs per step (accelerate bf16 mixed precision), minus DDP / EMA / text-embedding backward (they favour
    this arm).  A baseline leg: nothing here is on the product path."""
    import torch.nn.functional as F
    from oracle import f5_oracle as O
    from oracle.weights import make_dit_state_dict
    cfg = O.DiTConfig(dim=arch.dim, depth=arch.depth, heads=arch.heads)
    sd_cpu = make_dit_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(1234)
    x1 = (torch.randn(B, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5).to(dev)
    text = torch.randint(0, cfg.text_num_embeds, (B, int(0.16 * n)), generator=g)
    with torch.no_grad():
        te = O.text_embedding(sd_cpu, cfg, text[:1], n, False).expand(B, -1, -1).to(dev)
    sd = {k: (v.to(dev).requires_grad_(True) if v.is_floating_point() else v.to(dev)) for k, v in sd_cpu.items()}
    params = [v for k, v in sd.items() if v.requires_grad and k.startswith("transformer.") and not k.startswith("transformer.text_embed.")]
    opt = torch.optim.AdamW(params, lr=7.5e-5, betas=(0.9, 0.98), weight_decay=0.01, fused=True)
    rope_cpu, attn_cpu = O.rotary_freqs, O.attention

    def attention_sdpa(sd_, cfg_, p, x, mask, rope, drop=None, attn_drop=None):
        b, m, _ = x.shape
        H, d = cfg_.heads, cfg_.dim_head
        q = F.linear(x, sd_[p + "to_q.weight"], sd_[p + "to_q.bias"]).view(b, m, H, d).transpose(1, 2)
        k = F.linear(x, sd_[p + "to_k.weight"], sd_[p + "to_k.bias"]).view(b, m, H, d).transpose(1, 2)
        v = F.linear(x, sd_[p + "to_v.weight"], sd_[p + "to_v.bias"]).view(b, m, H, d).transpose(1, 2)
        pn = cfg_.pe_attn_head if cfg_.pe_attn_head is not None else H
        q = torch.cat((O.apply_rotary(q[:, :pn], rope), q[:, pn:]), dim=1)
        k = torch.cat((O.apply_rotary(k[:, :pn], rope), k[:, pn:]), dim=1)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, m, H * d)
        return F.linear(o, sd_[p + "to_out.0.weight"], sd_[p + "to_out.0.bias"])

    O.rotary_freqs = lambda m, d=64, theta=10000.0: rope_cpu(m, d, theta).to(dev)
    O.attention = attention_sdpa
    try:
        span = torch.zeros(B, n, dtype=torch.bool, device=dev)
        span[:, n // 8: n // 8 + int(0.8 * n)] = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                e0.record()
            x0 = torch.randn_like(x1)
            time_ = torch.rand(B, device=dev)
            t = time_[:, None, None]
            phi, flow = (1 - t) * x0 + t * x1, x1 - x0
            cond = torch.where(span[..., None], torch.zeros_like(x1), x1)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                pred = O.dit_forward(sd, cfg, phi, cond, text, time_, False, False, None, text_embed=te)
            loss = F.mse_loss(pred.float(), flow, reduction="none")[span].mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        e1.record()
        torch.cuda.synchronize()
    finally:
        O.rotary_freqs, O.attention = rope_cpu, attn_cpu
    ms = e0.elapsed_time(e1) / steps
    return {"value": B * n / (ms * 1e-3), "unit": "mel-frames/s", "ms_per_step": ms, "dtype": "bf16 autocast, fp32 master",
            "peak_alloc_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "loss": float(loss.detach()),
            "sample": f"full batch ({B} x {n} frames), forward + backward + clip + fused AdamW per timed step, mean of {steps}",
            "kernels": "torch eager autograd: cuBLAS linears, SDPA (flash) forward / backward, cuDNN grouped conv, fused AdamW"}

def cpu_train_measure(arch, n, steps, warmup):
    """CPU arm of the training step: torch autograd over the oracle's fp32 CFM.forward for ONE utterance of the workload
    (forward + backward; the optimizer is negligible next to them), all host threads."""
    from oracle import f5_oracle as O
    from oracle.weights import make_dit_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.DiTConfig(dim=arch.dim, depth=arch.depth, heads=arch.heads)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in make_dit_state_dict(cfg, 0).items()}
    g = torch.Generator().manual_seed(1234)
    x1 = (torch.randn(1, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5)
    x0 = torch.randn(1, n, cfg.mel_dim, generator=g)
    text = torch.randint(0, cfg.text_num_embeds, (1, int(0.16 * n)), generator=g)
    span = torch.zeros(1, n, dtype=torch.bool)
    span[0, n // 4: n // 4 + int(0.85 * n) // 1] = True
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = O.cfm_loss(sd, cfg, x1, text, span, x0, torch.tensor([0.4]), False, False)
        loss.backward()
        for v in sd.values():
            if v.is_floating_point():
                v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    return dict(value=n / per, step_s=per, cores=torch.get_num_threads(),
                sample=f"1 utterance ({n} frames) forward + backward (torch autograd over the fp32 oracle), mean of {len(times)}")


def run_train(args, cfg, B, n, desc, rank, local_rank, world):
    """--workload cfg5 / cfg5b: one data-parallel optimizer step per timed step"""
    import torch.distributed as dist
    if args.impl == "reference":
        if rank != 0:
            return
        config = {"workload": f"{args.workload}: {desc}", "batch_per_gpu": B, "frames_per_utterance": n}
        r = cpu_train_measure(cfg, n, max(1, min(args.steps, 2)), 1)
        print(json.dumps({"impl": "reference", "metric": "train_mel_frames_per_sec", "value": r["value"], "unit": "mel-frames/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["step_s"] * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": r["value"], "unit": "mel-frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                          "e2e": {"value": r["value"], "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    line = measure_train(args, args.workload, cfg, B, n, desc, rank, local_rank, world, dev, args.steps, args.warmup)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def measure_train(args, wname, cfg, B, n, desc, rank, local_rank, world, dev, steps, warmup, cpu_baseline=True, eager=None):
    """one data-parallel optimizer step per timed step (cfg-5): CFM.forward + backward + NCCL gradient all-reduce + clip + AdamW +
    EMA.  The process group (world > 1) is the caller's.  Returns the JSON line on rank 0, None elsewhere."""
    import torch.distributed as dist
    config = {"workload": f"{wname}: {desc}", "batch_per_gpu": B, "frames_per_utterance": n,
              "parallelism": f"dp{world} (batch sharded; ONE flat fp32 gradient all-reduce per step over NCCL)",
              "l2": "no flush: one step streams ~30 GB of saved activations, >> 126 MB L2",
              "profiling": "timed region un-instrumented; kernel classes / roofline from 2 extra event-bracketed steps"}
    from eraxvif5tts_b200 import _lib as L
    from eraxvif5tts_b200.train import TrainEngine
    L.load()
    model, _ = build_product_models(cfg, dev)
    train_dropout = float(os.environ.get("F5B_TRAIN_DROPOUT", "0"))  # cfg-5 is quoted at dropout 0 (parity setting); 0.1 = the reference's training default
    ckpt = os.environ.get("F5B_TRAIN_CHECKPOINT", "0") == "1"  # the reference's checkpoint_activations option (dit.py:221-223)
    attn_dropout = float(os.environ.get("F5B_TRAIN_ATTN_DROPOUT", "0"))  # SDPA's own dropout (modules.py:490 hard-codes 0.1); 0 = parity setting
    eng = TrainEngine(model, with_ema=(rank == 0), dropout=train_dropout, checkpoint_activations=ckpt,
                      attn_dropout=attn_dropout)  # EMA only on the main process (trainer.py:179-181)
    config["dropout"] = train_dropout
    config["attn_dropout"] = attn_dropout
    config["checkpoint_activations"] = ckpt
    if train_dropout > 0:
        config["workload"] = config["workload"].replace("dropout 0", f"dropout {train_dropout:g} (FeedForward + to_out sites)")
    if world > 1:
        eng.broadcast_params(0)
    g = torch.Generator().manual_seed(1234 + rank)
    mel_h = (torch.randn(B, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5).pin_memory()
    text_h = torch.randint(0, cfg.text_num_embeds, (B, int(0.16 * n)), generator=g).pin_memory()
    mel_d, text_d = mel_h.to(dev), text_h.to(dev)
    loss_h = torch.zeros(1).pin_memory()
    overlap = world > 1 and os.environ.get("F5B_ALLREDUCE_OVERLAP", "1") != "0"
    config["allreduce"] = ("bucketed (4 groups of blocks), launched under the remaining backward" if overlap else
                           "one flat all-reduce after the backward") if world > 1 else "none (1 GPU)"

    def step(mel, text):
        eng.zero_grad()
        loss, _, _ = eng.loss_and_grads(mel, text, overlap_allreduce=overlap)
        scale = eng.allreduce_grads()
        eng.step(grad_scale=scale)
        return loss

    def step_e2e():
        loss = step(mel_h.to(dev, non_blocking=True), text_h.to(dev, non_blocking=True))
        loss_h.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 1)):
        loss = step(mel_d, text_d)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the timed region runs WITHOUT per-launch event bracketing (about 780 event pairs per step cost ~3 % here); the kernel-class
    # breakdown and the roofline come from extra bracketed steps right after it
    L.prof_reset(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        loss = step(mel_d, text_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_per_step = 0
    prof_steps = 2
    if not args.no_profile:
        L.prof_reset(True)
        for _ in range(prof_steps):
            loss = step(mel_d, text_d)
        barrier()
    prof = L.prof_read()
    L.prof_reset(False)
    assert torch.isfinite(loss).all(), "non-finite loss"
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        eng.release()
        return None
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kinds = {k: {"launches_per_step": v["launches"] / prof_steps, "ms_per_step": v["ms"] / prof_steps, "share": v["ms"] / tot_ms,
                 **({"tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12} if v["flops"] > 0 and v["ms"] > 0 else {}),
                 **({"gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9} if v["ms"] > 0 else {})} for k, v in prof.items() if v["launches"]}
    roofline = None
    timed = {k: v for k, v in prof.items() if v["ms"] > 0 and v["flops"] > 0}
    if timed:
        top = max(timed, key=lambda k: timed[k]["ms"])
        v = timed[top]
        ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
        roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                    "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback", "launches": v["launches"],
                    "avg_launch_ms": v["ms"] / v["launches"], "flops_per_launch": v["flops"] / v["launches"]}
    frames = B * n * steps * world
    step_flops = 3 * dit_flops_per_forward(cfg, B, n)
    line = {"metric": "train_mel_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "mel-frames/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config, "clocks": sampler.summary(),
            "e2e": {"value": frames / e2e_s, "unit": "mel-frames/s", "h2d_bytes_per_step": mel_h.numel() * 4 + text_h.numel() * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3 / steps},
            "gpu_launches": int(sum(v["launches"] for v in prof.values()) / prof_steps * steps),
            "model_tflops_per_gpu": step_flops / (ms * 1e-3 / steps) / 1e12, "params": eng.n,
            "loss": float(loss), "roofline": roofline, "kernels": kinds}
    if world == 1 and cpu_baseline and not args.no_cpu_baseline:
        r = cpu_train_measure(cfg, n, 1, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "mel-frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    eng.release()
    del eng, model
    torch.cuda.empty_cache()
    if world == 1 and (args.torch_eager_gpu if eager is None else eager):
        torch.cuda.reset_peak_memory_stats(dev)
        line["torch_eager_gpu"] = torch_eager_gpu_train_measure(cfg, B, n, dev)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 (default, BASELINE.json's metric config), cfg1, cfg3, cfg3d18, cfg4 (fixed 256-utterance job, strong scaling): inference; cfg5, cfg5b: the training step")
    ap.add_argument("--batch", type=int, default=0, help="experiments only: override the workload's utterances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ragged", action="store_true", help="inference workloads: per-utterance durations U[0.8, 1] x the nominal length "
                    "(SURVEY.md 8d's ragged cfg-2 variant, U[1500, 1875]); the value counts un-padded frames")
    ap.add_argument("--no-profile", action="store_true", help="skip the extra event-bracketed step(s) behind the timed region (no kernel breakdown / roofline)")
    ap.add_argument("--torch-eager-gpu", action="store_true", help="also time the oracle as stock PyTorch eager bf16 on this GPU (second comparator); "
                    "on by default for the default workload at 1 GPU")
    ap.add_argument("--no-eager", action="store_true", help="default workload: skip the PyTorch-eager GPU comparator legs")
    ap.add_argument("--no-train", action="store_true", help="default workload: skip the cfg-5 training-step sub-record")
    ap.add_argument("--no-parity", action="store_true", help="default workload: skip the depth-22 / NFE-32 parity leg against the fp32 oracle")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    arch_kw, B, ref_frames, total, desc = WORKLOADS[args.workload]
    if args.batch > 0:  # experiment knob (sub-batch / L2-residency A/B); the named configurations fix the batch
        B = args.batch
        desc += f" [--batch {B} override]"
    cfg = Arch(**arch_kw)
    if args.workload.startswith("cfg5"):
        return run_train(args, cfg, B, total, desc, rank, local_rank, world)
    reps = 1
    if args.workload == "cfg4":
        if 16 % world:
            raise SystemExit("cfg4 splits 16 batches over the ranks: --gpus must divide 16")
        reps = 16 // world  # batches this rank samples per step
    frames_per_step = reps * B * total
    gen_frames_per_step = reps * B * (total - ref_frames)
    config = {"workload": f"{args.workload}: {desc}", "batch_per_gpu": B, "frames_per_utterance": total, "ref_frames": ref_frames,
              "nfe": NFE, "cfg_strength": CFG, "sway_sampling_coef": SWAY, "ode": "euler", "vocoder": "vocos-mel-24khz (random init)",
              "parallelism": f"dp{world} (utterance batch sharded, no collective on the sampling path)",
              "l2": "no flush: one DiT forward streams ~1.4 GB of activations + 0.65 GB of weights, >> 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_measure(cfg, ref_frames, total, args.steps, args.warmup, full=(args.workload == "cfg1"))
        line = {"impl": "reference", "metric": "mel_frames_per_sec", "value": r["value"], "unit": "mel-frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["pair_s"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "mel-frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "rtf": r["per_utterance_s"] / ((total - ref_frames) * 256 / 24000.0), "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------------------ ours
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)  # keeps NCCL's version banner off stdout (one JSON line there)
    from eraxvif5tts_b200 import _lib as L
    L.load()
    model, voc = build_product_models(cfg, dev)
    cond, text, duration, lens, wav = make_inputs(cfg, B, ref_frames, total, 1234 + rank)
    if args.ragged:
        duration = torch.randint(int(0.8 * total), total + 1, (B,), generator=torch.Generator().manual_seed(4321))
        duration[0] = total  # the padded length stays the nominal one
        frames_per_step = int(duration.sum())
        gen_frames_per_step = int((duration - ref_frames).sum())
        config["ragged"] = (f"durations U[{int(0.8 * total)}, {total}] (same draw on every rank), padded to {total}; value and e2e count "
                            f"the {frames_per_step} un-padded frames per step; roofline FLOPs count what the kernels execute")
    cond_d, text_d, dur_d, lens_d = cond.to(dev), text.to(dev), duration.to(dev), lens.to(dev)
    wav_h, text_h = wav.pin_memory(), text.pin_memory()
    gen_len = 256 * (total - ref_frames - 1)
    audio_h = torch.empty(B, gen_len, dtype=torch.float32).pin_memory()

    def step_device():
        for _ in range(reps):
            out, _ = model.sample(cond=cond_d, text=text_d, duration=dur_d, lens=lens_d, steps=NFE, cfg_strength=CFG,
                                  sway_sampling_coef=SWAY, seed=0, return_trajectory=False)
            a = voc.decode(out[:, ref_frames:].permute(0, 2, 1))
        return a

    def step_e2e():
        for _ in range(reps):
            w = wav_h.to(dev, non_blocking=True)
            t = text_h.to(dev, non_blocking=True)
            out, _ = model.sample(cond=w, text=t, duration=dur_d, steps=NFE, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=0,
                                  return_trajectory=False)
            a = voc.decode(out[:, ref_frames:].permute(0, 2, 1))
            audio_h.copy_(a, non_blocking=True)
            torch.cuda.synchronize()
        return audio_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # The timed region is UN-INSTRUMENTED: no per-launch events, the product's own launch path (small fused batches replay one
    # captured CUDA graph per ODE step).  The kernel-class breakdown / roofline come from extra event-bracketed steps right behind
    # it (eager launches: events cannot be recorded inside a graph replay), with the same inputs on the same warm device.
    from eraxvif5tts_b200.model import cfm as cfm_mod
    graph_mode = os.environ.get("F5B_CUDA_GRAPH", "") != "0" and (2 * B * total <= cfm_mod.GRAPH_MAX_ROWS or os.environ.get("F5B_CUDA_GRAPH") == "1")
    L.prof_reset(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        a = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_timed = int(sum(v["launches"] for v in L.prof_read().values()))
    prof, prof_steps = {}, 1
    if not args.no_profile:
        L.prof_reset(True)
        for _ in range(prof_steps):
            step_device()
        prof = L.prof_read()
    L.prof_reset(False)
    assert torch.isfinite(a).all(), "non-finite audio"
    # e2e through the public API with host buffers
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])

    extras = default_workload_extras if (args.workload == "cfg2" and not args.ragged) else None
    if rank != 0:
        del model, voc
        torch.cuda.empty_cache()
        if extras is not None:
            extras(args, None, rank, local_rank, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" if peaks else "fallback"
    peak_gbs = peaks.get("hbm_gbs", 6650.0)

    kinds = {}
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        d = {"launches_per_step": v["launches"] / prof_steps, "ms_per_step": v["ms"] / prof_steps}
        if v["ms"] > 0:
            if v["flops"] > 0:
                d["tflops"] = v["flops"] / (v["ms"] * 1e-3) / 1e12
            d["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            d["share"] = v["ms"] / sum(x["ms"] for x in prof.values())
        kinds[k] = d
    roofline = None
    timed = {k: v for k, v in prof.items() if v["ms"] > 0}
    if timed:
        top = max(timed, key=lambda k: timed[k]["ms"])
        v = timed[top]
        if v["flops"] > 0:
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "traffic": None, "peak_source": peak_src, "launches": v["launches"], "avg_launch_ms": v["ms"] / v["launches"],
                        "flops_per_launch": v["flops"] / v["launches"]}
        else:
            ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                        "traffic": None, "peak_source": peak_src, "launches": v["launches"], "avg_launch_ms": v["ms"] / v["launches"]}

    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if roofline is not None and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == args.workload and roofline["kernel"] in tj:
            roofline["traffic"] = tj[roofline["kernel"]]["bytes_per_launch"]
            roofline["traffic_source"] = "ncu --set full capture, profiles/r02_traffic.json"
            roofline["algorithmic_bytes_per_launch"] = prof[roofline["kernel"]]["bytes"] / max(prof[roofline["kernel"]]["launches"], 1)
    total_frames = frames_per_step * args.steps * world
    value = total_frames / (ms * 1e-3)
    e2e_value = total_frames / e2e_s
    launches = launches_timed
    config["cuda_graph"] = "one captured graph per ODE step (replayed 32x per sample)" if graph_mode else "off"
    from eraxvif5tts_b200.model.backbones import dit as _dit
    _senv = os.environ.get("F5B_SPLIT_CFG", "")
    config["cfg_branches"] = ("two concurrent chains (cond / uncond forwards forked inside the captured step; launch-bound shapes)"
                              if graph_mode and _senv != "0" and (_senv == "1" or 2 * B * total <= _dit.SPLIT_CFG_MAX_ROWS)
                              else "one fused 2B-row batch")
    config["profiling"] = f"timed region un-instrumented; kernel classes / roofline from {prof_steps} extra event-bracketed eager step(s) behind it"
    config["dependent_launch"] = ("on (fused batch <= 16384 rows: kernel prologues overlap the previous kernel's tail)" if 2 * B * total <= L.PDL_MAX_ROWS
                                  else "off (GPU-bound batch; measured 1.8 % slower with it)")
    fwd_flops = dit_flops_per_forward(cfg, 2 * B, total) * NFE * reps
    line = {"metric": "mel_frames_per_sec", "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "cfg4" else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "mel-frames/s", "h2d_bytes_per_step": reps * (wav_h.numel() * 4 + text_h.numel() * 8),
                    "d2h_bytes_per_step": reps * audio_h.numel() * 4, "ms_per_step": e2e_s * 1e3 / args.steps},
            "gpu_launches": launches,
            "rtf": (ms * 1e-3 / args.steps) / (gen_frames_per_step * 256 / 24000.0),
            "gen_frames_per_sec": gen_frames_per_step * args.steps * world / (ms * 1e-3),
            "dit_tflops_per_gpu": fwd_flops / (ms * 1e-3 / args.steps) / 1e12,
            "roofline": roofline, "kernels": kinds}
    del model, voc
    torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_measure(cfg, ref_frames, total, 2, 1, full=(args.workload == "cfg1"))
        line["cpu_baseline"] = {"value": r["value"], "unit": "mel-frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if world == 1 and (args.torch_eager_gpu or (extras is not None and not args.no_eager)):
        line["torch_eager_gpu"] = torch_eager_gpu_measure(cfg, B, ref_frames, total, dev)
    if extras is not None:
        extras(args, line, rank, local_rank, world, dev)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def default_workload_extras(args, line, rank, local_rank, world, dev):
    """Sub-records the default `python bench.py --gpus N` line carries besides the headline (every rank calls this; rank 0 holds
    `line`):
      train  : the cfg-5 training step (32 x 1200 frames per GPU; forward + backward + NCCL gradient all-reduce over the N ranks +
               clip + AdamW + EMA) — ms/step, frames/s, model TFLOP/s, all-reduce mode; at N = 1 with its PyTorch-eager comparator;
      parity : (N = 1) the north-star acceptance check at the BASELINE configuration — F5TTS_Base depth 22, NFE 32, sway, CFG, ragged
               batch — product vs the fp32 oracle on this GPU (oracle/acceptance.py; the oracle is the checker, after all timing)."""
    if not args.no_train:
        arch_kw, B, _, n, desc = WORKLOADS["cfg5"]
        t = measure_train(args, "cfg5", Arch(**arch_kw), B, n, desc, rank, local_rank, world, dev, steps=max(args.steps, 3), warmup=3,
                          cpu_baseline=False, eager=(world == 1 and not args.no_eager))
        if line is not None:
            keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "config", "e2e", "gpu_launches",
                    "model_tflops_per_gpu", "params", "loss", "roofline", "kernels", "torch_eager_gpu")
            line["train"] = {k: t[k] for k in keep if k in t}
    if line is not None and world == 1 and not args.no_parity:
        line["parity"] = parity_leg(dev)


def parity_leg(dev):
    from oracle import acceptance as A
    from oracle import f5_oracle as O
    from oracle.weights import make_dit_state_dict
    from eraxvif5tts_b200.model import CFM, DiT
    cfg = O.DiTConfig()  # F5TTS_Base: dim 1024, depth 22, heads 16
    sd = make_dit_state_dict(cfg, 0)
    tr = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim,
             text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, text_mask_padding=cfg.text_mask_padding,
             conv_layers=cfg.conv_layers, pe_attn_head=cfg.pe_attn_head)
    model = CFM(transformer=tr, mel_spec_kwargs=dict(n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=cfg.mel_dim,
                                                     target_sample_rate=24000, mel_spec_type="vocos"), odeint_kwargs=dict(method="euler"))
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).eval()
    res = A.sample_parity(model, sd, cfg, 563, [1875, 1610], steps=NFE, cfg_strength=CFG, sway=SWAY, device=dev)
    res["tolerances"] = {"velocity_max_abs_bf16": A.VEL_TOL_BF16, "mel_mean_abs": A.MEL_MEAN_TOL}
    res["pass"] = bool(res["velocity_max_abs"] <= A.VEL_TOL_BF16 and res["mel_mean_abs_generated"] <= A.MEL_MEAN_TOL)
    res["oracle"] = "oracle/f5_oracle.py (fp32, cuBLAS / cuDNN without TF32) on the same GPU, identical bf16-exact weights, noise and inputs"
    # the tf32 operand mode (north_star's fp32 clause: 1e-3 relative error of the velocity field): same check, then its throughput on
    # cfg-2's batch — device-resident inputs, CUDA events, un-instrumented, 2 warm-up + 3 timed sample() calls
    model.transformer.set_precision("tf32")
    r32 = A.sample_parity(model, sd, cfg, 563, [1875, 1610], steps=NFE, cfg_strength=CFG, sway=SWAY, device=dev)
    t = {"precision": "tf32 tensor-core operands (kind::tf32), fp32 activations", "velocity_rel_fro": r32["velocity_rel_fro"],
         "velocity_max_abs": r32["velocity_max_abs"], "mel_mean_abs_generated": r32["mel_mean_abs_generated"],
         "tolerance_velocity_rel_fro": A.VEL_RTOL_FP32,
         "pass": bool(r32["velocity_rel_fro"] <= A.VEL_RTOL_FP32 and r32["mel_mean_abs_generated"] <= A.MEL_MEAN_TOL)}
    _, B, ref_frames, total, _ = WORKLOADS["cfg2"]
    cond, text, duration, lens, _ = make_inputs(Arch(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads), B, ref_frames, total, 1234)
    cond, text, duration, lens = cond.to(dev), text.to(dev), duration.to(dev), lens.to(dev)

    def one():
        model.sample(cond=cond, text=text, duration=duration, lens=lens, steps=NFE, cfg_strength=CFG, sway_sampling_coef=SWAY, seed=0,
                     return_trajectory=False)
    for _ in range(2):
        one()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(3):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 3
    t.update({"workload": "cfg2 sampling (CFM.sample only, no vocoder)", "ms_per_step": ms, "value": B * total / (ms * 1e-3),
              "unit": "mel-frames/s", "model_tflops": dit_flops_per_forward(cfg, 2 * B, total) * NFE / (ms * 1e-3) / 1e12})
    res["tf32_mode"] = t
    return res


if __name__ == "__main__":
    main()
