"""CFM with the reference's constructor and `sample` signature (/root/reference/src/f5_tts/model/cfm.py:30-208).

Host logic (durations, masks, noise, sway-sampled time grid) follows the reference line by line in torch; the ODE loop is
re-designed for the B200 path:
  * the time MLP and every AdaLN modulation of the whole schedule are ONE GEMM before the loop (t is shared by the batch);
  * TextEmbedding and the step-invariant [cond | text] part of InputEmbedding.proj are computed once per branch;
  * the cond and uncond forwards of classifier-free guidance run as ONE 2B-row batch through the DiT kernels;
  * CFG combine + Euler update + bf16 re-pack of the state is one kernel.
"""
from __future__ import annotations

import os
from typing import Callable

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.utils.rnn import pad_sequence

from .. import _lib as L
from .. import ops
from .modules import MelSpec
from .utils import default, exists, lens_to_mask, list_str_to_idx, list_str_to_tensor

f32, bf16 = torch.float32, torch.bfloat16

# One captured CUDA graph per ODE step (164 kernel launches + the CFG/Euler update), replayed for every step and every later
# sample() call of the same shape: decisive for launch-bound small batches (the reference's serial B=1 chunk loop: 8.2k -> 13.1k
# frames/s) and still worth ~1 % at batch 16.  F5B_CUDA_GRAPH=0 disables it; per-launch event profiling disables it too.
GRAPH_MAX_ROWS = 1 << 30


class CFM(nn.Module):
    def __init__(self, transformer: nn.Module, sigma=0.0, odeint_kwargs: dict = dict(method="euler"), audio_drop_prob=0.35,
                 cond_drop_prob=0.25, num_channels=None, mel_spec_module: nn.Module | None = None, mel_spec_kwargs: dict = dict(),
                 frac_lengths_mask: tuple[float, float] = (0.7, 1.0), vocab_char_map: dict | None = None):
        super().__init__()
        self.frac_lengths_mask = frac_lengths_mask
        self.mel_spec = default(mel_spec_module, MelSpec(**mel_spec_kwargs))
        self.num_channels = default(num_channels, self.mel_spec.n_mel_channels)
        self.audio_drop_prob = audio_drop_prob  # reference values 0.35 / 0.25 (cfm.py:42-43)
        self.cond_drop_prob = cond_drop_prob
        self.transformer = transformer
        self.dim = transformer.dim
        self.sigma = sigma
        self.odeint_kwargs = odeint_kwargs
        self.vocab_char_map = vocab_char_map

    @property
    def device(self):
        return next(self.parameters()).device

    @L.on_own_device
    @torch.no_grad()
    def sample(self, cond, text, duration, *, lens=None, steps=32, cfg_strength=1.0, sway_sampling_coef=None, seed: int | None = None,
               max_duration=4096, vocoder: Callable | None = None, no_ref_audio=False, duplicate_test=False, t_inter=0.1,
               edit_mask=None, noise: torch.Tensor | None = None, return_trajectory: bool = True):
        """cfm.py:82-208.  Extensions (keyword-only, default = reference behaviour): `noise` [b, n, mel] replaces the internally
        drawn y0 (parity tests against a CPU oracle need identical noise); `return_trajectory=False` returns only the last state
        in `trajectory` (saves steps x b x n x mel floats)."""
        self.eval()
        eng = self.transformer.engine()
        device = eng.device
        method = self.odeint_kwargs.get("method", "euler")
        if method not in ("euler", "midpoint"):
            raise NotImplementedError(f"odeint method {method!r}: the reference uses fixed-grid euler or midpoint (cfm.py:37)")

        # raw wave -> mel (cfm.py:103-106)
        cond = cond.to(device)
        if cond.ndim == 2:
            cond = self.mel_spec.forward_token_major(cond)
            assert cond.shape[-1] == self.num_channels
        cond = cond.to(f32)
        batch, cond_seq_len = cond.shape[:2]
        if not exists(lens):
            lens = torch.full((batch,), cond_seq_len, device=device, dtype=torch.long)
        lens = lens.to(device)

        # text (cfm.py:116-121)
        if isinstance(text, list):
            if exists(self.vocab_char_map):
                text = list_str_to_idx(text, self.vocab_char_map).to(device)
            else:
                text = list_str_to_tensor(text).to(device)
            assert text.shape[0] == batch
        text = text.to(device)

        # duration (cfm.py:125-136)
        cond_mask = lens_to_mask(lens)
        if edit_mask is not None:
            cond_mask = cond_mask & edit_mask.to(device)
        if isinstance(duration, int):
            duration = torch.full((batch,), duration, device=device, dtype=torch.long)
        duration = duration.to(device)
        duration = torch.maximum(torch.maximum((text != -1).sum(dim=-1), lens) + 1, duration)
        duration = duration.clamp(max=max_duration)
        dur_host = duration.tolist()  # one D2H sync per sample(); the reference syncs here too (int(dur) in randn)
        n = max(dur_host)

        if duplicate_test:
            test_cond = F.pad(cond, (0, 0, cond_seq_len, n - 2 * cond_seq_len), value=0.0)
        cond = F.pad(cond, (0, 0, 0, n - cond_seq_len), value=0.0)
        if no_ref_audio:
            cond = torch.zeros_like(cond)
        cond_mask = F.pad(cond_mask, (0, n - cond_mask.shape[-1]), value=False).unsqueeze(-1)
        step_cond = torch.where(cond_mask, cond, torch.zeros_like(cond))
        # key-padding mask only for batch > 1 (cfm.py:152-155), as a per-row length
        lens32 = duration.to(torch.int32).contiguous() if batch > 1 else None

        # noise (cfm.py:178-183): per item, re-seeding the global generator
        if noise is None:
            y0 = []
            for dur in dur_host:
                if exists(seed):
                    torch.manual_seed(seed)
                y0.append(torch.randn(dur, self.num_channels, device=device, dtype=f32))
            y0 = pad_sequence(y0, padding_value=0, batch_first=True)
        else:
            y0 = noise.to(device=device, dtype=f32).clone()
            assert y0.shape == (batch, n, self.num_channels)

        t_start = 0
        if duplicate_test:
            t_start = t_inter
            y0 = (1 - t_start) * y0 + t_start * test_cond
            steps = int(steps * (1 - t_start))
        t = torch.linspace(t_start, 1, steps + 1, device=device, dtype=f32)
        if sway_sampling_coef is not None:
            t = t + sway_sampling_coef * (torch.cos(torch.pi / 2 * t) - 1 + t)

        # ---- step-invariant work -------------------------------------------------------------------------------------
        use_cfg = cfg_strength >= 1e-5
        Bf = 2 * batch if use_cfg else batch
        genv = os.environ.get("F5B_CUDA_GRAPH", "")
        use_graph = (method == "euler" and genv != "0" and (genv == "1" or Bf * n <= GRAPH_MAX_ROWS)
                     and not L.load().f5b_prof_enabled())
        # dependent launches (kernel prologues under the previous kernel's tail) pay in the launch-bound regime only
        L.set_dependent_launch(Bf * n <= L.PDL_MAX_ROWS)
        # the reference's SDPA dropout at inference (DiT.set_attn_dropout, default off): one mask stream per ODE evaluation
        attn_p = float(getattr(self.transformer, "attn_dropout_p", 0.0))
        attn_base = (int(seed) if exists(seed) else int(torch.randint(0, 2 ** 31, (1,)).item())) if attn_p > 0 else 0
        sess = eng.step_session(batch, Bf, n, lens32 is not None, attn_p) if use_graph else None
        c0 = sess["c0"] if use_graph else torch.empty(Bf, n, eng.dim, dtype=f32, device=device)
        te_c = eng.text_embed(text, n, False)
        eng.input_const(step_cond, te_c, out=c0[:batch])
        if use_cfg:
            te_u = eng.text_embed(text, n, True)
            eng.input_const(None, te_u, out=c0[batch:])
        dt = t[1:] - t[:-1]
        if method == "euler":
            times = t[:-1]
        else:
            times = torch.stack((t[:-1], t[:-1] + 0.5 * dt), dim=1).reshape(-1)
        mod = eng.modulation(times)  # [evals, mod_dim]
        dt_host = dt.tolist()

        rows = batch * n
        if use_graph:
            y, yb, pred = sess["y"], sess["yb"], sess["pred"]
            y.copy_(y0)
            if lens32 is not None:
                sess["lens"].copy_(lens32)
            lens32 = sess["lens"]
        else:
            y = y0.contiguous()
            yb = torch.empty(rows, 128, dtype=eng.act_dtype, device=device)
            pred = torch.empty(Bf, n, self.num_channels, dtype=f32, device=device)
        eng.pack_state(y, yb)
        tf32 = eng.precision == "tf32"

        def euler(state, step):
            # CFG combine + Euler update; the bf16 mode refreshes the packed state in the same kernel, the tf32 mode re-packs in fp32
            ops.cfg_euler(state, pc, pu, cfg_strength, step, None if tf32 else yb)
            if tf32:
                eng.pack_state(state, yb)
        pc = pred[:batch]
        pu = pred[batch:] if use_cfg else None
        traj = [y.clone()] if return_trajectory else None
        ymid = torch.empty_like(y) if method == "midpoint" else None
        if use_graph:
            # per step: [modulation row | cfg | dt] -> stepbuf (one small D2D copy), then replay the captured step
            # (the dropout word travels as the mantissa of a float in [1, 2): any copy preserves its bits)
            words = torch.tensor([0x3F800000 | ((attn_base * 0x9E3779B1 + i * 0x85EBCA77) & 0x7FFFFF) for i in range(steps)],
                                 dtype=torch.int32, device=device).view(f32).unsqueeze(1)
            table = torch.cat((mod, torch.full((steps, 1), float(cfg_strength), device=device), dt.unsqueeze(1), words), dim=1).contiguous()

        def adrop(k):
            return (attn_p, attn_base * 1000003 + k, None) if attn_p > 0 else None

        # ---- ODE loop (fn closure cfm.py:159-173 + torchdiffeq fixed-grid solver) ---------------------------------------
        for i in range(steps):
            if use_graph:
                sess["stepbuf"].copy_(table[i], non_blocking=True)
                sess["graph"].replay()
                L.prof_add(sess["delta"])
            elif method == "euler":
                eng.forward(yb, batch, c0, Bf, n, mod[i], 0, lens32, pred, attn_drop=adrop(2 * i))
                euler(y, dt_host[i])
            else:
                eng.forward(yb, batch, c0, Bf, n, mod[2 * i], 0, lens32, pred, attn_drop=adrop(2 * i))
                ymid.copy_(y)
                euler(ymid, 0.5 * dt_host[i])
                eng.forward(yb, batch, c0, Bf, n, mod[2 * i + 1], 0, lens32, pred, attn_drop=adrop(2 * i + 1))
                euler(y, dt_host[i])
            if return_trajectory:
                traj.append(y.clone())
        self.transformer.clear_cache()

        if use_graph:
            y = y.clone()  # the session buffer is reused by the next call
        trajectory = torch.stack(traj) if return_trajectory else y.unsqueeze(0)
        out = torch.where(cond_mask, cond, y)
        if exists(vocoder):
            out = out.permute(0, 2, 1)
            out = vocoder(out)
        return out, trajectory

    @L.on_own_device
    @torch.no_grad()
    def forward(self, inp, text, *, lens=None, noise_scheduler=None, draws: dict | None = None):
        """Flow-matching loss (cfm.py:210-283): returns (loss, cond, pred) like the reference, computed by the CUDA library.  The
        tensors carry no autograd graph: training goes through `eraxvif5tts_b200.train.TrainEngine.loss_and_grads`, which runs this
        same forward in its activation-saving form followed by the hand-written backward.
        `draws` (extension) fixes the random choices for parity tests: rand_span_mask, x0, time, drop_audio_cond, drop_text."""
        from random import random
        from .utils import mask_from_frac_lengths
        eng = self.transformer.engine()
        device = eng.device
        inp = inp.to(device)
        if inp.ndim == 2:
            inp = self.mel_spec.forward_token_major(inp)
            assert inp.shape[-1] == self.num_channels
        x1 = inp.to(f32).contiguous()
        batch, seq_len = x1.shape[:2]
        if isinstance(text, list):
            if exists(self.vocab_char_map):
                text = list_str_to_idx(text, self.vocab_char_map).to(device)
            else:
                text = list_str_to_tensor(text).to(device)
            assert text.shape[0] == batch
        text = text.to(device)
        if not exists(lens):
            lens = torch.full((batch,), seq_len, device=device)
        lens = lens.to(device)
        mask = lens_to_mask(lens, length=seq_len)
        draws = draws or {}
        if "rand_span_mask" in draws:
            rand_span_mask = draws["rand_span_mask"].to(device)
        else:
            frac_lengths = torch.zeros((batch,), device=device).float().uniform_(*self.frac_lengths_mask)
            rand_span_mask = mask_from_frac_lengths(lens, frac_lengths)
        rand_span_mask = rand_span_mask & mask
        x0 = draws["x0"].to(device=device, dtype=f32).contiguous() if "x0" in draws else torch.randn_like(x1)
        time = draws["time"].to(device=device, dtype=f32).contiguous() if "time" in draws else torch.rand((batch,), dtype=f32, device=device)
        if "drop_audio_cond" in draws:
            drop_audio_cond, drop_text = bool(draws["drop_audio_cond"]), bool(draws["drop_text"])
        else:
            drop_audio_cond = random() < self.audio_drop_prob  # per BATCH, Python RNG (cfm.py:266-271)
            if random() < self.cond_drop_prob:
                drop_audio_cond, drop_text = True, True
            else:
                drop_text = False
        lib = L.load()
        phi, flow, cond = torch.empty_like(x1), torch.empty_like(x1), torch.empty_like(x1)
        span_u8 = rand_span_mask.to(torch.uint8).contiguous()
        L.check(lib.f5b_fm_prepare(x1.data_ptr(), x0.data_ptr(), time.data_ptr(), span_u8.data_ptr(), phi.data_ptr(), flow.data_ptr(),
                                   cond.data_ptr(), batch, seq_len, self.num_channels, L.stream()), "f5b_fm_prepare")
        # no mask is passed to the transformer in training (cfm.py:275-277)
        pred = self.transformer(x=phi, cond=cond, text=text, time=time, drop_audio_cond=drop_audio_cond, drop_text=drop_text)
        ws = torch.empty(2048, dtype=f32, device=device)
        out2 = torch.empty(2, dtype=f32, device=device)
        L.check(lib.f5b_masked_mse(pred.data_ptr(), flow.data_ptr(), span_u8.data_ptr(), ws.data_ptr(), out2.data_ptr(), batch * seq_len,
                                   self.num_channels, L.stream()), "f5b_masked_mse")
        return out2[0], cond, pred
