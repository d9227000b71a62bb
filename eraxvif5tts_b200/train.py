"""Training step of the hot path: CFM.forward -> loss -> backward -> (all-reduce) -> clip + AdamW + EMA
(/root/reference/src/f5_tts/model/cfm.py:210-283, /root/reference/src/f5_tts/model/trainer.py:1250-1287, 1321), re-designed
around three flat device buffers laid out in the ORDER THE KERNELS WANT:

    master  fp32  every parameter, grouped into the fused tensors of F5bDitDesc (q|k|v weights of a block adjacent, the 22 blocks
                  stacked, all AdaLN linears stacked ...); the nn.Parameters of the DiT module (reference key names, so
                  state_dict()/checkpoints are unchanged) become views into it
    mirror  bf16  same layout, written by the fused AdamW kernel in the same pass that updates the master -> the GEMM operands
    grads   fp32  same layout; the wgrad GEMMs TMA-reduce-add straight into it, the DDP average is ONE all-reduce over it

so a step has no per-parameter host work at all.  The two tensors whose kernel layout is not a view of the parameter
(InputEmbedding.proj split by source, the grouped conv weights) are re-packed by small kernels after the update.
No autograd graph and no PyTorch math is involved; torch provides device memory, streams and torch.distributed."""
from __future__ import annotations

import ctypes as C
import os
from random import random

import torch

from . import _lib as L
from .optim import EmaSchedule

bf16, f32 = torch.bfloat16, torch.float32


def _segments(dit):
    """(desc field, [parameters in kernel order], mirror?) — mirror=True: the kernels read the bf16 copy"""
    sd = dict(dit.named_parameters())
    depth, Lc = dit.depth, dit.conv_layers
    blk = "transformer_blocks.{}."
    tb = "text_embed.text_blocks.{}."

    def per(fmt, n):
        return [sd[fmt.format(i)] for i in range(n)]

    seg = [
        ("time_w0", [sd["time_embed.time_mlp.0.weight"]], True), ("time_b0", [sd["time_embed.time_mlp.0.bias"]], False),
        ("time_w2", [sd["time_embed.time_mlp.2.weight"]], True), ("time_b2", [sd["time_embed.time_mlp.2.bias"]], False),
        ("mod_w", per(blk + "attn_norm.linear.weight", depth) + [sd["norm_out.linear.weight"]], True),
        ("mod_b", per(blk + "attn_norm.linear.bias", depth) + [sd["norm_out.linear.bias"]], False),
        ("text_table", [sd["text_embed.text_embed.weight"]], False),
    ]
    if Lc > 0:
        seg += [
            ("tb_dw_w", per(tb + "dwconv.weight", Lc), False), ("tb_dw_b", per(tb + "dwconv.bias", Lc), False),
            ("tb_ln_w", per(tb + "norm.weight", Lc), False), ("tb_ln_b", per(tb + "norm.bias", Lc), False),
            ("tb_pw1_w", per(tb + "pwconv1.weight", Lc), True), ("tb_pw1_b", per(tb + "pwconv1.bias", Lc), False),
            ("tb_grn_g", per(tb + "grn.gamma", Lc), False), ("tb_grn_b", per(tb + "grn.beta", Lc), False),
            ("tb_pw2_w", per(tb + "pwconv2.weight", Lc), True), ("tb_pw2_b", per(tb + "pwconv2.bias", Lc), False),
        ]
    cw = "input_embed.conv_pos_embed.conv1d.{}."
    seg += [
        ("_in_w", [sd["input_embed.proj.weight"]], False), ("in_b", [sd["input_embed.proj.bias"]], False),
        ("_cp_w1", [sd[cw.format(0) + "weight"]], False), ("cp_b1", [sd[cw.format(0) + "bias"]], False),
        ("_cp_w2", [sd[cw.format(2) + "weight"]], False), ("cp_b2", [sd[cw.format(2) + "bias"]], False),
        ("qkv_w", [sd[blk.format(i) + f"attn.to_{c}.weight"] for i in range(depth) for c in "qkv"], True),
        ("qkv_b", [sd[blk.format(i) + f"attn.to_{c}.bias"] for i in range(depth) for c in "qkv"], False),
        ("out_w", per(blk + "attn.to_out.0.weight", depth), True), ("out_b", per(blk + "attn.to_out.0.bias", depth), False),
        ("ff1_w", per(blk + "ff.ff.0.0.weight", depth), True), ("ff1_b", per(blk + "ff.ff.0.0.bias", depth), False),
        ("ff2_w", per(blk + "ff.ff.2.weight", depth), True), ("ff2_b", per(blk + "ff.ff.2.bias", depth), False),
        ("proj_w", [sd["proj_out.weight"]], True), ("proj_b", [sd["proj_out.bias"]], False),
    ]
    covered = {id(p) for _, ps, _ in seg for p in ps}
    missing = [k for k, p in sd.items() if id(p) not in covered]
    if missing:
        raise L.F5bError(f"TrainEngine: parameters without a kernel layout: {missing}")
    return seg


class TrainEngine:
    """One data-parallel replica of the training step for a `CFM` whose transformer is a `DiT`."""

    def __init__(self, cfm, lr: float = 7.5e-5, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0, with_ema=False,
                 ema_schedule: EmaSchedule | None = None, dropout: float = 0.0, checkpoint_activations: bool | None = None,
                 attn_dropout: float = 0.0):
        """checkpoint_activations (default: the DiT's own `checkpoint_activations` flag, dit.py:121,158,221-223): every block keeps only
        its fp32 input and the backward re-runs the block's forward (4.6 GB of saved activations instead of 29.6 GB at 32 x 1200
        frames; one extra forward per step; gradients bit-identical).
        dropout: the DiT's train-mode dropout (the reference builds DiT(dropout=0.1), model/backbones/dit.py:132), applied after
        FeedForward's GELU and behind attention's to_out (modules.py:342-353, :436-440).
        attn_dropout: the dropout inside scaled_dot_product_attention (modules.py:490; the fork hard-codes dropout_p = 0.1), on the
        normalised attention probabilities; default 0 (the parity setting), same counter-based mask generator and seed.
        0 (default) is the setting the gradient-parity tests run at."""
        self.cfm, self.dit = cfm, cfm.transformer
        self.dropout = float(dropout)
        self.attn_dropout = float(attn_dropout)
        self.checkpoint_activations = bool(getattr(cfm.transformer, "checkpoint_activations", False)
                                           if checkpoint_activations is None else checkpoint_activations)
        self.last_losses = None
        dit = self.dit
        dev = dit.proj_out.weight.device
        if dev.type != "cuda":
            raise L.F5bError("TrainEngine needs the model on a CUDA device (B200); there is no CPU fallback")
        self.device, self.lib = dev, L.load()
        seg = _segments(dit)
        self.offsets, off = {}, 0
        for name, ps, _ in seg:
            off = (off + 7) // 8 * 8  # 16-byte aligned bf16 views (TMA base addresses)
            self.offsets[name] = off
            off += sum(p.numel() for p in ps)
        self.n = n = (off + 7) // 8 * 8
        self.p = torch.zeros(n, dtype=f32, device=dev)
        self.g = torch.zeros(n, dtype=f32, device=dev)
        self.m = torch.zeros(n, dtype=f32, device=dev)
        self.v = torch.zeros(n, dtype=f32, device=dev)
        self.mirror = torch.zeros(n, dtype=bf16, device=dev)
        self.params = []
        with torch.no_grad():
            for name, ps, _ in seg:
                o = self.offsets[name]
                for p in ps:
                    s = p.numel()
                    self.p[o:o + s].copy_(p.detach().reshape(-1))
                    p.data = self.p[o:o + s].view_as(p)   # parameters (and their .grad) become views of the flat buffers
                    p.grad = self.g[o:o + s].view_as(p)
                    self.params.append((p, o, s))
                    o += s
            self.mirror.copy_(self.p)
        self.ema = self.p.clone() if with_ema else None
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.ema_schedule = ema_schedule or EmaSchedule()
        self.step_count, self.ema_calls = 0, 0
        self._red = torch.empty(1024, dtype=f32, device=dev)
        self._sumsq = torch.zeros(1, dtype=f32, device=dev)

        D, mel, T = dit.dim, dit.mel_dim, dit.text_dim
        self.D, self.mel, self.T = D, mel, T
        self.in_w = dict(dit.named_parameters())["input_embed.proj.weight"]
        self.in_wx = torch.zeros(D, 128, dtype=bf16, device=dev)
        self.in_wct = torch.zeros(D, 128 + T, dtype=bf16, device=dev)
        self.g_in_wx = torch.zeros(D, 128, dtype=f32, device=dev)
        self.g_in_wct = torch.zeros(D, 128 + T, dtype=f32, device=dev)
        npk = self.lib.f5b_convpos_packed_elems(D, 16, 31)
        self.cp = {k: torch.zeros(npk, dtype=bf16, device=dev) for k in ("w1", "w2", "w1_t", "w2_t")}

        d = L.DitDesc()
        d.dim, d.depth, d.heads, d.dim_head, d.ff_mult = D, dit.depth, dit.heads, dit.dim_head, dit.ff_mult
        d.mel_dim, d.text_dim, d.conv_layers = mel, T, dit.conv_layers
        d.rope_heads = dit.heads if dit.pe_attn_head is None else int(dit.pe_attn_head)
        d.text_mask_padding = int(bool(dit.text_mask_padding))
        d.convpos_kernel, d.convpos_groups = 31, 16
        d.vocab_rows = dit.text_embed.text_embed.weight.shape[0]
        gr = L.DitGrads()
        for name, _, mirrored in seg:
            o = self.offsets[name]
            if name.startswith("_"):
                continue
            setattr(d, name, (self.mirror if mirrored else self.p)[o:].data_ptr())
            setattr(gr, name, self.g[o:].data_ptr())
        self.text_pos = (dit.text_embed.freqs_cis.to(device=dev, dtype=f32).contiguous() if dit.conv_layers > 0
                         else torch.zeros(1, T, dtype=f32, device=dev))
        d.text_pos = self.text_pos.data_ptr()
        d.in_wx, d.in_wct = self.in_wx.data_ptr(), self.in_wct.data_ptr()
        d.cp_w1, d.cp_w2 = self.cp["w1"].data_ptr(), self.cp["w2"].data_ptr()
        gr.in_wx, gr.in_wct = self.g_in_wx.data_ptr(), self.g_in_wct.data_ptr()
        gr.cp_w1, gr.cp_w2 = self.g[self.offsets["_cp_w1"]:].data_ptr(), self.g[self.offsets["_cp_w2"]:].data_ptr()
        self.desc, self.grads = d, gr
        h = L.vp()
        L.check(self.lib.f5b_dit_create(C.byref(d), C.byref(h)), "f5b_dit_create")
        self.handle = h
        self._ws = None
        self._rope = {}
        # gradient segments that are complete once a block has been differentiated (stacked [depth, ...] tensors) and the rest
        F = dit.ff_mult * D
        self._stacked = [("qkv_w", 3 * D * D), ("qkv_b", 3 * D), ("out_w", D * D), ("out_b", D), ("ff1_w", F * D), ("ff1_b", F),
                         ("ff2_w", D * F), ("ff2_b", D)]
        covered = sorted((self.offsets[k], self.offsets[k] + dit.depth * per) for k, per in self._stacked)
        self._rest, pos = [], 0
        for a, b in covered:
            if a > pos:
                self._rest.append((pos, a))
            pos = b
        if pos < self.n:
            self._rest.append((pos, self.n))
        self._pending = None
        self.refresh_packed()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.f5b_dit_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------------------------------------ buffers
    def refresh_packed(self):
        """re-pack the two kernel layouts that are not views of their parameter (after every optimizer step / weight load)"""
        mel = self.mel
        with torch.no_grad():
            w = self.in_w
            self.in_wx[:, :mel].copy_(w[:, :mel])
            self.in_wct[:, :mel].copy_(w[:, mel:2 * mel])
            self.in_wct[:, 128:].copy_(w[:, 2 * mel:])
        s = L.stream()
        for j in (1, 2):
            src = self.p[self.offsets[f"_cp_w{j}"]:].data_ptr()
            L.check(self.lib.f5b_pack_convpos_weight(src, self.cp[f"w{j}"].data_ptr(), self.D, 16, 31, s), "f5b_pack_convpos_weight")
            L.check(self.lib.f5b_pack_convpos_weight_t(src, self.cp[f"w{j}_t"].data_ptr(), self.D, 16, 31, s), "f5b_pack_convpos_weight_t")

    def sync_from_master(self):
        """call after writing parameters from outside the optimizer (checkpoint load): rebuild the bf16 mirror + packed copies"""
        with torch.no_grad():
            self.mirror.copy_(self.p)
        self.refresh_packed()
        self.dit.invalidate()

    def release(self):
        """drop the activation workspace (tens of GB at training batch sizes); the next loss_and_grads re-allocates it"""
        self._ws = None

    def workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def rope_table(self, n):
        if n not in self._rope:
            inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=self.device).float() / 64))
            fr = torch.outer(torch.arange(n, device=self.device).float(), inv)
            self._rope[n] = torch.stack((fr.cos(), fr.sin()), dim=-1).contiguous()
        return self._rope[n]

    def zero_grad(self):
        self.g.zero_()
        self.g_in_wx.zero_()
        self.g_in_wct.zero_()

    # ------------------------------------------------------------------------------------------------ forward + backward
    @L.on_own_device
    @torch.no_grad()
    def loss_and_grads(self, inp, text, *, lens=None, draws: dict | None = None, overlap_allreduce: bool = False, group=None,
                       buckets: int = 4, distill: dict | None = None):
        """CFM.forward (cfm.py:210-283) followed by loss.backward(): returns (loss, cond, pred) and ACCUMULATES d loss / d theta into
        the flat gradient buffer (= every parameter's .grad).  `draws` fixes the random choices (rand_span_mask, x0, time,
        drop_audio_cond, drop_text) for parity tests.
        overlap_allreduce (use on the LAST micro-batch of an update, world > 1): the backward is issued in `buckets` groups of blocks
        from the top of the network down, and the NCCL all-reduce of each group's finished gradient segments is launched while the
        groups below are still being differentiated (DDP's bucketed overlap); `allreduce_grads()` then only waits.
        distill = dict(teacher=<CFM or DiT>, alpha=0.5, loss_type="mse" | "l1", spec_l1_weight=0.0): the distillation step of
        train/distil_reload.py:1044-1093 in one launch sequence -- the frozen teacher's inference forward (always conditioned:
        drop_audio_cond = drop_text = False) on the same (x_t, cond, text, time), then the student's train forward, and
        total = (1 - alpha) student + alpha distill + spec_l1_weight spec_l1 differentiated w.r.t. the student.  Returns the total
        loss; `self.last_losses` holds the device tensor (total, student, distill, spec_l1, masked frames)."""
        from .model.utils import exists, lens_to_mask, list_str_to_idx, list_str_to_tensor, mask_from_frac_lengths
        cfm, lib, dev = self.cfm, self.lib, self.device
        inp = inp.to(dev)
        if inp.ndim == 2:
            inp = cfm.mel_spec.forward_token_major(inp)
        x1 = inp.to(f32).contiguous()
        B, n = x1.shape[:2]
        if isinstance(text, list):
            text = (list_str_to_idx(text, cfm.vocab_char_map) if exists(cfm.vocab_char_map) else list_str_to_tensor(text))
        text = text.to(dev).long().contiguous()
        if not exists(lens):
            lens = torch.full((B,), n, device=dev)
        lens = lens.to(dev)
        mask = lens_to_mask(lens, length=n)
        draws = draws or {}
        if "rand_span_mask" in draws:
            span = draws["rand_span_mask"].to(dev)
        else:
            frac = torch.zeros((B,), device=dev).float().uniform_(*cfm.frac_lengths_mask)
            span = mask_from_frac_lengths(lens, frac)
        span = span & mask
        x0 = draws["x0"].to(device=dev, dtype=f32).contiguous() if "x0" in draws else torch.randn_like(x1)
        time = draws["time"].to(device=dev, dtype=f32).contiguous() if "time" in draws else torch.rand((B,), dtype=f32, device=dev)
        if "drop_audio_cond" in draws:
            drop_audio_cond, drop_text = bool(draws["drop_audio_cond"]), bool(draws["drop_text"])
        else:
            drop_audio_cond = random() < cfm.audio_drop_prob  # per BATCH, Python RNG (cfm.py:266-271)
            if random() < cfm.cond_drop_prob:
                drop_audio_cond, drop_text = True, True
            else:
                drop_text = False
        s = L.stream()
        C_ = cfm.num_channels
        phi, flow, cond = torch.empty_like(x1), torch.empty_like(x1), torch.empty_like(x1)
        span_u8 = span.to(torch.uint8).contiguous()
        L.check(lib.f5b_fm_prepare(x1.data_ptr(), x0.data_ptr(), time.data_ptr(), span_u8.data_ptr(), phi.data_ptr(), flow.data_ptr(),
                                   cond.data_ptr(), B, n, C_, s), "f5b_fm_prepare")
        te = torch.empty(B * n, self.T, dtype=f32, device=dev)
        tws = torch.empty(lib.f5b_dit_text_train_ws_bytes(self.handle, B, n), dtype=torch.uint8, device=dev)
        L.check(lib.f5b_dit_text_embed_train(self.handle, text.data_ptr(), text.shape[1], B, n, int(drop_text), te.data_ptr(),
                                             tws.data_ptr(), tws.numel(), s), "f5b_dit_text_embed_train")
        L.check(lib.f5b_train_set_checkpoint(int(self.checkpoint_activations)), "f5b_train_set_checkpoint")
        nbytes = lib.f5b_dit_train_ws_bytes(self.handle, B, n)
        ws = self.workspace(nbytes)
        rope = self.rope_table(n)
        pred = torch.empty(B, n, C_, dtype=f32, device=dev)
        L.set_dependent_launch(False)  # GPU-bound: A/B 105.3 / 106.4 ms off vs 106.4 / 105.9 ms on (include/f5b200.h)
        # the dropout masks of this micro-step are a function of this seed; the backward below regenerates them
        seed = int(draws["dropout_seed"]) if "dropout_seed" in draws else int(torch.randint(0, 2 ** 62, (1,)).item())
        L.check(lib.f5b_train_set_dropout(self.dropout, seed), "f5b_train_set_dropout")
        L.check(lib.f5b_train_set_attn_dropout(self.attn_dropout), "f5b_train_set_attn_dropout")
        # no mask is passed to the transformer in training (cfm.py:275-277)
        L.check(lib.f5b_dit_train_forward(self.handle, phi.data_ptr(), None if drop_audio_cond else cond.data_ptr(), te.data_ptr(),
                                          time.data_ptr(), B, n, None, rope.data_ptr(), pred.data_ptr(), ws.data_ptr(), ws.numel(), s),
                "f5b_dit_train_forward")
        dpred = torch.empty(B * n, 128, dtype=bf16, device=dev)
        if distill is None:
            red = torch.empty(2048, dtype=f32, device=dev)
            out2 = torch.empty(2, dtype=f32, device=dev)
            L.check(lib.f5b_masked_mse(pred.data_ptr(), flow.data_ptr(), span_u8.data_ptr(), red.data_ptr(), out2.data_ptr(), B * n, C_, s),
                    "f5b_masked_mse")
            L.check(lib.f5b_mse_grad(pred.data_ptr(), flow.data_ptr(), span_u8.data_ptr(), out2.data_ptr(), dpred.data_ptr(), B * n, C_, 128,
                                     s), "f5b_mse_grad")
            self.last_losses = out2
        else:
            teacher = distill["teacher"]
            teacher = getattr(teacher, "transformer", teacher)
            kind = distill.get("loss_type", "mse")
            if kind not in ("mse", "l1"):
                raise ValueError(f"Unsupported distill_loss_type: {kind}")  # distil_reload.py:1077
            alpha, w = float(distill.get("alpha", 0.5)), float(distill.get("spec_l1_weight", 0.0))
            tpred = teacher(x=phi, cond=cond, text=text, time=time, drop_audio_cond=False, drop_text=False).to(f32).contiguous()
            red = torch.empty(4096, dtype=f32, device=dev)
            out2 = torch.empty(5, dtype=f32, device=dev)
            L.check(lib.f5b_distill_loss(pred.data_ptr(), flow.data_ptr(), tpred.data_ptr(), span_u8.data_ptr(), red.data_ptr(),
                                         out2.data_ptr(), B * n, C_, int(kind == "l1"), alpha, w, s), "f5b_distill_loss")
            L.check(lib.f5b_distill_grad(pred.data_ptr(), flow.data_ptr(), tpred.data_ptr(), span_u8.data_ptr(), out2.data_ptr(),
                                         dpred.data_ptr(), B * n, C_, 128, int(kind == "l1"), alpha, w, s), "f5b_distill_grad")
            self.last_losses = out2
        dtext = torch.empty(B * n, self.T, dtype=bf16, device=dev)
        import torch.distributed as dist
        overlap = overlap_allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

        def part(parts, lo, hi):
            L.check(lib.f5b_dit_train_backward_part(self.handle, dpred.data_ptr(), self.cp["w1_t"].data_ptr(), self.cp["w2_t"].data_ptr(),
                                                    C.byref(self.grads), dtext.data_ptr(), B, n, None, rope.data_ptr(), ws.data_ptr(),
                                                    ws.numel(), parts, lo, hi, s), "f5b_dit_train_backward_part")

        works = []
        if overlap:
            depth = self.dit.depth
            nb = max(1, min(buckets, depth))
            bounds = [round(depth * (nb - k) / nb) for k in range(nb + 1)]  # depth ... 0
            part(1, 0, 0)
            for k in range(nb):
                hi, lo = bounds[k], bounds[k + 1]
                if hi <= lo:
                    continue
                part(2, lo, hi)
                for name, per in self._stacked:  # these blocks' gradients are final: reduce them under the remaining backward
                    o = self.offsets[name]
                    works.append(dist.all_reduce(self.g[o + lo * per:o + hi * per], op=dist.ReduceOp.SUM, group=group, async_op=True))
            part(4, 0, 0)
        else:
            part(7, 0, self.dit.depth)
        L.check(lib.f5b_dit_text_embed_backward(self.handle, text.data_ptr(), text.shape[1], B, n, int(drop_text), dtext.data_ptr(),
                                                C.byref(self.grads), tws.data_ptr(), tws.numel(), s), "f5b_dit_text_embed_backward")
        if overlap:
            self._fold_split_grads()
            for a, b in self._rest:
                works.append(dist.all_reduce(self.g[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True))
            self._pending = (works, 1.0 / dist.get_world_size(group))
        return out2[0], cond, pred

    def _fold_split_grads(self):
        """InputEmbedding.proj's gradient is produced per source (x | cond, text) in the padded kernel layout: fold it back"""
        mel = self.mel
        g = self.in_w.grad
        g[:, :mel] += self.g_in_wx[:, :mel]
        g[:, mel:2 * mel] += self.g_in_wct[:, :mel]
        g[:, 2 * mel:] += self.g_in_wct[:, 128:]
        self.g_in_wx.zero_()
        self.g_in_wct.zero_()

    # ------------------------------------------------------------------------------------------------ optimizer
    def allreduce_grads(self, group=None) -> float:
        """DDP's gradient averaging (trainer.py:1280 via accelerate) as ONE flat all-reduce; returns the scale still to apply"""
        from .parallel import allreduce_flat_
        if self._pending is not None:  # launched under the backward (loss_and_grads(overlap_allreduce=True)): just wait
            works, scale = self._pending
            self._pending = None
            for w in works:
                w.wait()
            return scale
        self._fold_split_grads()
        return allreduce_flat_(self.g, group)

    def broadcast_params(self, src: int = 0, group=None):
        from .parallel import broadcast_flat_
        broadcast_flat_(self.p, src, group)
        self.sync_from_master()

    @L.on_own_device
    @torch.no_grad()
    def step(self, lr: float | None = None, grad_scale: float = 1.0, update_ema: bool = True):
        """clip_grad_norm_ + AdamW + EMA in one fused pass that also writes the bf16 mirror (trainer.py:1280-1287, 1321)"""
        self._fold_split_grads()
        self.step_count += 1
        lr = self.lr if lr is None else lr
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        s = L.stream()
        if clip:
            L.check(self.lib.f5b_grad_sumsq(self.g.data_ptr(), self.n, self._red.data_ptr(), self._sumsq.data_ptr(), s), "f5b_grad_sumsq")
        decay = -1.0
        if self.ema is not None and update_ema:
            self.ema_calls += 1
            dcy = self.ema_schedule.decay_for_call(self.ema_calls)
            if dcy == "copy":
                decay = 0.0
            elif dcy is not None:
                decay = float(dcy)
        L.check(self.lib.f5b_adamw_ema_step(self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), L.ptr(self.ema),
                                            self.mirror.data_ptr(), self.n, lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                            self.step_count, self._sumsq.data_ptr() if clip else None, float(self.max_grad_norm or 0.0),
                                            grad_scale, decay, s), "f5b_adamw_ema_step")
        self.refresh_packed()
        self.dit.invalidate()  # the inference engine's packed copies are stale

    def grad_norm(self, grad_scale: float = 1.0) -> torch.Tensor:
        L.check(self.lib.f5b_grad_sumsq(self.g.data_ptr(), self.n, self._red.data_ptr(), self._sumsq.data_ptr(), L.stream()), "f5b_grad_sumsq")
        return self._sumsq.sqrt() * grad_scale

    # ------------------------------------------------------------------------------------------------ checkpoints
    def checkpoint(self, update: int, scheduler_state: dict | None = None) -> dict:
        """The reference's training checkpoint (Trainer.save_checkpoint, trainer.py:521-530): `model_state_dict`,
        `optimizer_state_dict` in torch.optim.AdamW's layout (parameters numbered in `model.parameters()` order, per-parameter
        `step` / `exp_avg` / `exp_avg_sq`), `ema_model_state_dict` in ema_pytorch's layout (`ema_model.` prefix + `initted`, `step`),
        `scheduler_state_dict`, `update` — so the reference's loaders and tools keep working on checkpoints written here."""
        cfm = self.cfm
        where = {id(p): (o, s) for p, o, s in self.params}
        plist = list(cfm.parameters())
        state = {}
        for i, p in enumerate(plist):
            o, n_ = where[id(p)]
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.m[o:o + n_].view_as(p).clone(),
                        "exp_avg_sq": self.v[o:o + n_].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(plist)))}
        ck = dict(model_state_dict={k: v.detach().clone() for k, v in cfm.state_dict().items()},
                  optimizer_state_dict={"state": state, "param_groups": [group]},
                  scheduler_state_dict=scheduler_state or {}, update=update)
        if self.ema is not None:
            ema = {"initted": torch.tensor(self.ema_calls > 0), "step": torch.tensor(self.ema_calls)}
            names = {id(p): k for k, p in cfm.named_parameters()}
            for k, v in cfm.state_dict().items():
                ema["ema_model." + k] = v.detach().clone()
            for p, o, n_ in self.params:
                ema["ema_model." + names[id(p)]] = self.ema[o:o + n_].view_as(p).clone()
            ck["ema_model_state_dict"] = ema
        return ck

    def save_checkpoint(self, path: str, update: int, scheduler_state: dict | None = None) -> None:
        torch.save(self.checkpoint(update, scheduler_state), path)

    @staticmethod
    def _model_state_dict(ckpt: dict):
        """trainer.py:676-728: the weights live under the first non-empty of `model_state_dict`, `ema_model_state_dict`,
        `state_dict`, `model` (or the file is a flat .safetensors dict); a prefix carried by >= 80 % of the keys (`ema_model.` for an
        EMA source, `module.`, `model.`, `_orig_mod.`) is stripped; ema_pytorch's `initted` / `step` entries are dropped."""
        raw, is_ema = None, False
        for key in ("model_state_dict", "ema_model_state_dict", "state_dict", "model", "state_dict_loaded_from_safetensors"):
            v = ckpt.get(key)
            if isinstance(v, dict) and v:
                raw, is_ema = v, key == "ema_model_state_dict"
                break
        if raw is None:
            raise KeyError(f"no model state dict in the checkpoint (top-level keys: {list(ckpt.keys())})")
        prefixes = ["module.", "model.", "_orig_mod."]
        if is_ema and any(k.startswith("ema_model.") for k in raw):
            prefixes.insert(0, "ema_model.")
        used = None
        first = next(iter(raw))
        for pre in prefixes:
            if first.startswith(pre) and sum(1 for k in raw if k.startswith(pre)) >= 0.8 * len(raw):
                used = pre
                break
        out = {}
        for k, v in raw.items():
            k2 = k[len(used):] if used and k.startswith(used) else k
            if k2 not in ("initted", "step"):
                out[k2] = v
        return out

    @torch.no_grad()
    def load_checkpoint(self, ckpt, grad_accumulation_steps: int = 1) -> int:
        """Load a checkpoint in any of the reference's formats (trainer.py:600-827).  A full training checkpoint
        (`optimizer_state_dict` + `update`) restores weights, Adam moments and EMA and returns `update + 1`, the update the
        reference resumes at (:801-812); weights-only files (pretrained_*.pt / .safetensors, EMA-only, pruned) load the weights
        (missing keys keep their initial values, like load_state_dict(strict=False)) and return 0."""
        if isinstance(ckpt, (str, os.PathLike)):
            path = str(ckpt)
            if path.endswith(".safetensors"):
                from safetensors.torch import load_file
                ckpt = {"state_dict_loaded_from_safetensors": load_file(path, device="cpu")}
            else:
                try:
                    ckpt = torch.load(path, map_location="cpu", weights_only=True)
                except Exception:  # noqa: BLE001 — reference checkpoints may pickle plain python objects (pruning_info, ...)
                    ckpt = torch.load(path, map_location="cpu", weights_only=False)
        if not isinstance(ckpt, dict):
            raise TypeError("checkpoint is not a dictionary")
        cfm = self.cfm
        sd = self._model_state_dict(ckpt)
        full = "optimizer_state_dict" in ckpt and "update" in ckpt
        names = {id(p): k for k, p in cfm.named_parameters()}
        known = set(names.values())
        if not any(k in known for k in sd):
            raise KeyError("the checkpoint's state dict shares no key with the model")
        if not full:
            for p, o, n_ in self.params:
                k = names[id(p)]
                if k in sd:
                    if sd[k].numel() != n_:
                        raise ValueError(f"checkpoint tensor {k} has shape {tuple(sd[k].shape)}, the model expects {tuple(p.shape)}")
                    self.p[o:o + n_].copy_(sd[k].reshape(-1))
            if self.ema is not None:
                self.ema.copy_(self.p)
            self.sync_from_master()
            return 0
        plist = list(cfm.parameters())
        index = {id(p): i for i, p in enumerate(plist)}
        opt = ckpt.get("optimizer_state_dict", {}).get("state", {})
        ema = ckpt.get("ema_model_state_dict")
        for p, o, n_ in self.params:
            k = names[id(p)]
            self.p[o:o + n_].copy_(sd[k].reshape(-1))
            st = opt.get(index[id(p)]) or opt.get(str(index[id(p)]))
            if st is not None:
                self.m[o:o + n_].copy_(st["exp_avg"].reshape(-1))
                self.v[o:o + n_].copy_(st["exp_avg_sq"].reshape(-1))
                self.step_count = int(float(st["step"]))
            if self.ema is not None and ema is not None and ("ema_model." + k) in ema:
                self.ema[o:o + n_].copy_(ema["ema_model." + k].reshape(-1))
        if ema is not None and "step" in ema:
            self.ema_calls = int(ema["step"])
        self.sync_from_master()
        return int(ckpt["update"]) + 1  # trainer.py:812 `start_update += 1`

    def ema_state_dict(self) -> dict:
        names = {id(p): k for k, p in self.dit.named_parameters()}
        return {"ema_model.transformer." + names[id(p)]: self.ema[o:o + s].view_as(p).clone() for p, o, s in self.params}
