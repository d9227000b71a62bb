"""Optimizer step of the reference's training loop (/root/reference/src/f5_tts/model/trainer.py:316-323, 1179-1188, 1280-1287, 1321)
re-designed around FLAT fp32 buffers: every parameter / gradient / Adam moment / EMA weight is a view into one contiguous
buffer, so gradient averaging is ONE NCCL all-reduce over NVLink and clip + AdamW + EMA is ONE fused HBM-bound kernel
(`f5b_adamw_ema_step`).  `FlatAdamW` is the stand-alone optimizer over any module's parameters (gradients supplied by the caller;
tests drive it against torch.optim.AdamW); the training step proper (`train.TrainEngine`) uses the same kernels over buffers laid
out in kernel order and fills the gradient buffer with the hand-written backward."""
from __future__ import annotations

import math

import torch

from . import _lib as L


class WarmupLinearDecay:
    """SequentialLR(LinearLR(1e-8 -> 1, warmup), LinearLR(1 -> 1e-8, decay)) of trainer.py:1179-1188; `lr(update)` = the rate
    used BY update number `update` (0-based).

    `steps_per_update`: the reference hands the scheduler to `accelerator.prepare`, and accelerate's AcceleratedScheduler steps
    the wrapped scheduler `num_processes` times per optimizer update (split_batches=False).  That is why the reference multiplies
    the warm-up length by num_processes (:1179-1181): in optimizer updates its warm-up lasts `num_warmup_updates` and the decay
    spans `(total - warmup * P) / P` updates.  Pass `steps_per_update = num_processes` with the reference's (multiplied) lengths to
    get exactly that: the schedule is evaluated at `update * steps_per_update`."""

    def __init__(self, base_lr: float, warmup_updates: int, total_updates: int, steps_per_update: int = 1):
        self.base_lr, self.warmup, self.decay = base_lr, max(1, warmup_updates), max(1, total_updates - warmup_updates)
        self.steps_per_update = max(1, int(steps_per_update))

    def lr(self, update: int) -> float:
        update = update * self.steps_per_update
        if update < self.warmup:
            f = 1e-8 + (1.0 - 1e-8) * update / self.warmup
        else:
            f = 1.0 + (1e-8 - 1.0) * min(update - self.warmup, self.decay) / self.decay
        return self.base_lr * f


class EmaSchedule:
    """ema_pytorch.EMA defaults as constructed at trainer.py:180 (third-party, unpinned): beta 0.9999, update_after_step 100,
    update_every 10, inv_gamma 1, power 2/3.  `decay_for_call(k)` for the k-th `update()` call (1-based) returns
    None (no-op), 'copy' (EMA := online weights) or the decay of the lerp."""

    def __init__(self, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0, power=2.0 / 3.0, min_value=0.0):
        self.beta, self.after, self.every, self.inv_gamma, self.power, self.min_value = beta, update_after_step, update_every, inv_gamma, power, min_value

    def decay_for_call(self, k: int):
        """ema_pytorch.EMA.update(): `step = self.step; self.step += 1` — the gating (`update_every`, `update_after_step`) uses the
        PRE-increment value, then update_moving_average() calls get_current_decay(), which reads self.step AFTER the increment:
        epoch = (step + 1) - update_after_step - 1 = step - update_after_step."""
        step = k - 1
        if step % self.every != 0:
            return None
        if step <= self.after:
            return "copy"
        epoch = max(step - self.after, 0)
        value = 1.0 - (1.0 + epoch / self.inv_gamma) ** (-self.power)
        return 0.0 if epoch <= 0 else min(max(value, self.min_value), self.beta)


class FlatAdamW:
    """AdamW(betas=(0.9, 0.98), eps=1e-8) + global-norm clipping + optional EMA over the parameters of `module`."""

    def __init__(self, module: torch.nn.Module, lr: float = 7.5e-5, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                 with_ema: bool = False, ema_schedule: EmaSchedule | None = None):
        self.module = module
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise L.F5bError("FlatAdamW needs the module on a CUDA device (no CPU fallback)")
        self.device, self.lib = dev, L.load()
        self.sizes = [p.numel() for p in params]
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + (s + 3) // 4 * 4)  # keep every view 16-byte aligned
        n = self.offsets[-1]
        self.n = n
        self.p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.ema = torch.zeros(n, dtype=torch.float32, device=dev) if with_ema else None
        self.params = params
        with torch.no_grad():
            for p, off, s in zip(params, self.offsets, self.sizes):
                self.p[off:off + s].copy_(p.detach().reshape(-1))
                p.data = self.p[off:off + s].view_as(p)          # parameters become views of the flat master
                p.grad = self.g[off:off + s].view_as(p)           # and so do their gradients
        if with_ema:
            self.ema.copy_(self.p)
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.ema_schedule = ema_schedule or EmaSchedule()
        self.step_count = 0
        self.ema_calls = 0
        self._ws = torch.empty(1024, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_grad(self):
        self.g.zero_()

    def allreduce_grads(self, group=None) -> float:
        """DDP's gradient averaging (trainer.py:1280 via accelerate) as one flat all-reduce; returns the scale still to be applied
        (1/world when the backend summed)."""
        from .parallel import allreduce_flat_
        return allreduce_flat_(self.g, group)

    def broadcast_params(self, src: int = 0, group=None):
        from .parallel import broadcast_flat_
        broadcast_flat_(self.p, src, group)

    def grad_norm(self, grad_scale: float = 1.0) -> torch.Tensor:
        L.check(self.lib.f5b_grad_sumsq(self.g.data_ptr(), self.n, self._ws.data_ptr(), self._sumsq.data_ptr(), L.stream()), "f5b_grad_sumsq")
        return self._sumsq.sqrt() * grad_scale

    @L.on_own_device
    def step(self, lr: float | None = None, grad_scale: float = 1.0, update_ema: bool = True):
        """one optimizer.step() (+ ema_model.update() when `with_ema`); gradients are expected in `self.g` / `p.grad`"""
        self.step_count += 1
        lr = self.lr if lr is None else lr
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            L.check(self.lib.f5b_grad_sumsq(self.g.data_ptr(), self.n, self._ws.data_ptr(), self._sumsq.data_ptr(), L.stream()), "f5b_grad_sumsq")
        decay = -1.0
        if self.ema is not None and update_ema:
            self.ema_calls += 1
            d = self.ema_schedule.decay_for_call(self.ema_calls)
            if d == "copy":
                decay = 0.0          # lerp with weight 1 = copy of the freshly updated weights
            elif d is not None:
                decay = float(d)
        L.check(self.lib.f5b_adamw_ema_step(self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                            L.ptr(self.ema), None, self.n, lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                            self.step_count, self._sumsq.data_ptr() if clip else None,
                                            float(self.max_grad_norm or 0.0), grad_scale, decay, L.stream()), "f5b_adamw_ema_step")
        inv = getattr(self.module, "invalidate", None)
        if callable(inv):
            inv()  # packed bf16 engine copies are stale
        for mod in self.module.modules():
            if mod is not self.module and callable(getattr(mod, "invalidate", None)):
                mod.invalidate()

    def ema_state_dict(self) -> dict:
        """EMA weights under the reference's checkpoint naming (`ema_model.<key>`, trainer.py:521-598)"""
        out = {}
        names = {id(p): k for k, p in self.module.named_parameters()}
        for p, off, s in zip(self.params, self.offsets, self.sizes):
            out["ema_model." + names[id(p)]] = self.ema[off:off + s].view_as(p).clone()
        return out
