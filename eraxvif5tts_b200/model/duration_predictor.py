"""DurationPredictor with the reference's constructor, parameter names (state_dict keys) and forward signatures
(model/duration_predictor.py:4-66).  The modules only hold the parameters; the eval-mode forward is two kernels in
libf5b200.so (csrc/align.cu: f5b_duration_predictor).  There is no autograd graph: `loss_and_grads` is the train-mode forward
(dropout), the duration loss of train/distil_reload.py:1096-1124 and the hand-written backward, filling every parameter's .grad."""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib as L

f32 = torch.float32


class DurationPredictor(nn.Module):
    def __init__(self, text_num_embeds, in_channels, filter_channels, kernel_size, p_dropout, gin_channels=0):
        super().__init__()
        if kernel_size % 2 != 1:
            raise NotImplementedError("odd kernel sizes only (padding = kernel_size // 2 keeps the length)")
        self.text_embed = nn.Embedding(text_num_embeds + 1, in_channels)  # 0 is the filler token
        self.in_channels, self.filter_channels, self.kernel_size = in_channels, filter_channels, kernel_size
        self.p_dropout, self.gin_channels = p_dropout, gin_channels
        self.drop = nn.Dropout(p_dropout)
        self.conv_1 = nn.Conv1d(in_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_1 = nn.GroupNorm(1, filter_channels)
        self.conv_2 = nn.Conv1d(filter_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_2 = nn.GroupNorm(1, filter_channels)
        self.proj = nn.Conv1d(filter_channels, 1, 1)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, in_channels, 1)

    def _run(self, ids, mask, id_shift, g):
        if g is not None:
            raise NotImplementedError("speaker conditioning g (gin_channels) is unused by the reference's call sites and is not built")
        dev = self.proj.weight.device
        if dev.type != "cuda":
            raise L.F5bError("DurationPredictor needs its parameters on a CUDA device (B200); there is no CPU fallback")
        if self.training and self.p_dropout > 0:
            raise NotImplementedError("forward() is the eval-mode forward: call .eval(), or loss_and_grads() for a training step")
        ids = ids.to(device=dev, dtype=torch.int64).contiguous()
        mask = mask.to(device=dev, dtype=f32).contiguous()
        b, nt = ids.shape
        lo, hi = int(ids.min()) + id_shift, int(ids.max()) + id_shift
        if lo < 0 or hi >= self.text_embed.weight.shape[0]:
            raise IndexError("index out of range in self")  # nn.Embedding's error
        p = [t.detach().to(f32).contiguous() for t in (self.text_embed.weight, self.conv_1.weight, self.conv_1.bias, self.norm_1.weight,
                                                       self.norm_1.bias, self.conv_2.weight, self.conv_2.bias, self.norm_2.weight,
                                                       self.norm_2.bias, self.proj.weight, self.proj.bias)]
        h1 = torch.empty(b, nt, self.filter_channels, dtype=f32, device=dev)
        out = torch.empty(b, 1, nt, dtype=f32, device=dev)
        L.check(L.load().f5b_duration_predictor(ids.data_ptr(), id_shift, mask.data_ptr(), p[0].data_ptr(), p[0].shape[0],
                                                *[t.data_ptr() for t in p[1:]], h1.data_ptr(), out.data_ptr(), b, nt, self.in_channels,
                                                self.filter_channels, self.kernel_size, L.stream()), "f5b_duration_predictor")
        return out

    @L.on_own_device
    @torch.no_grad()
    def forward(self, x, x_mask, g=None):
        """x int [b, nt] text tokens padded with -1 (list_str_to_idx), x_mask [b, nt] -> log-durations [b, 1, nt]"""
        return self._run(x, x_mask, 1, g)

    @torch.no_grad()
    def phoneme_forward(self, phoneme_indices, phoneme_mask, g=None):
        """same network on phoneme indices (no +1 shift), duration_predictor.py:46-66"""
        return self._run(phoneme_indices, phoneme_mask, 0, g)

    # ------------------------------------------------------------------------------------------------ training
    def _param_struct(self, grads: bool):
        names = (self.text_embed.weight, self.conv_1.weight, self.conv_1.bias, self.norm_1.weight, self.norm_1.bias, self.conv_2.weight,
                 self.conv_2.bias, self.norm_2.weight, self.norm_2.bias, self.proj.weight, self.proj.bias)
        st = L.DurPredParams()
        for (field, _), p in zip(L.DurPredParams._fields_, names):
            if p.dtype != f32 or not p.is_contiguous():
                raise L.F5bError("DurationPredictor training needs contiguous fp32 parameters")
            if grads:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                setattr(st, field, p.grad.data_ptr())
            else:
                setattr(st, field, p.data_ptr())
        return st

    @L.on_own_device
    @torch.no_grad()
    def loss_and_grads(self, x, x_mask, attn=None, target_logw=None, seed=None, phoneme: bool = False, per_item: bool = False,
                       weight: float = 1.0):
        """One training step's loss and gradients (train/distil_reload.py:1096-1124): logw = self(x, x_mask) in the module's current
        mode (train: Dropout(p_dropout) after each GroupNorm), logw_ = log(attn.sum(2) + 1e-6) * mask (attn: hard alignment
        [b, nt, mel_len]; or pass target_logw [b, nt] directly), loss = sum((logw - logw_)^2) / sum(mask).  ACCUMULATES d loss / d theta
        into every parameter's .grad (created if missing) and returns (loss, logw [b, 1, nt]); step them with any optimizer.
        Quirk kept by default: the script subtracts logw_ [b, nt] from logw [b, 1, nt], which BROADCASTS to [b, b, nt] -- for b > 1
        every item's prediction is also compared with every other item's target (:1111).  per_item=True computes the evidently
        intended loss (item i against its own target only); both agree for b = 1.
        weight scales the accumulated gradients (duration_loss_weight of :1118-1119: total = ... + weight * dur_loss); the returned
        loss is unscaled."""
        dev = self.proj.weight.device
        if dev.type != "cuda":
            raise L.F5bError("DurationPredictor needs its parameters on a CUDA device (B200); there is no CPU fallback")
        id_shift = 0 if phoneme else 1
        ids = x.to(device=dev, dtype=torch.int64).contiguous()
        mask = x_mask.to(device=dev, dtype=f32).contiguous()
        b, nt = ids.shape
        lo, hi = int(ids.min()) + id_shift, int(ids.max()) + id_shift
        if lo < 0 or hi >= self.text_embed.weight.shape[0]:
            raise IndexError("index out of range in self")
        if (attn is None) == (target_logw is None):
            raise ValueError("pass exactly one of attn / target_logw")
        if attn is not None:
            target_logw = torch.log(attn.to(device=dev, dtype=f32).sum(dim=2) + 1e-6)
        target = (target_logw.to(device=dev, dtype=f32).reshape(b, nt) * mask).contiguous()
        p_drop = float(self.p_dropout) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        F_ = self.filter_channels
        h1, a, c, dpre = (torch.empty(b, nt, F_, dtype=f32, device=dev) for _ in range(4))
        stats = torch.empty(b, 4, dtype=f32, device=dev)
        out = torch.empty(b, 1, nt, dtype=f32, device=dev)
        lib, s = L.load(), L.stream()
        ps = self._param_struct(False)
        L.check(lib.f5b_duration_predictor_train_forward(ids.data_ptr(), id_shift, mask.data_ptr(), C.byref(ps), p_drop, seed, h1.data_ptr(),
                                                         a.data_ptr(), c.data_ptr(), stats.data_ptr(), out.data_ptr(), b, nt,
                                                         self.in_channels, F_, self.kernel_size, s), "f5b_duration_predictor_train_forward")
        denom = mask.sum()
        logw = out.view(b, nt)
        if per_item:
            diff = logw - target
            loss = (diff * diff).sum() / denom
            dlogw = (2.0 * weight * diff / denom).contiguous()
        else:  # [b, 1, nt] - [b, nt] -> [b, b, nt]
            diff = logw[:, None, :] - target[None, :, :]
            loss = (diff * diff).sum() / denom
            dlogw = (2.0 * weight * diff.sum(dim=1) / denom).contiguous()
        gs = self._param_struct(True)
        L.check(lib.f5b_duration_predictor_backward(dlogw.data_ptr(), ids.data_ptr(), id_shift, mask.data_ptr(), C.byref(ps), C.byref(gs),
                                                    p_drop, seed, h1.data_ptr(), a.data_ptr(), c.data_ptr(), stats.data_ptr(),
                                                    dpre.data_ptr(), b, nt, self.in_channels, F_, self.kernel_size, s),
                "f5b_duration_predictor_backward")
        return loss, out
