"""Vocos-compatible vocoder (`charactr/vocos-mel-24khz` architecture) with the interface the reference uses:
`vocoder.decode(mel [b, 100, T]) -> wav [b, 256 (T-1)]` (call sites /root/reference/src/f5_tts/infer/f5tts_wrapper.py:524,
infer/utils_infer.py:488) and the upstream `vocos` package's state_dict keys, so `pytorch_model.bin` loads
key-for-key.  ConvNeXt backbone GEMMs run on the tcgen05 engine, depth-wise conv + LayerNorm and the iSTFT head (exp / cos /
sin / irFFT / overlap-add) are fused memory-bound kernels — all inside libf5b200.so."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L

bf16, f32 = torch.bfloat16, torch.float32


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the math runs in libf5b200.so via Vocos.decode")


class _ConvNeXtBlock(_Holder):
    def __init__(self, dim, intermediate_dim, layer_scale_init_value):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(intermediate_dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim))


class _Backbone(_Holder):
    def __init__(self, input_channels, dim, intermediate_dim, num_layers):
        super().__init__()
        self.embed = nn.Conv1d(input_channels, dim, kernel_size=7, padding=3)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.convnext = nn.ModuleList([_ConvNeXtBlock(dim, intermediate_dim, 1.0 / num_layers) for _ in range(num_layers)])
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-6)


class _ISTFT(_Holder):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))


class _Head(_Holder):
    def __init__(self, dim, n_fft):
        super().__init__()
        self.out = nn.Linear(dim, n_fft + 2)
        self.istft = _ISTFT(n_fft)


class Vocos(nn.Module):
    def __init__(self, n_mels=100, dim=512, intermediate_dim=1536, num_layers=8, n_fft=1024, hop_length=256):
        super().__init__()
        self.n_mels, self.dim, self.intermediate_dim, self.num_layers = n_mels, dim, intermediate_dim, num_layers
        self.n_fft, self.hop_length = n_fft, hop_length
        self.backbone = _Backbone(n_mels, dim, intermediate_dim, num_layers)
        self.head = _Head(dim, n_fft)
        self._engine = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    @classmethod
    def from_hparams(cls, config_path=None):
        """vocos.Vocos.from_hparams(config.yaml) (utils_infer.py:113-118): the reference only ever loads vocos-mel-24khz."""
        return cls()

    def invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        # upstream checkpoints also carry feature_extractor.* buffers (the encoder side, unused by decode)
        sd = {k: v for k, v in state_dict.items() if not k.startswith("feature_extractor.")}
        return super().load_state_dict(sd, strict=strict, **kw)

    def _build(self, device):
        lib = L.load()
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        keep = []

        def dev(t, dtype):
            t = t.to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t

        Lr, D, I = self.num_layers, self.dim, self.intermediate_dim
        d = L.VocosDesc()
        d.n_mels, d.dim, d.intermediate, d.num_layers, d.n_fft, d.hop = self.n_mels, D, I, Lr, self.n_fft, self.hop_length
        ld = (7 * self.n_mels + 7) // 8 * 8
        ew = sd["backbone.embed.weight"]  # [D, n_mels, 7] -> im2col column k*n_mels + c
        ewp = torch.zeros(D, ld, dtype=ew.dtype, device=ew.device)
        ewp[:, :7 * self.n_mels] = ew.permute(0, 2, 1).reshape(D, 7 * self.n_mels)
        P = dict(embed_w=dev(ewp, bf16), embed_b=dev(sd["backbone.embed.bias"], f32), norm_w=dev(sd["backbone.norm.weight"], f32),
                 norm_b=dev(sd["backbone.norm.bias"], f32))
        cn = "backbone.convnext.{}."

        def stack(name, dtype, reshape=None):
            ts = [sd[cn.format(i) + name] for i in range(Lr)]
            if reshape:
                ts = [t.reshape(reshape) for t in ts]
            return dev(torch.stack(ts), dtype)

        P.update(dw_w=stack("dwconv.weight", f32, (D, 7)), dw_b=stack("dwconv.bias", f32), ln_w=stack("norm.weight", f32),
                 ln_b=stack("norm.bias", f32), pw1_w=stack("pwconv1.weight", bf16), pw1_b=stack("pwconv1.bias", f32),
                 pw2_w=stack("pwconv2.weight", bf16), pw2_b=stack("pwconv2.bias", f32), gamma=stack("gamma", f32),
                 fln_w=dev(sd["backbone.final_layer_norm.weight"], f32), fln_b=dev(sd["backbone.final_layer_norm.bias"], f32),
                 head_w=dev(sd["head.out.weight"], bf16), head_b=dev(sd["head.out.bias"], f32))
        d.ld_embed = ld
        for k, v in P.items():
            setattr(d, k, v.data_ptr())
        h = L.vp()
        L.check(lib.f5b_vocos_create(C.byref(d), C.byref(h)), "f5b_vocos_create")
        self._engine = dict(handle=h, keep=keep, desc=d, device=torch.device(device), lib=lib, ws=None)

    @L.on_own_device
    @torch.no_grad()
    def decode(self, features_input: torch.Tensor, **kwargs) -> torch.Tensor:
        """mel [b, n_mels, T] -> wav [b, hop (T-1)]"""
        dev_ = self.head.out.weight.device
        if dev_.type != "cuda":
            raise L.F5bError("Vocos.decode needs the module on a CUDA device (no CPU fallback)")
        if self._engine is None or self._engine["device"] != dev_:
            self._build(dev_)
        e = self._engine
        B, _, T = features_input.shape
        mel = features_input.to(device=dev_, dtype=f32).permute(0, 2, 1).contiguous()  # token-major [b, T, n_mels]
        nbytes = e["lib"].f5b_vocos_workspace_bytes(e["handle"], B, T)
        if e["ws"] is None or e["ws"].numel() < nbytes:
            e["ws"] = torch.empty(nbytes, dtype=torch.uint8, device=dev_)
        wav = torch.empty(B, self.hop_length * (T - 1), dtype=f32, device=dev_)
        L.check(e["lib"].f5b_vocos_decode(e["handle"], mel.data_ptr(), B, T, wav.data_ptr(), e["ws"].data_ptr(), e["ws"].numel(),
                                          L.stream()), "f5b_vocos_decode")
        return wav

    def forward(self, features_input, **kwargs):
        return self.decode(features_input, **kwargs)

    def __del__(self):
        try:
            if self._engine is not None:
                self._engine["lib"].f5b_vocos_destroy(self._engine["handle"])
        except Exception:  # noqa: BLE001
            pass
