"""Load the REAL reference modules (unmodified, by path)
(TEST INFRASTRUCTURE: used by oracle/gen_golden.py, by the `not gpu` test that pins the restatement, and by bench.py's CPU
arm.  /root/reference does not exist on the GPU box; there the four modules are loaded from the byte-compiled files under
`oracle/_ref/` that `python -m oracle.make_ref` built from the reference sources in the build container — git-ignored, travels
with the snapshot like a built .so.)

Recipe: inject tiny stand-ins for the third-party packages that are absent here
(torchdiffeq, x_transformers, librosa, jieba, pypinyin), register stub parent packages so
`f5_tts/model/__init__.py` (which drags in the trainer) never runs, then exec the reference's own
utils.py, modules.py, backbones/dit.py and cfm.py from where they lie.  No reference source is copied.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("F5_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REF_ROOT, "src", "f5_tts")
_PYC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "f5_tts")  # built by oracle/make_ref.py
_loaded = None


def source_available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "model", "cfm.py"))


def available() -> bool:
    """the reference's modules can be loaded: from source (build container) or from oracle/_ref/*.code (GPU box)"""
    return source_available() or os.path.isfile(os.path.join(_PYC, "model", "cfm.code"))


def _shim_modules():
    from . import f5_oracle as O

    td = types.ModuleType("torchdiffeq")

    def odeint(fn, y0, t, method="euler", **kw):
        # fixed-grid solver handing t_i to fn as a 0-dim tensor in y's dtype (torchdiffeq behaviour)
        return O.odeint_fixed(lambda ti, y: fn(ti.to(y.dtype), y), y0, t, method)

    td.odeint = odeint

    xt = types.ModuleType("x_transformers")
    xtx = types.ModuleType("x_transformers.x_transformers")

    class RotaryEmbedding(torch.nn.Module):
        def __init__(self, dim, **kw):
            super().__init__()
            self.dim = dim
            self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))

        def forward_from_seq_len(self, seq_len):
            return O.rotary_freqs(seq_len, self.dim).unsqueeze(0), 1.0

    def apply_rotary_pos_emb(t, freqs, scale=1):
        return O.apply_rotary(t, freqs[:, -t.shape[-2]:, :])

    xtx.RotaryEmbedding = RotaryEmbedding
    xtx.apply_rotary_pos_emb = apply_rotary_pos_emb
    xt.x_transformers = xtx

    lib = types.ModuleType("librosa")
    libf = types.ModuleType("librosa.filters")

    def _no_mel(*a, **k):
        raise RuntimeError("librosa mel (bigvgan path) is not part of the hot path")

    libf.mel = _no_mel
    lib.filters = libf
    jieba = types.ModuleType("jieba")
    pyp = types.ModuleType("pypinyin")
    pyp.lazy_pinyin = lambda *a, **k: []
    pyp.Style = types.SimpleNamespace(TONE3=0)
    return {"torchdiffeq": td, "x_transformers": xt, "x_transformers.x_transformers": xtx,
            "librosa": lib, "librosa.filters": libf, "jieba": jieba, "pypinyin": pyp}


def load():
    """Returns a namespace with the reference's own `utils`, `modules`, `dit`, `cfm` modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT} and no byte-compiled modules under {_PYC}")
    from_source = source_available()
    sys.dont_write_bytecode = True
    for k, m in _shim_modules().items():
        sys.modules.setdefault(k, m)
    for name, sub in (("f5_tts", ""), ("f5_tts.model", "model"), ("f5_tts.model.backbones", "model/backbones")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(_SRC if from_source else _PYC, sub)]
            sys.modules[name] = pkg

    def _exec(name, rel):
        if from_source:
            spec = importlib.util.spec_from_file_location(name, os.path.join(_SRC, rel))
        else:
            path = os.path.join(_PYC, rel[:-3] + ".code")  # a .pyc under a neutral extension (oracle/make_ref.py)
            spec = importlib.util.spec_from_file_location(name, path, loader=importlib.machinery.SourcelessFileLoader(name, path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    utils = _exec("f5_tts.model.utils", "model/utils.py")
    modules = _exec("f5_tts.model.modules", "model/modules.py")
    dit = _exec("f5_tts.model.backbones.dit", "model/backbones/dit.py")
    cfm = _exec("f5_tts.model.cfm", "model/cfm.py")

    # documented oracle adjustment #1: SDPA dropout_p forced to 0.0 (reference modules.py:490 passes 0.1)
    real_sdpa = torch.nn.functional.scaled_dot_product_attention

    def sdpa_no_dropout(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, **kw):
        return real_sdpa(q, k, v, attn_mask=attn_mask, dropout_p=0.0, is_causal=is_causal, **kw)

    modules.F = types.SimpleNamespace(**{n: getattr(torch.nn.functional, n) for n in dir(torch.nn.functional)})
    modules.F.scaled_dot_product_attention = sdpa_no_dropout
    _loaded = types.SimpleNamespace(utils=utils, modules=modules, dit=dit, cfm=cfm)
    return _loaded


def build_reference_cfm(cfg, sd, method="euler"):
    """Instantiate the reference's CFM(DiT) and load `sd` (oracle.weights.make_dit_state_dict) into it."""
    ref = load()
    tr = ref.dit.DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head, ff_mult=cfg.ff_mult,
                     mel_dim=cfg.mel_dim, text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim,
                     text_mask_padding=cfg.text_mask_padding, conv_layers=cfg.conv_layers,
                     pe_attn_head=cfg.pe_attn_head)
    model = ref.cfm.CFM(transformer=tr, odeint_kwargs=dict(method=method),
                        mel_spec_kwargs=dict(n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=cfg.mel_dim,
                                             target_sample_rate=24000, mel_spec_type="vocos")).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    bad = [k for k in missing if "inv_freq" not in k]
    assert not bad and not unexpected, (bad, unexpected)
    return model
