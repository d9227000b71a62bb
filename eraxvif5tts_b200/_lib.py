"""ctypes binding of libf5b200.so (include/f5b200.h).  The product path has NO fallback: if the library is missing or a
call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("F5B_LIB") or os.path.join(_PKG, "lib", "libf5b200.so")  # F5B_LIB: kernel-variant experiments

vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

EPI_BF16, EPI_F32, EPI_QKV_ROPE, EPI_GATE_RESID, EPI_BF16_DUAL = 0, 1, 2, 3, 4
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SILU = 0, 1, 2, 3


class F5bError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32), ("epi", i32), ("act", i32), ("bias", vp), ("out", vp), ("ldc", i32),
                ("out2", vp), ("ldc2", i32), ("out3", vp), ("addsrc", vp), ("ld_add", i32), ("rows_per_batch", i32),
                ("gate", vp), ("gate_bstride", i64), ("lens", vp), ("batch_mod", i32), ("rope", vp), ("rope_heads", i32),
                ("heads", i32), ("tf32", i32)]


class DitDesc(C.Structure):
    _fields_ = [(k, i32) for k in ("dim", "depth", "heads", "dim_head", "ff_mult", "mel_dim", "text_dim", "conv_layers",
                                   "rope_heads", "text_mask_padding", "convpos_kernel", "convpos_groups", "vocab_rows",
                                   "precision")] + \
               [(k, vp) for k in ("time_w0", "time_b0", "time_w2", "time_b2", "mod_w", "mod_b", "text_table", "text_pos",
                                  "tb_dw_w", "tb_dw_b", "tb_ln_w", "tb_ln_b", "tb_pw1_w", "tb_pw1_b", "tb_grn_g", "tb_grn_b",
                                  "tb_pw2_w", "tb_pw2_b", "in_wx", "in_wct", "in_b", "cp_w1", "cp_b1", "cp_w2", "cp_b2",
                                  "qkv_w", "qkv_b", "out_w", "out_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b", "proj_w", "proj_b")]


class DitGrads(C.Structure):
    """F5bDitGrads: f32 gradient buffers (NULL = skip)"""
    _fields_ = [(k, vp) for k in ("time_w0", "time_b0", "time_w2", "time_b2", "mod_w", "mod_b", "text_table", "tb_dw_w", "tb_dw_b",
                                  "tb_ln_w", "tb_ln_b", "tb_pw1_w", "tb_pw1_b", "tb_grn_g", "tb_grn_b", "tb_pw2_w", "tb_pw2_b",
                                  "in_wx", "in_wct", "in_b", "cp_w1", "cp_b1", "cp_w2", "cp_b2", "qkv_w", "qkv_b", "out_w", "out_b",
                                  "ff1_w", "ff1_b", "ff2_w", "ff2_b", "proj_w", "proj_b")]


class DurPredParams(C.Structure):
    """F5bDurPredParams: fp32 parameter (or gradient) pointers of the DurationPredictor"""
    _fields_ = [(k, vp) for k in ("table", "conv1_w", "conv1_b", "norm1_w", "norm1_b", "conv2_w", "conv2_b", "norm2_w", "norm2_b",
                                  "proj_w", "proj_b")]


class VocosDesc(C.Structure):
    _fields_ = [(k, i32) for k in ("n_mels", "dim", "intermediate", "num_layers", "n_fft", "hop")] + \
               [("embed_w", vp), ("embed_b", vp), ("ld_embed", i32)] + \
               [(k, vp) for k in ("norm_w", "norm_b", "dw_w", "dw_b", "ln_w", "ln_b", "pw1_w", "pw1_b", "pw2_w", "pw2_b", "gamma",
                                  "fln_w", "fln_b", "head_w", "head_b")]


# name -> (restype, argtypes); every symbol declared in include/f5b200.h
SIGNATURES = {
    "f5b_last_error": (C.c_char_p, []),
    "f5b_abi_version": (C.c_int, []),
    "f5b_gemm": (C.c_int, [vp, C.c_int, vp, C.c_int, C.POINTER(GemmArgs), vp]),
    "f5b_gemm_tn": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_ln_modulate": (C.c_int, [vp, vp, vp, i64, C.c_int, vp, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_ln_affine": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, f32, vp]),
    "f5b_attn_fwd": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_attn_fwd_lse": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_attn_bwd": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, f32, vp,
                               C.c_int, vp]),
    "f5b_gate_add": (C.c_int, [vp, vp, vp, i64, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_gate_add_ln_modulate": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_gate_bwd": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_act_fwd": (C.c_int, [vp, vp, i64, C.c_int, vp]),
    "f5b_act_bwd": (C.c_int, [vp, vp, vp, vp, i64, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_ln_modulate_bwd": (C.c_int, [vp, vp, vp, i64, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_mse_grad": (C.c_int, [vp, vp, vp, vp, vp, i64, C.c_int, C.c_int, vp]),
    "f5b_dit_train_ws_bytes": (sz, [vp, C.c_int, C.c_int]),
    "f5b_dit_train_forward": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, sz, vp]),
    "f5b_dit_train_backward": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, sz, vp]),
    "f5b_set_dependent_launch": (C.c_int, [C.c_int]),
    "f5b_crossfade_append": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_int, vp]),
    "f5b_pcm16": (C.c_int, [vp, vp, C.c_int64, vp]),
    "f5b_align_viterbi": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_align_window": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_duration_predictor": (C.c_int, [vp, C.c_int] + [vp] * 2 + [C.c_int] + [vp] * 12 + [C.c_int] * 5 + [vp]),
    "f5b_duration_predictor_train_forward": (C.c_int, [vp, C.c_int, vp, vp, C.c_float, C.c_uint64, vp, vp, vp, vp, vp] + [C.c_int] * 5 + [vp]),
    "f5b_duration_predictor_backward": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_uint64, vp, vp, vp, vp, vp] + [C.c_int] * 5 + [vp]),
    "f5b_distill_loss": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp]),
    "f5b_distill_grad": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp]),
    "f5b_train_set_dropout": (C.c_int, [C.c_float, C.c_uint64]),
    "f5b_train_set_checkpoint": (C.c_int, [C.c_int]),
    "f5b_train_set_attn_dropout": (C.c_int, [C.c_float]),
    "f5b_dit_set_attn_dropout": (C.c_int, [C.c_float, C.c_uint64, vp]),
    "f5b_dit_train_backward_part": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, sz, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_dit_text_train_ws_bytes": (sz, [vp, C.c_int, C.c_int]),
    "f5b_dit_text_embed_train": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, sz, vp]),
    "f5b_dit_text_embed_backward": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, sz, vp]),
    "f5b_ln_affine_bwd": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_grn_gelu_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_dwconv7_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_text_lookup_bwd": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_convpos": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_pack_convpos_weight": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_pack_convpos_weight_t": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_convpos_packed_elems": (sz, [C.c_int, C.c_int, C.c_int]),
    "f5b_dwconv7_ln": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_grn": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_text_lookup": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_mask_rows_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
    "f5b_time_sinus": (C.c_int, [vp, vp, C.c_int, vp]),
    "f5b_silu_bf16": (C.c_int, [vp, vp, i64, vp]),
    "f5b_pack_bf16": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_cfg_euler": (C.c_int, [vp, vp, vp, f32, f32, vp, C.c_int, vp, C.c_int, C.c_int, vp]),
    "f5b_melspec": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_istft_head": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
    "f5b_im2col7": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_rope_table": (C.c_int, [vp, C.c_int, vp]),
    "f5b_dit_create": (C.c_int, [C.POINTER(DitDesc), C.POINTER(vp)]),
    "f5b_dit_destroy": (None, [vp]),
    "f5b_dit_workspace_bytes": (sz, [vp, C.c_int, C.c_int]),
    "f5b_dit_modulation": (C.c_int, [vp, vp, C.c_int, vp, vp, vp]),
    "f5b_dit_modulation_ws_bytes": (sz, [vp, C.c_int]),
    "f5b_dit_text_embed": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "f5b_dit_text_ws_bytes": (sz, [vp, C.c_int, C.c_int]),
    "f5b_dit_input_const": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "f5b_dit_forward": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, i64, vp, vp, vp, vp, sz, vp]),
    "f5b_vocos_create": (C.c_int, [C.POINTER(VocosDesc), C.POINTER(vp)]),
    "f5b_vocos_destroy": (None, [vp]),
    "f5b_vocos_workspace_bytes": (sz, [vp, C.c_int, C.c_int]),
    "f5b_cfg_euler_dev": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp]),
    "f5b_fm_prepare": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_masked_mse": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]),
    "f5b_grad_sumsq": (C.c_int, [vp, i64, vp, vp, vp]),
    "f5b_adamw_ema_step": (C.c_int, [vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, C.c_int, vp, f32, f32, f32, vp]),
    "f5b_prof_enabled": (C.c_int, []),
    "f5b_prof_add": (None, [C.POINTER(C.c_double), C.c_int]),
    "f5b_prof_reset": (None, [C.c_int]),
    "f5b_prof_read": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "f5b_vocos_decode": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, sz, vp]),
    "f5b_ln_modulate_tf32": (C.c_int, [vp, vp, vp, i64, C.c_int, vp, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_attn_fwd_tf32": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, f32, vp]),
    "f5b_attn_tf32_ws_floats": (sz, [C.c_int, C.c_int, C.c_int]),
    "f5b_convpos_tf32": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_pack_convpos_weight_tf32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "f5b_convpos_packed_elems_tf32": (sz, [C.c_int, C.c_int, C.c_int]),
    "f5b_time_sinus_tf32": (C.c_int, [vp, vp, C.c_int, vp]),
    "f5b_pack_tf32": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libf5b200.so (built in-tree by eraxvif5tts_b200.build).  Raises if it is missing — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise F5bError(f"{LIB_PATH} not found: build it with `python -m eraxvif5tts_b200.build` "
                       "(the CUDA extension is mandatory; there is no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.f5b_abi_version() != 1:
        raise F5bError("libf5b200.so ABI version mismatch; rebuild")
    if os.environ.get("F5B_ATTN_VARIANT"):  # kernel A/B switch for experiments (0 = default)
        lib.f5b_debug_attn_variant(int(os.environ["F5B_ATTN_VARIANT"]))
    if os.environ.get("F5B_ATTN_POLY"):
        lib.f5b_debug_attn_poly(int(os.environ["F5B_ATTN_POLY"]))
    if os.environ.get("F5B_PDL"):  # A/B override of set_dependent_launch() below
        lib.f5b_set_dependent_launch(int(os.environ["F5B_PDL"]))
    _lib = lib
    return lib


PDL_MAX_ROWS = 16384  # fused rows (2B x n) up to which an ODE step is launch-bound enough for dependent launches to pay


def set_dependent_launch(on: bool) -> None:
    """f5b_set_dependent_launch (include/f5b200.h) unless F5B_PDL pins it for an experiment"""
    if not os.environ.get("F5B_PDL"):
        load().f5b_set_dependent_launch(int(bool(on)))


KERNEL_KINDS = ("gemm", "attention", "convpos", "norm", "elementwise", "spectral", "vocos")


def prof_reset(enable: bool) -> None:
    load().f5b_prof_reset(int(enable))


def prof_read() -> dict:
    """{kind: dict(launches, ms, flops, bytes)} since the last prof_reset (device-synchronising)."""
    buf = (C.c_double * (4 * len(KERNEL_KINDS)))()
    check(load().f5b_prof_read(buf, len(KERNEL_KINDS)), "f5b_prof_read")
    return {k: dict(launches=int(buf[4 * i]), ms=buf[4 * i + 1], flops=buf[4 * i + 2], bytes=buf[4 * i + 3])
            for i, k in enumerate(KERNEL_KINDS)}


def prof_raw() -> list:
    """raw counters without synchronising semantics beyond prof_read's (used to diff launches around a graph capture)"""
    buf = (C.c_double * (4 * len(KERNEL_KINDS)))()
    check(load().f5b_prof_read(buf, len(KERNEL_KINDS)), "f5b_prof_read")
    return list(buf)


def prof_add(delta: list) -> None:
    buf = (C.c_double * (4 * len(KERNEL_KINDS)))(*delta)
    load().f5b_prof_add(buf, len(KERNEL_KINDS))


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().f5b_last_error()
        raise F5bError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """device pointer of a torch tensor (None -> NULL)"""
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def on_own_device(fn):
    """Method decorator for every product entry point that launches library kernels: the C library launches on the CURRENT CUDA
    device and `stream()` is the current device's stream, so an object living on cuda:N must make cuda:N current for the duration
    of the call (a model on cuda:1 used without torch.cuda.set_device(1) would otherwise launch on device 0 against device-1
    pointers).  The device is the object's `device` attribute / property, else that of its first parameter."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        import torch
        dev = getattr(self, "device", None)
        if dev is None and hasattr(self, "parameters"):
            dev = next(self.parameters()).device
        dev = torch.device(dev) if dev is not None else None
        if dev is None or dev.type != "cuda" or (dev.index is not None and dev.index == torch.cuda.current_device()):
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped
