"""DiT with the reference's constructor, forward signature and state_dict keys
(/root/reference/src/f5_tts/model/backbones/dit.py:103-233).  The nn.Module tree below only HOLDS parameters (so reference
checkpoints load key-for-key); the math is done by libf5b200.so on bf16 copies packed into fused layouts by `DiTEngine`."""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn

from ... import _lib as L

# fused rows (2B x n) up to which the CFG branches run as two concurrent chains (step_session).  0 = never by default: measured on
# cfg-1 (B = 1 x 940 frames) the fork costs 67.9 / 68.2 -> 72.9 / 72.7 ms per utterance and on cfg-3 688 -> 700 ms per batch (two
# 1-CTA-per-SM persistent GEMM grids do not co-reside, and the fork / join edges lose the programmatic dependent launches), so the
# fused 2B-row batch stays the product path; F5B_SPLIT_CFG=1 selects the fork (bit-identical, tests/test_gpu_edges.py).
SPLIT_CFG_MAX_ROWS = 0

bf16, f32 = torch.bfloat16, torch.float32


# ----------------------------------------------------------------------------------------------------------------------
# parameter containers (names = the reference's module tree)
# ----------------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the math runs in libf5b200.so via DiT.forward / CFM.sample")


class _TimestepEmbedding(_Holder):
    def __init__(self, dim, freq_embed_dim=256):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.Linear(freq_embed_dim, dim), nn.SiLU(), nn.Linear(dim, dim))


class _GRN(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1, 1, dim))
        self.beta = nn.Parameter(torch.zeros(1, 1, dim))


class _ConvNeXtV2Block(_Holder):
    def __init__(self, dim, intermediate_dim):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.act = nn.GELU()
        self.grn = _GRN(intermediate_dim)
        self.pwconv2 = nn.Linear(intermediate_dim, dim)


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0) -> torch.Tensor:
    """model/modules.py:196-207: cat(cos, sin) of outer(pos, theta^(-2j/dim))."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    freqs = torch.outer(torch.arange(end), freqs).float()
    return torch.cat([torch.cos(freqs), torch.sin(freqs)], dim=-1)


class _TextEmbedding(_Holder):
    def __init__(self, text_num_embeds, text_dim, mask_padding=True, conv_layers=0, conv_mult=2):
        super().__init__()
        self.text_embed = nn.Embedding(text_num_embeds + 1, text_dim)
        self.mask_padding = mask_padding
        self.extra_modeling = conv_layers > 0
        if conv_layers > 0:
            self.precompute_max_pos = 4096
            self.register_buffer("freqs_cis", precompute_freqs_cis(text_dim, self.precompute_max_pos), persistent=False)
            self.text_blocks = nn.Sequential(*[_ConvNeXtV2Block(text_dim, text_dim * conv_mult) for _ in range(conv_layers)])


class _ConvPositionEmbedding(_Holder):
    def __init__(self, dim, kernel_size=31, groups=16):
        super().__init__()
        assert kernel_size % 2 != 0
        self.conv1d = nn.Sequential(
            nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.Mish(),
            nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.Mish())


class _InputEmbedding(_Holder):
    def __init__(self, mel_dim, text_dim, out_dim):
        super().__init__()
        self.proj = nn.Linear(mel_dim * 2 + text_dim, out_dim)
        self.conv_pos_embed = _ConvPositionEmbedding(dim=out_dim)


class _RotaryEmbedding(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class _AdaLayerNorm(_Holder):
    def __init__(self, dim, mult):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, dim * mult)


class _Attention(_Holder):
    def __init__(self, dim, heads, dim_head, dropout):
        super().__init__()
        inner = heads * dim_head
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(dropout)])


class _FeedForward(_Holder):
    def __init__(self, dim, mult, dropout):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(nn.Sequential(nn.Linear(dim, inner), nn.GELU(approximate="tanh")), nn.Dropout(dropout),
                                nn.Linear(inner, dim))


class _DiTBlock(_Holder):
    def __init__(self, dim, heads, dim_head, ff_mult, dropout):
        super().__init__()
        self.attn_norm = _AdaLayerNorm(dim, 6)
        self.attn = _Attention(dim, heads, dim_head, dropout)
        self.ff_norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = _FeedForward(dim, ff_mult, dropout)


# ----------------------------------------------------------------------------------------------------------------------
# packed device weights + C handle
# ----------------------------------------------------------------------------------------------------------------------
def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest tf32 value (10 explicit mantissa bits, ties away from zero = PTX cvt.rna.tf32.f32), kept as fp32: the
    tcgen05 kind::tf32 MMA ignores the low 13 bits of each operand word, so rounding here makes that truncation exact."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class DiTEngine:
    """Fused-layout copies of a DiT's parameters on one CUDA device and the `F5bDit` handle built on them: bf16 operands
    (precision "bf16", the throughput mode) or fp32 words rounded to tf32 (precision "tf32", the fp32-tolerance mode)."""

    def __init__(self, dit: "DiT", device, precision: str = "bf16"):
        from ... import ops
        lib = L.load()
        if precision not in ("bf16", "tf32"):
            raise ValueError(f"precision must be 'bf16' or 'tf32', got {precision!r}")
        self.precision = precision
        tf32 = precision == "tf32"
        self.act_dtype = f32 if tf32 else bf16
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.F5bError("DiTEngine needs a CUDA device (B200); there is no CPU fallback")
        sd = {k: v.detach() for k, v in dit.state_dict().items()}
        D, depth, H = dit.dim, dit.depth, dit.heads
        mel, T, Lc = dit.mel_dim, dit.text_dim, dit.conv_layers
        self.keep = []  # tensors the handle points into

        def dev(t, dtype):
            if dtype is bf16 and tf32:  # tensor-core operand: fp32 words rounded to tf32 instead of bf16
                t = round_tf32(t.to(device=self.device, dtype=f32))
            else:
                t = t.to(device=self.device, dtype=dtype).contiguous()
            self.keep.append(t)
            return t

        def stack(fmt, dtype, n=depth, reshape=None):
            ts = [sd[fmt.format(i)] for i in range(n)]
            if reshape is not None:
                ts = [t.reshape(reshape) for t in ts]
            return dev(torch.stack(ts), dtype)

        d = L.DitDesc()
        d.dim, d.depth, d.heads, d.dim_head, d.ff_mult = D, depth, H, dit.dim_head, dit.ff_mult
        d.mel_dim, d.text_dim, d.conv_layers = mel, T, Lc
        d.rope_heads = H if dit.pe_attn_head is None else int(dit.pe_attn_head)
        d.text_mask_padding = int(bool(dit.text_mask_padding))
        d.convpos_kernel, d.convpos_groups = 31, 16
        d.vocab_rows = sd["text_embed.text_embed.weight"].shape[0]
        d.precision = 1 if tf32 else 0
        P = {}
        P["time_w0"], P["time_b0"] = dev(sd["time_embed.time_mlp.0.weight"], bf16), dev(sd["time_embed.time_mlp.0.bias"], f32)
        P["time_w2"], P["time_b2"] = dev(sd["time_embed.time_mlp.2.weight"], bf16), dev(sd["time_embed.time_mlp.2.bias"], f32)
        blk = "transformer_blocks.{}."
        P["mod_w"] = dev(torch.cat([sd[blk.format(i) + "attn_norm.linear.weight"] for i in range(depth)] + [sd["norm_out.linear.weight"]]), bf16)
        P["mod_b"] = dev(torch.cat([sd[blk.format(i) + "attn_norm.linear.bias"] for i in range(depth)] + [sd["norm_out.linear.bias"]]), f32)
        P["text_table"] = dev(sd["text_embed.text_embed.weight"], f32)
        if Lc > 0:
            P["text_pos"] = dev(dit.text_embed.freqs_cis, f32)
            tb = "text_embed.text_blocks.{}."
            P["tb_dw_w"], P["tb_dw_b"] = stack(tb + "dwconv.weight", f32, Lc, (T, 7)), stack(tb + "dwconv.bias", f32, Lc)
            P["tb_ln_w"], P["tb_ln_b"] = stack(tb + "norm.weight", f32, Lc), stack(tb + "norm.bias", f32, Lc)
            P["tb_pw1_w"], P["tb_pw1_b"] = stack(tb + "pwconv1.weight", bf16, Lc), stack(tb + "pwconv1.bias", f32, Lc)
            P["tb_grn_g"], P["tb_grn_b"] = stack(tb + "grn.gamma", f32, Lc, (2 * T,)), stack(tb + "grn.beta", f32, Lc, (2 * T,))
            P["tb_pw2_w"], P["tb_pw2_b"] = stack(tb + "pwconv2.weight", bf16, Lc), stack(tb + "pwconv2.bias", f32, Lc)
        else:
            P["text_pos"] = dev(torch.zeros(1, T), f32)
        pw = sd["input_embed.proj.weight"]  # [D, 2*mel + T] columns: x | cond | text (dit.py:95)
        wx = torch.zeros(D, 128, dtype=pw.dtype, device=pw.device)
        wx[:, :mel] = pw[:, :mel]
        wct = torch.zeros(D, 128 + T, dtype=pw.dtype, device=pw.device)
        wct[:, :mel] = pw[:, mel:2 * mel]
        wct[:, 128:] = pw[:, 2 * mel:]
        P["in_wx"], P["in_wct"], P["in_b"] = dev(wx, bf16), dev(wct, bf16), dev(sd["input_embed.proj.bias"], f32)
        cw = "input_embed.conv_pos_embed.conv1d.{}."
        with torch.cuda.device(self.device):
            for j, c in ((1, 0), (2, 2)):
                w32 = dev(sd[cw.format(c) + "weight"], f32)
                pk = ops.pack_convpos_weight_tf32(w32, 16) if tf32 else ops.pack_convpos_weight(w32, 16)
                self.keep.append(pk)
                P[f"cp_w{j}"], P[f"cp_b{j}"] = pk, dev(sd[cw.format(c) + "bias"], f32)
            torch.cuda.synchronize()
        P["qkv_w"] = dev(torch.stack([torch.cat([sd[blk.format(i) + f"attn.to_{n}.weight"] for n in "qkv"]) for i in range(depth)]), bf16)
        P["qkv_b"] = dev(torch.stack([torch.cat([sd[blk.format(i) + f"attn.to_{n}.bias"] for n in "qkv"]) for i in range(depth)]), f32)
        P["out_w"], P["out_b"] = stack(blk + "attn.to_out.0.weight", bf16), stack(blk + "attn.to_out.0.bias", f32)
        P["ff1_w"], P["ff1_b"] = stack(blk + "ff.ff.0.0.weight", bf16), stack(blk + "ff.ff.0.0.bias", f32)
        P["ff2_w"], P["ff2_b"] = stack(blk + "ff.ff.2.weight", bf16), stack(blk + "ff.ff.2.bias", f32)
        P["proj_w"], P["proj_b"] = dev(sd["proj_out.weight"], bf16), dev(sd["proj_out.bias"], f32)
        for k, v in P.items():
            setattr(d, k, v.data_ptr())
        self.P = P
        self.desc = d
        h = L.vp()
        L.check(lib.f5b_dit_create(C.byref(d), C.byref(h)), "f5b_dit_create")
        self.handle = h
        self.lib = lib
        self.dim, self.depth, self.mel_dim, self.text_dim = D, depth, mel, T
        self.mod_dim = depth * 6 * D + 2 * D
        self._ws = None
        self._rope = {}
        self._sessions = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.f5b_dit_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass

    # ---- buffers ----
    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def rope_table(self, n: int) -> torch.Tensor:
        """x_transformers RotaryEmbedding.forward_from_seq_len(n) (call site dit.py:215) as a (cos, sin) table [n, 32, 2];
        evaluated with the same torch expressions as the reference so the angles match bit for bit."""
        if n not in self._rope:
            inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=self.device).float() / 64))
            fr = torch.outer(torch.arange(n, device=self.device).float(), inv)
            if len(self._rope) >= 64:
                self._rope.pop(next(iter(self._rope)))  # captured graphs keep their own reference (step_session)
            self._rope[n] = torch.stack((fr.cos(), fr.sin()), dim=-1).contiguous()
        return self._rope[n]

    # ---- CUDA-graph step sessions (launch-bound regimes: small batches, e.g. the reference's serial B=1 chunks) ----
    def step_session(self, Bx: int, Bf: int, n: int, masked: bool, attn_p: float = 0.0):
        """Persistent buffers + one captured CUDA graph of {DiT.forward over the fused batch, CFG + Euler update} for a given
        shape.  Per ODE step only `stepbuf` (that step's modulation row followed by (cfg, dt)) changes, so the same graph is
        replayed for all steps and all later sample() calls of this shape."""
        # CFG branches as two CONCURRENT chains (opt-in experiment, see SPLIT_CFG_MAX_ROWS): the conditioned and the unconditioned
        # forward share no data but the read-only state, so the captured step can fork into two B-row forwards on two streams
        # instead of one fused 2B-row batch (bit-identical results; measured slower)
        env = os.environ.get("F5B_SPLIT_CFG", "")
        split = Bf == 2 * Bx and env != "0" and (env == "1" or Bf * n <= SPLIT_CFG_MAX_ROWS) and not attn_p > 0
        key = (Bx, Bf, n, masked, float(attn_p), split)
        sess = self._sessions.get(key)
        if sess is not None:
            self._sessions[key] = self._sessions.pop(key)  # LRU order
            return sess
        from ... import ops
        while len(self._sessions) >= 4:
            self._sessions.pop(next(iter(self._sessions)))
        dev, mel = self.device, self.mel_dim
        sess = dict(
            y=torch.zeros(Bx, n, mel, dtype=f32, device=dev), yb=torch.zeros(Bx * n, 128, dtype=self.act_dtype, device=dev),
            c0=torch.zeros(Bf, n, self.dim, dtype=f32, device=dev), pred=torch.zeros(Bf, n, mel, dtype=f32, device=dev),
            stepbuf=torch.zeros(self.mod_dim + 3, dtype=f32, device=dev),  # modulation row | cfg | dt | attention-dropout word
            lens=torch.full((Bx,), n, dtype=torch.int32, device=dev) if masked else None, graph=None, delta=None,
            rope=self.rope_table(n), ws=None, split=split)  # the graph bakes these pointers in: the session keeps them alive
        if split:
            nb = self.lib.f5b_dit_workspace_bytes(self.handle, Bx, n)
            sess["ws"] = torch.empty(nb, dtype=torch.uint8, device=dev)
            sess["ws2"] = torch.empty(nb, dtype=torch.uint8, device=dev)
            sess["fork"] = torch.cuda.Stream(device=dev)
        else:
            sess["ws"] = self.workspace(self.lib.f5b_dit_workspace_bytes(self.handle, Bf, n))
        pc, pu = sess["pred"][:Bx], (sess["pred"][Bx:] if Bf > Bx else None)
        params = sess["stepbuf"][self.mod_dim:]
        # SDPA dropout at inference (DiT.set_attn_dropout): p is baked into the graph, the mask stream follows the device word
        adrop = (float(attn_p), 0, sess["stepbuf"][self.mod_dim + 2:]) if attn_p > 0 else None

        def body():
            if split:
                cur, fork = torch.cuda.current_stream(dev), sess["fork"]
                fork.wait_stream(cur)
                self.forward(sess["yb"], Bx, sess["c0"][:Bx], Bx, n, sess["stepbuf"], 0, sess["lens"], pc, ws=sess["ws"])
                with torch.cuda.stream(fork):
                    self.forward(sess["yb"], Bx, sess["c0"][Bx:], Bx, n, sess["stepbuf"], 0, sess["lens"], pu, ws=sess["ws2"])
                cur.wait_stream(fork)
            else:
                self.forward(sess["yb"], Bx, sess["c0"], Bf, n, sess["stepbuf"], 0, sess["lens"], sess["pred"], attn_drop=adrop)
            if self.precision == "tf32":
                L.check(self.lib.f5b_cfg_euler_dev(sess["y"].data_ptr(), pc.data_ptr(), L.ptr(pu), params.data_ptr(), None, 128, None,
                                                   Bx * n, mel, L.stream()), "f5b_cfg_euler_dev")
                self.pack_state(sess["y"], sess["yb"])
            else:
                L.check(self.lib.f5b_cfg_euler_dev(sess["y"].data_ptr(), pc.data_ptr(), L.ptr(pu), params.data_ptr(),
                                                   sess["yb"].data_ptr(), 128, None, Bx * n, mel, L.stream()), "f5b_cfg_euler_dev")

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()  # warm-up outside capture (first-use attribute setup, tensor-map cache)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        before = L.prof_raw()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        after = L.prof_raw()
        sess["graph"] = g
        sess["delta"] = [a - b for a, b in zip(after, before)]
        L.prof_add([-d for d in sess["delta"]])  # the capture itself launched nothing
        self._sessions[key] = sess
        return sess

    # ---- pieces of DiT.forward ----
    def modulation(self, t: torch.Tensor) -> torch.Tensor:
        """t f32 [M] -> f32 [M, depth*6D + 2D] (time MLP + every AdaLN linear)."""
        t = t.to(device=self.device, dtype=f32).contiguous()
        M = t.numel()
        mod = torch.empty(M, self.mod_dim, dtype=f32, device=self.device)
        ws = torch.empty(self.lib.f5b_dit_modulation_ws_bytes(self.handle, M), dtype=torch.uint8, device=self.device)
        L.check(self.lib.f5b_dit_modulation(self.handle, t.data_ptr(), M, mod.data_ptr(), ws.data_ptr(), L.stream()), "f5b_dit_modulation")
        return mod

    def text_embed(self, text: torch.Tensor, n: int, drop_text: bool) -> torch.Tensor:
        """int64 [B, nt] (-1 padded) -> f32 [B, n, text_dim]   (TextEmbedding.forward, dit.py:49-79)"""
        text = text.to(device=self.device, dtype=torch.int64).contiguous()
        B, nt = text.shape
        out = torch.empty(B, n, self.text_dim, dtype=f32, device=self.device)
        ws = torch.empty(self.lib.f5b_dit_text_ws_bytes(self.handle, B, n), dtype=torch.uint8, device=self.device)
        L.check(self.lib.f5b_dit_text_embed(self.handle, text.data_ptr(), nt, B, n, int(drop_text), out.data_ptr(), ws.data_ptr(),
                                            L.stream()), "f5b_dit_text_embed")
        return out

    def input_const(self, cond, text_embed: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """[cond | text_embed] W^T + b -> f32 [B, n, D]; cond None = drop_audio_cond (dit.py:92-95)."""
        B, n, _ = text_embed.shape
        c0 = out if out is not None else torch.empty(B, n, self.dim, dtype=f32, device=self.device)
        ws = torch.empty(B * n * (128 + self.text_dim) * (4 if self.precision == "tf32" else 2), dtype=torch.uint8, device=self.device)
        if cond is not None:
            cond = cond.to(device=self.device, dtype=f32).contiguous()
        L.check(self.lib.f5b_dit_input_const(self.handle, L.ptr(cond), text_embed.data_ptr(), B, n, c0.data_ptr(), ws.data_ptr(),
                                             L.stream()), "f5b_dit_input_const")
        return c0

    def pack_state(self, y: torch.Tensor, yb: torch.Tensor | None = None) -> torch.Tensor:
        """ODE state / DiT input f32 [..., mel] -> the zero-padded [rows, 128] operand of the input GEMM in the engine's operand
        type (bf16, or fp32 rounded to tf32)."""
        from ... import ops
        rows = y.numel() // self.mel_dim
        if yb is None:
            yb = torch.empty(rows, 128, dtype=self.act_dtype, device=self.device)
        (ops.pack_tf32 if self.precision == "tf32" else ops.pack_bf16)(y.view(rows, self.mel_dim), yb, self.mel_dim, 128)
        return yb

    def forward(self, x_bf16, Bx, c0, Bf, n, mod, mod_bstride, lens, pred, ws=None, attn_drop=None):
        """x_bf16: the packed state from `pack_state` (bf16, or fp32 in the tf32 mode); ws: a private workspace (default: the
        engine's shared one); attn_drop: None or (p, seed, device word tensor or None) — the reference's SDPA dropout at inference
        (modules.py:490), see DiT.set_attn_dropout"""
        ap, aseed, adev = attn_drop if attn_drop is not None else (0.0, 0, None)
        L.check(self.lib.f5b_dit_set_attn_dropout(float(ap), int(aseed) & ((1 << 64) - 1), L.ptr(adev)), "f5b_dit_set_attn_dropout")
        assert x_bf16.dtype == self.act_dtype
        nbytes = self.lib.f5b_dit_workspace_bytes(self.handle, Bf, n)
        if ws is None:
            ws = self.workspace(nbytes)
        assert ws.numel() >= nbytes and c0.is_contiguous() and pred.is_contiguous()
        L.check(self.lib.f5b_dit_forward(self.handle, x_bf16.data_ptr(), Bx, c0.data_ptr(), Bf, n, mod.data_ptr(), mod_bstride,
                                         L.ptr(lens), self.rope_table(n).data_ptr(), pred.data_ptr(), ws.data_ptr(), ws.numel(),
                                         L.stream()), "f5b_dit_forward")
        return pred


class DiT(nn.Module):
    """Same constructor / forward / clear_cache as the reference DiT (dit.py:103-233)."""

    def __init__(self, *, dim, depth=8, heads=8, dim_head=64, dropout=0.1, ff_mult=4, mel_dim=100, text_num_embeds=256,
                 text_dim=None, text_mask_padding=True, qk_norm=None, conv_layers=0, pe_attn_head=None,
                 long_skip_connection=False, checkpoint_activations=False, precision="bf16"):
        super().__init__()
        if qk_norm is not None:
            raise NotImplementedError("qk_norm is unused by every F5TTS config of the reference (configs/*.yaml) and is not built")
        if long_skip_connection:
            raise NotImplementedError("long_skip_connection is unused by the reference's configs and is not built")
        if dim_head != 64 or heads * dim_head != dim:
            raise NotImplementedError("the CUDA path is built for dim_head=64 and heads*dim_head == dim (all F5TTS configs)")
        if text_dim is None:
            text_dim = mel_dim
        self.time_embed = _TimestepEmbedding(dim)
        self.text_embed = _TextEmbedding(text_num_embeds, text_dim, mask_padding=text_mask_padding, conv_layers=conv_layers)
        self.text_cond, self.text_uncond = None, None  # text cache (dit.py:131)
        self.input_embed = _InputEmbedding(mel_dim, text_dim, dim)
        self.rotary_embed = _RotaryEmbedding(dim_head)
        self.dim, self.depth, self.heads, self.dim_head, self.ff_mult = dim, depth, heads, dim_head, ff_mult
        self.mel_dim, self.text_dim, self.conv_layers = mel_dim, text_dim, conv_layers
        self.text_mask_padding, self.pe_attn_head = text_mask_padding, pe_attn_head
        self.transformer_blocks = nn.ModuleList([_DiTBlock(dim, heads, dim_head, ff_mult, dropout) for _ in range(depth)])
        self.long_skip_connection = None
        self.norm_out = _AdaLayerNorm(dim, 2)
        self.proj_out = nn.Linear(dim, mel_dim)
        self.checkpoint_activations = checkpoint_activations
        self.precision = precision  # extension: "bf16" (throughput) or "tf32" (1e-3 of the fp32 reference); set_precision() switches
        self.attn_dropout_p = 0.0   # extension: SDPA dropout at inference (set_attn_dropout); 0 = the parity setting
        self.initialize_weights()
        self._engine = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    def initialize_weights(self):
        """AdaLN-zero init, dit.py:162-172"""
        for block in self.transformer_blocks:
            nn.init.constant_(block.attn_norm.linear.weight, 0)
            nn.init.constant_(block.attn_norm.linear.bias, 0)
        nn.init.constant_(self.norm_out.linear.weight, 0)
        nn.init.constant_(self.norm_out.linear.bias, 0)
        nn.init.constant_(self.proj_out.weight, 0)
        nn.init.constant_(self.proj_out.bias, 0)

    # ---- engine management ----
    def invalidate(self):
        self._engine = None
        self.clear_cache()

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .half() move the masters: re-pack lazily
        self._engine = None
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision: str):
        """Operand mode of the inference path: "bf16" (default) or "tf32"; the packed weights are rebuilt lazily."""
        if precision not in ("bf16", "tf32"):
            raise ValueError(f"precision must be 'bf16' or 'tf32', got {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self.invalidate()
        return self

    def set_attn_dropout(self, p: float):
        """The reference passes dropout_p = 0.1 to F.scaled_dot_product_attention unconditionally (model/modules.py:490), so its own
        inference is stochastic in attention even under eval().  Parity is defined at p = 0 (default); p > 0
        reproduces that behaviour statistically (same distribution, the product's own mask stream; bf16 operand mode).  The training
        step has its own switch (`TrainEngine(attn_dropout=)`)."""
        if not 0.0 <= p < 1.0:
            raise ValueError("attn dropout p must be in [0, 1)")
        self.attn_dropout_p = float(p)
        return self

    def engine(self) -> DiTEngine:
        dev = self.proj_out.weight.device
        if self._engine is None or self._engine.device != dev or self._engine.precision != self.precision:
            self._engine = DiTEngine(self, dev, self.precision)
        return self._engine

    def clear_cache(self):
        self.text_cond, self.text_uncond = None, None

    @L.on_own_device
    @torch.no_grad()
    def forward(self, x, cond, text, time, drop_audio_cond, drop_text, mask=None, cache=False, attn_dropout_seed: int | None = None):
        """x, cond [b, n, mel]; text int [b, nt]; time [] or [b]; mask bool [b, n] (a prefix / key-padding mask as built by
        lens_to_mask, cfm.py:152-153) -> [b, n, mel].  No autograd graph: the training step is `train.TrainEngine`.
        attn_dropout_seed (extension): fixes the mask stream when set_attn_dropout(p > 0) is active (default: a fresh draw)."""
        eng = self.engine()
        b, n = x.shape[0], x.shape[1]
        time = time.to(device=eng.device, dtype=f32)
        mod = eng.modulation(time.reshape(-1))
        mod_bstride = 0 if time.ndim == 0 or time.numel() == 1 else eng.mod_dim
        if cache:
            if drop_text:
                if self.text_uncond is None:
                    self.text_uncond = eng.text_embed(text, n, True)
                te = self.text_uncond
            else:
                if self.text_cond is None:
                    self.text_cond = eng.text_embed(text, n, False)
                te = self.text_cond
        else:
            te = eng.text_embed(text, n, drop_text)
        c0 = eng.input_const(None if drop_audio_cond else cond, te)
        xb = eng.pack_state(x.to(device=eng.device, dtype=f32).contiguous())
        lens = None
        if mask is not None:
            lens = mask.sum(dim=-1).to(torch.int32).contiguous()
        pred = torch.empty(b, n, self.mel_dim, dtype=f32, device=eng.device)
        adrop = None
        if self.attn_dropout_p > 0:
            s = int(torch.randint(0, 2 ** 62, (1,)).item()) if attn_dropout_seed is None else int(attn_dropout_seed)
            adrop = (self.attn_dropout_p, s, None)
        eng.forward(xb, b, c0, b, n, mod, mod_bstride, lens, pred, attn_drop=adrop)
        return pred.to(x.dtype)
