"""CPU tests: host-side logic of the drop-in classes, state_dict key compatibility with the reference, C-ABI symbol export,
and the utterance sharding (gloo, world_size 2).  No CUDA compute is called."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from eraxvif5tts_b200 import build, _lib
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "f5b200.h")).read()
    declared = set(re.findall(r"\b(f5b_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/f5b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.load().f5b_abi_version() == 1


def test_python_enum_constants_match_the_c_header():
    """the epilogue / activation codes the Python side passes in F5bGemmArgs are the header's (a drifted constant would select another
    epilogue silently)"""
    from eraxvif5tts_b200 import _lib
    header = open(os.path.join(ROOT, "include", "f5b200.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(F5B_(?:EPI|ACT)_[A-Z0-9_]+)\s*=\s*(\d+)", header)}
    assert len(enums) >= 9
    for name, val in enums.items():
        py = name[len("F5B_"):]
        if hasattr(_lib, py):
            assert getattr(_lib, py) == val, (name, val, getattr(_lib, py))
    for py in ("EPI_BF16", "EPI_F32", "EPI_QKV_ROPE", "EPI_GATE_RESID", "EPI_BF16_DUAL", "ACT_NONE", "ACT_GELU_TANH", "ACT_GELU_ERF", "ACT_SILU"):
        assert "F5B_" + py in enums and getattr(_lib, py) == enums["F5B_" + py]


def test_no_cpu_fallback():
    from eraxvif5tts_b200 import _lib, ops
    with pytest.raises(_lib.F5bError):
        ops.ln_modulate(torch.zeros(4, 64), None, None, 0, 0, 4)  # CPU tensor must be rejected, not computed
    from eraxvif5tts_b200.infer import F5TTSWrapper
    with pytest.raises(RuntimeError):
        F5TTSWrapper(model_name="F5TTS_Base", vocab_char_map={"a": 0}, device="cpu")


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "eraxvif5tts_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{fn} imports the oracle"


def test_state_dict_keys_match_reference_layout():
    """SURVEY.md §10 key list (taken from the real reference modules; oracle.weights mirrors it)."""
    from eraxvif5tts_b200.model import DiT
    from oracle.f5_oracle import DiTConfig
    from oracle.weights import dit_key_shapes
    cfg = DiTConfig.tiny()
    d = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim, text_num_embeds=cfg.text_num_embeds,
            text_dim=cfg.text_dim, text_mask_padding=False, conv_layers=cfg.conv_layers, pe_attn_head=1)
    ours = {k: tuple(v.shape) for k, v in d.state_dict().items()}
    ref = {k[len("transformer."):]: tuple(s) for k, s, _ in dit_key_shapes(cfg)}
    ref["rotary_embed.inv_freq"] = (32,)
    assert ours == ref
    # AdaLN-zero init (dit.py:162-172)
    assert float(d.norm_out.linear.weight.abs().max()) == 0 and float(d.proj_out.weight.abs().max()) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/f5_tts"), reason="reference tree only in the build container")
def test_state_dict_keys_match_live_reference():
    from eraxvif5tts_b200.model import DiT
    from oracle import ref_shim
    ref = ref_shim.load()
    kw = dict(dim=128, depth=2, heads=2, ff_mult=2, mel_dim=100, text_num_embeds=40, text_dim=64, text_mask_padding=False, conv_layers=2,
              pe_attn_head=1)
    a = {k: tuple(v.shape) for k, v in DiT(**kw).state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.dit.DiT(**kw).state_dict().items()}
    assert a == b


def test_vocos_keys_match_upstream_layout():
    from eraxvif5tts_b200 import Vocos
    from oracle.f5_oracle import VocosConfig
    from oracle.weights import vocos_key_shapes
    vc = VocosConfig()
    ours = {k: tuple(v.shape) for k, v in Vocos().state_dict().items()}
    ref = {k: tuple(s) for k, s, _ in vocos_key_shapes(vc)}
    ref["head.istft.window"] = (1024,)
    assert ours == ref


def test_chunk_text_and_tokenizer(tmp_path):
    from eraxvif5tts_b200.infer.utils_infer import chunk_text, resolve_arch
    from eraxvif5tts_b200.model.utils import get_tokenizer, lens_to_mask, list_str_to_idx, mask_from_frac_lengths
    chunks = chunk_text("One. Two, three! Four? Five; six: seven.", max_chars=12)
    assert all(len(c.encode()) <= 12 for c in chunks) and "".join(chunks).replace(" ", "") == "One.Two,three!Four?Five;six:seven."
    assert chunk_text("", 10) == []
    vf = tmp_path / "vocab.txt"
    vf.write_text(" \na\nb\nc\n", encoding="utf-8")
    m, n = get_tokenizer(str(vf), "custom")
    assert n == 4 and m[" "] == 0 and m["c"] == 3  # first line is a literal space
    idx = list_str_to_idx(["ab c", "zz"], m)
    assert idx.tolist() == [[1, 2, 0, 3], [0, 0, -1, -1]]  # unknown -> 0, pad -1
    assert lens_to_mask(torch.tensor([1, 3])).tolist() == [[True, False, False], [True, True, True]]
    torch.manual_seed(0)
    mk = mask_from_frac_lengths(torch.tensor([10, 20]), torch.tensor([0.5, 1.0]))
    assert mk.shape == (2, 20) and int(mk[0].sum()) == 5 and int(mk[1].sum()) == 20
    assert resolve_arch("F5TTS_Base")["depth"] == 22 and resolve_arch("F5TTS_Small")["dim"] == 768
    with pytest.raises(ValueError):
        resolve_arch("nope")


def test_melspec_filterbank_matches_torchaudio():
    import torchaudio
    from eraxvif5tts_b200.model.modules import MelSpec, melscale_fbanks_htk
    fb = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000, norm=None, mel_scale="htk")
    assert (melscale_fbanks_htk(513, 0.0, 12000.0, 100, 24000) - fb).abs().max() < 1e-6
    m = MelSpec()
    r = m.fb_ranges
    for j in range(100):  # every non-zero of column j lies inside [f0, f1)
        nzr = (fb[:, j] > 0).nonzero().flatten()
        assert int(r[j, 0]) <= int(nzr.min()) and int(nzr.max()) < int(r[j, 1])


def test_cfm_forward_needs_cuda():
    """CFM.forward (loss, forward only) has no CPU path either"""
    from eraxvif5tts_b200 import _lib
    from eraxvif5tts_b200.model import CFM, DiT
    m = CFM(transformer=DiT(dim=128, depth=1, heads=2, ff_mult=2, text_dim=64, conv_layers=1, text_num_embeds=10), mel_spec_kwargs={})
    with pytest.raises(_lib.F5bError):
        m(torch.zeros(1, 8, 100), torch.zeros(1, 3, dtype=torch.long))


def test_shard_indices_cover_everything_once():
    from eraxvif5tts_b200.parallel import shard_indices
    for n, w, bs in ((256, 8, 16), (10, 4, 3), (3, 8, 16), (0, 2, 4)):
        seen = []
        for r in range(w):
            for b in shard_indices(n, r, w, bs):
                assert len(b) <= bs
                seen += b
        assert sorted(seen) == list(range(n))
    assert shard_indices(256, 1, 8, 16)[0] == list(range(16, 32))


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from eraxvif5tts_b200.parallel import shard_indices, gather_counts, max_over_ranks, allreduce_flat_, broadcast_flat_
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
r = dist.get_rank()
mine = shard_indices(37, r, 2, 4)
cnt = sum(len(b) for b in mine)
counts = gather_counts(cnt)
t = max_over_ranks(1.0 + r)
g = torch.full((1000,), float(r + 1))
scale = allreduce_flat_(g)
assert scale == 0.5 and bool((g * scale == 1.5).all())
w = torch.full((10,), float(r))
broadcast_flat_(w, src=1)
assert bool((w == 1.0).all())
dist.barrier()
assert sum(counts) == 37 and t == 2.0, (counts, t)
print("rank", r, "ok", counts)
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=29731))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok [19, 18]" in o or "ok [20, 17]" in o for o in outs), outs


def _tiny_cfm():
    from eraxvif5tts_b200.model import CFM, DiT
    return CFM(transformer=DiT(dim=128, depth=3, heads=2, ff_mult=2, text_dim=64, conv_layers=1, text_num_embeds=10), mel_spec_kwargs={})


def test_load_checkpoint_formats(tmp_path):
    """reference checkpoint layouts (infer/f5tts_wrapper.py:201-254, model/trainer.py:521-598, model_pruning/*): .safetensors flat
    EMA dict, .pt with ema_model_state_dict (ema_model. prefix + initted/step), .pt with model_state_dict (+ DDP 'module.' prefix,
    legacy mel keys), pruned .pt with pruning_info"""
    from safetensors.torch import save_file
    from eraxvif5tts_b200.infer.utils_infer import load_checkpoint
    torch.manual_seed(0)
    src = _tiny_cfm()
    with torch.no_grad():
        for p in src.parameters():
            p.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in src.state_dict().items()}

    def check(path, **kw):
        m = _tiny_cfm()
        load_checkpoint(m, str(path), "cpu", **kw)
        for k, v in sd.items():
            assert torch.equal(m.state_dict()[k], v), k

    ema = {"ema_model." + k: v for k, v in sd.items()}
    ema.update({"initted": torch.tensor(True), "step": torch.tensor(7)})
    save_file({k: v.contiguous() for k, v in ema.items() if k not in ("initted", "step")}, str(tmp_path / "m.safetensors"))
    check(tmp_path / "m.safetensors", use_ema=True)
    torch.save({"ema_model_state_dict": ema, "model_state_dict": {k: torch.zeros_like(v) for k, v in sd.items()}, "update": 5},
               tmp_path / "ema.pt")
    check(tmp_path / "ema.pt", use_ema=True)
    legacy = {"module." + k: v for k, v in sd.items()}
    legacy["mel_spec.mel_stft.mel_scale.fb"] = torch.zeros(3)
    legacy["mel_spec.mel_stft.spectrogram.window"] = torch.zeros(3)
    torch.save({"model_state_dict": legacy}, tmp_path / "plain.pt")
    check(tmp_path / "plain.pt", use_ema=False)
    torch.save({"model_state_dict": sd, "pruning_info": {"kept_blocks": [0, 1, 2], "original_depth": 5}}, tmp_path / "pruned.pt")
    check(tmp_path / "pruned.pt", use_ema=False)


def test_convert_char_to_pinyin_latin_path():
    from eraxvif5tts_b200.infer.f5tts_wrapper import convert_char_to_pinyin
    out = convert_char_to_pinyin(["Xin chào; “thế giới”"])
    assert out[0] == list('Xin chào, "thế giới"')


def test_wrapper_duration_rule_matches_reference_formula():
    """f5tts_wrapper.py:498-503: duration = ref_len + int(ref_len / ref_bytes * gen_bytes / speed), speed 0.3 for < 10 bytes"""
    from eraxvif5tts_b200.infer.f5tts_wrapper import F5TTSWrapper
    w = F5TTSWrapper.__new__(F5TTSWrapper)
    w.ref_text, w.ref_audio_len, w.target_sample_rate, w.hop_length = "abcdefghij. ", 200, 24000, 256
    assert w._chunk_duration("x" * 24, 1.0, None) == 200 + int(200 / 12 * 24 / 1.0)
    assert w._chunk_duration("short", 1.0, None) == 200 + int(200 / 12 * 5 / 0.3)
    assert w._chunk_duration("anything", 1.0, 2.0) == int(2.0 * 24000 / 256)


def test_lr_and_ema_schedules_match_torch_and_ema_pytorch_rules():
    """WarmupLinearDecay == SequentialLR(LinearLR, LinearLR) of trainer.py:1179-1188; EmaSchedule == ema_pytorch defaults"""
    from torch.optim.lr_scheduler import LinearLR, SequentialLR
    from eraxvif5tts_b200.optim import EmaSchedule, WarmupLinearDecay
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=7.5e-5)
    w, total = 5, 20
    sch = SequentialLR(opt, [LinearLR(opt, 1e-8, 1.0, w), LinearLR(opt, 1.0, 1e-8, total - w)], milestones=[w])
    ours = WarmupLinearDecay(7.5e-5, w, total)
    for u in range(total):
        assert abs(opt.param_groups[0]["lr"] - ours.lr(u)) <= 1e-12 + 1e-6 * ours.lr(u), u
        opt.step()
        sch.step()
    e = EmaSchedule()
    assert e.decay_for_call(2) is None and e.decay_for_call(1) == "copy" and e.decay_for_call(101) == "copy"
    # ema_pytorch: update() gates on the pre-increment step, get_current_decay() reads self.step after the increment:
    # call 111 -> step 110 -> epoch = 111 - 100 - 1 = 10
    d = e.decay_for_call(111)
    assert abs(d - (1 - (1 + 10) ** (-2 / 3))) < 1e-12
    assert e.decay_for_call(10 ** 8 + 1) == 0.9999
    # world > 1: the reference multiplies the warm-up by num_processes (trainer.py:1179-1181) BECAUSE accelerate's prepared scheduler
    # steps num_processes times per optimizer update; the effective warm-up stays num_warmup_updates updates
    world, nw, per_proc_total = 4, 3, 10
    opt = torch.optim.SGD([p], lr=7.5e-5)
    wu, tot = nw * world, per_proc_total
    sch = SequentialLR(opt, [LinearLR(opt, 1e-8, 1.0, wu), LinearLR(opt, 1.0, 1e-8, max(tot - wu, 1))], milestones=[wu])
    ours = WarmupLinearDecay(7.5e-5, wu, tot, steps_per_update=world)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for u in range(per_proc_total):
            assert abs(opt.param_groups[0]["lr"] - ours.lr(u)) <= 1e-12 + 1e-6 * ours.lr(u), u
            opt.step()
            for _ in range(world):
                sch.step()
    assert abs(ours.lr(nw) - 7.5e-5) < 1e-12  # full rate after num_warmup_updates optimizer updates, whatever the world size


def test_trainer_checkpoint_search_order_and_state_dict_cleaning(tmp_path):
    """trainer.py:600-728: model_last.pt > latest model_<n> (.pt or .safetensors) > first pretrained_*; weights under
    model_state_dict / ema_model_state_dict / state_dict / model with ema_model. / module. / _orig_mod. prefixes stripped"""
    import types
    from eraxvif5tts_b200.model.trainer import Trainer
    from eraxvif5tts_b200.train import TrainEngine
    d = tmp_path / "ck"
    d.mkdir()
    find = lambda: Trainer.find_checkpoint(types.SimpleNamespace(checkpoint_path=str(d)))  # noqa: E731
    assert find() is None
    (d / "notes.txt").write_text("x")
    assert find() is None
    (d / "pretrained_b.safetensors").write_bytes(b"")
    (d / "pretrained_a.pt").write_bytes(b"")
    assert find().endswith("pretrained_a.pt")
    (d / "model_900.pt").write_bytes(b"")
    (d / "model_1200.safetensors").write_bytes(b"")
    assert find().endswith("model_1200.safetensors")
    (d / "model_last.pt").write_bytes(b"")
    assert find().endswith("model_last.pt")
    e = tmp_path / "odd"
    e.mkdir()
    (e / "weights.pt").write_bytes(b"")
    with pytest.raises(FileNotFoundError):
        Trainer.find_checkpoint(types.SimpleNamespace(checkpoint_path=str(e)))
    t = torch.zeros(1)
    clean = TrainEngine._model_state_dict
    assert set(clean({"model_state_dict": {"a.w": t, "b.w": t}})) == {"a.w", "b.w"}
    ema = {f"ema_model.l{i}.w": t for i in range(10)}  # the prefix must be carried by >= 80 % of the keys (reference rule)
    ema.update(initted=t, step=t)
    assert set(clean({"ema_model_state_dict": ema})) == {f"l{i}.w" for i in range(10)}
    assert set(clean({"state_dict": {"module.a.w": t, "module.b.w": t}})) == {"a.w", "b.w"}
    assert set(clean({"model": {"_orig_mod.a.w": t}})) == {"a.w"}
    assert set(clean({"state_dict_loaded_from_safetensors": {"a.w": t}})) == {"a.w"}
    assert set(clean({"model_state_dict": {}, "ema_model_state_dict": {"ema_model.a.w": t}})) == {"a.w"}  # empty dicts are skipped
    with pytest.raises(KeyError):
        clean({"optimizer_state_dict": {}})


class _FakeFrames:
    def __init__(self, lens):
        self.lens = lens

    def __len__(self):
        return len(self.lens)

    def get_frame_len(self, i):
        return self.lens[i]


def _load_reference_dataset_module():
    """the reference's dataset.py executed with its heavy imports stubbed (only DynamicBatchSampler / collate_fn are used)"""
    import importlib.util
    import sys
    import types
    path = "/root/reference/src/f5_tts/model/dataset.py"
    if not os.path.exists(path):
        return None
    stubs = {}
    for name in ("torchaudio", "datasets", "f5_tts", "f5_tts.model", "f5_tts.model.modules", "f5_tts.model.utils", "tqdm"):
        if name not in sys.modules:
            stubs[name] = types.ModuleType(name)
    stubs.get("datasets", sys.modules.get("datasets")).Dataset = getattr(sys.modules.get("datasets"), "Dataset", object)
    stubs.get("datasets", sys.modules.get("datasets")).load_from_disk = lambda *a, **k: None
    if "f5_tts.model.modules" in stubs:
        stubs["f5_tts.model.modules"].MelSpec = object
    if "f5_tts.model.utils" in stubs:
        stubs["f5_tts.model.utils"].default = lambda v, d: v if v is not None else d
    if "tqdm" in stubs:
        stubs["tqdm"].tqdm = lambda it, **k: it
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_ref_dataset", path)
        mod = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception:  # noqa: BLE001  (an import the stubs do not cover)
            return None
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_dynamic_batch_sampler_and_collate():
    from torch.utils.data import SequentialSampler
    from eraxvif5tts_b200.data import DynamicBatchSampler, collate_fn, collate_token_major, shard_batches
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(50, 3000, (500,), generator=g).tolist() + [5000]  # one utterance above the budget
    src = _FakeFrames(lens)
    bs = DynamicBatchSampler(SequentialSampler(src), frames_threshold=3200, max_samples=8, random_seed=666)
    flat = [i for b in bs.batches for i in b]
    assert sorted(flat) == list(range(500)) and 500 not in flat          # everything that fits, exactly once
    assert all(sum(lens[i] for i in b) <= 3200 and 1 <= len(b) <= 8 for b in bs.batches)
    assert [lens[i] for i in flat] == sorted(lens[:500])                 # length-sorted packing (high padding efficiency)
    bs.set_epoch(3)
    e3 = list(bs)
    bs.set_epoch(3)
    assert e3 == list(bs)
    bs.set_epoch(4)
    assert e3 != list(bs) and sorted(map(tuple, e3)) == sorted(map(tuple, bs.batches))
    shards = [shard_batches(e3, r, 4) for r in range(4)]
    assert len({len(s) for s in shards}) == 1 and sum(len(s) for s in shards) == len(e3) // 4 * 4
    assert not set(map(tuple, shards[0])) & set(map(tuple, shards[1]))
    ref = _load_reference_dataset_module()
    if ref is not None:  # identical batches and epoch order as the reference class
        rb = ref.DynamicBatchSampler(SequentialSampler(src), frames_threshold=3200, max_samples=8, random_seed=666)
        assert rb.batches == bs.batches
        rb.set_epoch(3)
        assert list(rb) == e3
    items = [dict(mel_spec=torch.randn(1, 100, t, generator=g), text="x" * (t // 10)) for t in (120, 77, 301)]
    a, b = collate_fn(items), collate_token_major(items, pin=False)
    assert a["mel"].shape == (3, 100, 301) and b["mel"].shape == (3, 301, 100)
    assert torch.equal(a["mel"].permute(0, 2, 1), b["mel"]) and torch.equal(a["mel_lengths"], b["mel_lengths"])
    assert a["text"] == b["text"] and a["text_lengths"].tolist() == [12, 7, 30]
    if ref is not None:
        r = ref.collate_fn(items)
        assert torch.equal(r["mel"], a["mel"]) and torch.equal(r["mel_lengths"], a["mel_lengths"]) and r["text"] == a["text"]


def test_dropout_mask_statistics():
    """keep rate of the generator the kernels and the oracle share: Bernoulli(1 - p) per element, kept values scaled by 1 / (1 - p),
    no visible correlation between sites / layers / neighbouring elements"""
    from oracle import f5_oracle as _O
    for p in (0.1, 0.5):
        a = _O.dropout_multipliers(p, 7, 3, 0, (64, 4096))
        b = _O.dropout_multipliers(p, 7, 3, 1, (64, 4096))
        keep_a, keep_b = (a > 0).float(), (b > 0).float()
        assert abs(float(keep_a.mean()) - (1 - p)) < 3e-3
        assert abs(float(a.mean()) - 1.0) < 6e-3
        assert float(a.max()) == pytest.approx(1 / (1 - p), rel=1e-4)
        both = float((keep_a * keep_b).mean())
        assert abs(both - (1 - p) ** 2) < 4e-3
        nb = float((keep_a[:, 1:] * keep_a[:, :-1]).mean())
        assert abs(nb - (1 - p) ** 2) < 4e-3


def test_trainer_batching_and_sharding_host_logic():
    """model.Trainer's host side without a GPU (the engine is not constructed): frame / sample batching, per-process sharding with
    equal batch counts on every rank, epoch reshuffling under resumable_with_seed"""
    import torch
    from eraxvif5tts_b200.data import DynamicBatchSampler
    from eraxvif5tts_b200.model.trainer import Trainer

    class DS(torch.utils.data.Dataset):
        lens = [50 + 7 * (i % 13) for i in range(57)]

        def get_frame_len(self, i):
            return self.lens[i]

        def __len__(self):
            return len(self.lens)

        def __getitem__(self, i):
            return dict(mel_spec=torch.zeros(100, self.lens[i]), text="ab")

    ds = DS()

    def make(rank, world, kind):
        t = Trainer.__new__(Trainer)
        t.batch_size_type, t.batch_size_per_gpu, t.max_samples = kind, (400 if kind == "frame" else 5), 6
        t.rank, t.num_processes = rank, world
        return t

    ref = DynamicBatchSampler(torch.utils.data.SequentialSampler(ds), 400, max_samples=6, random_seed=3)
    for world in (1, 2, 3):
        per_rank = []
        for r in range(world):
            t = make(r, world, "frame")
            sampler, gen = t._batches(ds, 3)
            assert gen is None and sampler.batches == ref.batches
            e0 = t._epoch_batches(ds, sampler, gen, 0)
            e1 = t._epoch_batches(ds, sampler, gen, 1)
            assert e0 != e1 or len(e0) <= 1  # set_epoch reshuffles
            assert all(sum(ds.lens[i] for i in b) <= 400 and len(b) <= 6 for b in e0)
            per_rank.append(e0)
        assert len({len(p) for p in per_rank}) == 1  # every rank steps the same number of times (the all-reduce never starves)
        flat = [i for p in per_rank for b in p for i in b]
        assert len(flat) == len(set(flat))
        if world == 1:
            assert sorted(flat) == list(range(len(ds)))
    t = make(0, 2, "sample")
    sampler, gen = t._batches(ds, 7)
    assert sampler is None
    a = t._epoch_batches(ds, sampler, gen, 0)
    b = t._epoch_batches(ds, sampler, gen, 1)
    assert a != b and all(len(x) <= 5 for x in a) and len(a) == (len(ds) + 4) // 5 // 2
    import pytest
    with pytest.raises(ValueError):
        make(0, 1, "tokens")._batches(ds, None)


def test_silence_aware_reference_clipping_matches_pydub_rules():
    """preprocess_reference's clip_short (f5tts_wrapper.py:272-301) = pydub.silence.split_on_silence + the 6 s / 12 s collection rule.
    pydub is absent offline: detect_silence is checked against a literal transcription of pydub's per-slice loop, and the clipping
    rule on a signal whose silences are known by construction."""
    import math
    from eraxvif5tts_b200.infer.f5tts_wrapper import clip_reference, detect_silence, split_on_silence

    def literal(x, sr, min_silence_len, silence_thresh, seek_step):
        seg_len = int(round(1000.0 * len(x) / sr))
        if seg_len < min_silence_len:
            return []
        th = 10 ** (silence_thresh / 20)
        last = seg_len - min_silence_len
        starts = list(range(0, last + 1, seek_step))
        if last % seek_step:
            starts.append(last)
        sil = []
        for i in starts:
            sl = x[int(i * sr / 1000): int((i + min_silence_len) * sr / 1000)]
            if (math.sqrt(float((sl.double() ** 2).mean())) if len(sl) else 0.0) <= th:
                sil.append(i)
        if not sil:
            return []
        out, prev = [], sil.pop(0)
        cur = prev
        for s_ in sil:
            if not (s_ == prev + seek_step) and s_ > prev + min_silence_len:
                out.append([cur, prev + min_silence_len])
                cur = s_
            prev = s_
        out.append([cur, prev + min_silence_len])
        return out

    g = torch.Generator().manual_seed(0)
    sr = 8000
    x = torch.cat([torch.randn(int(d * sr), generator=g) * a for d, a in
                   ((1.3, 0.2), (1.4, 0.0005), (2.2, 0.3), (0.3, 0.001), (3.0, 0.25), (1.2, 0.0), (4.0, 0.2), (1.5, 0.0002), (5.0, 0.3))]).clamp(-1, 1)
    for args in ((1000, -50, 10), (100, -40, 10), (300, -45, 7)):
        assert detect_silence(x, sr, *args) == literal(x, sr, *args), args
    assert detect_silence(x, sr, 1000, -50, 10) == [[1300, 2700], [8200, 9400], [13400, 14900]]
    pieces = split_on_silence(x, sr, 1000, -50, 1000, 10)
    assert [round(len(p) / sr, 2) for p in pieces] == [2.0, 6.8, 5.35, 5.75]  # paddings of 1 s meet half way inside shorter silences
    # rule 1: stop once > 6 s are collected and the next piece would pass 12 s -> 2.0 + 6.8
    assert abs(len(clip_reference(x, sr)) / sr - 8.8) < 1e-3
    # no silence at all: hard cut at 12 s (rule 3); short audio: untouched
    noisy = torch.randn(20 * sr, generator=g) * 0.2
    assert len(clip_reference(noisy, sr)) == 12 * sr
    short = torch.randn(3 * sr, generator=g) * 0.2
    assert torch.equal(clip_reference(short, sr), short)
    # only short pauses (200 ms below -40 dBFS): rule 2 finds them where rule 1 cannot
    y = torch.cat([torch.cat((torch.randn(int(2.3 * sr), generator=g) * 0.2, torch.randn(int(0.2 * sr), generator=g) * 0.002)) for _ in range(8)])
    w = clip_reference(y, sr)
    assert 6.0 < len(w) / sr <= 12.0 and len(w) < len(y)


def test_silence_edge_trimming_and_reference_file_helpers(tmp_path):
    """remove_silence_edges (utils_infer.py:273-286 = F5TTSWrapper._remove_silence_edges) against a literal transcription of pydub's
    two loops (10 ms chunks from the start while dBFS < threshold; 1 ms slices from the end until one is louder), then the file-level
    mirrors preprocess_ref_audio_text (:292-360) and remove_silence_for_generated_wav (:569-578) on PCM wav files."""
    import math
    from eraxvif5tts_b200.infer import utils_infer as UI
    from eraxvif5tts_b200.infer.f5tts_wrapper import _read_wav, _write_wav, remove_silence_edges

    def literal(x, sr, thr):
        def seg_len(sig):
            return int(round(1000.0 * len(sig) / sr))

        def dbfs(sl):
            r = math.sqrt(float((sl.double() ** 2).mean())) if len(sl) else 0.0
            return 20 * math.log10(r) if r > 0 else -math.inf
        trim = 0
        while trim < seg_len(x) and dbfs(x[int(trim * sr / 1000): int(min(trim + 10, seg_len(x)) * sr / 1000)]) < thr:
            trim += 10
        x = x[int(min(trim, seg_len(x)) * sr / 1000):]
        dur = len(x) / sr
        for i in range(seg_len(x) - 1, -1, -1):
            if dbfs(x[int(i * sr / 1000): int((i + 1) * sr / 1000)]) > thr:
                break
            dur -= 0.001
        return x[: int(int(dur * 1000) * sr / 1000)]

    g = torch.Generator().manual_seed(1)
    sr = 8000
    for lead, tail, amp in ((0.237, 0.4113, 0.2), (0.0, 0.0, 0.3), (0.055, 1.2, 0.05)):
        x = torch.cat((torch.randn(int(lead * sr), generator=g) * 1e-4, torch.randn(int(1.7 * sr), generator=g) * amp,
                       torch.randn(int(tail * sr), generator=g) * 1e-4)).clamp(-1, 1)
        got, want = remove_silence_edges(x, sr, -42), literal(x, sr, -42)
        assert got.numel() == want.numel() and torch.equal(got, want), (lead, tail, got.numel(), want.numel())
        assert abs(got.numel() / sr - 1.7) < 0.02
    assert remove_silence_edges(torch.zeros(sr), sr, -42).numel() == 0  # all silent
    assert UI.remove_silence_edges(x, -42, sr).numel() == remove_silence_edges(x, sr, -42).numel()
    # preprocess_ref_audio_text: 20 s of speech-like noise with a 1.3 s pause after 7 s -> clipped at the pause, edges trimmed, + 50 ms
    sr = 24000
    sig = torch.cat((torch.zeros(int(0.2 * sr)), torch.randn(7 * sr, generator=g) * 0.2, torch.zeros(int(1.3 * sr)),
                     torch.randn(12 * sr, generator=g) * 0.2)).clamp(-1, 1)
    src = str(tmp_path / "ref.wav")
    _write_wav(src, sig.numpy(), sr)
    msgs = []
    path, text = UI.preprocess_ref_audio_text(src, "xin chào", show_info=msgs.append)
    out, sr2 = _read_wav(path)
    assert sr2 == sr and text == "xin chào. " and any("clipping" in m for m in msgs)
    assert abs(out.shape[-1] / sr - (7.0 + 0.05)) < 0.03  # first piece = lead 0.2 + 7 s + half the pause; both quiet edges trimmed again
    path2, text2 = UI.preprocess_ref_audio_text(src, "Đã có dấu chấm.", clip_short=False, show_info=lambda m: None)
    out2, _ = _read_wav(path2)
    assert text2 == "Đã có dấu chấm. " and abs(out2.shape[-1] / sr - (20.3 + 0.05)) < 0.03
    import pytest
    with pytest.raises(RuntimeError):
        UI.preprocess_ref_audio_text(src, "  ")
    # remove_silence_for_generated_wav: the 1.3 s pause shrinks to 2 x 500 ms
    gen = str(tmp_path / "gen.wav")
    _write_wav(gen, sig[int(0.2 * sr):].numpy(), sr)
    UI.remove_silence_for_generated_wav(gen)
    out3, _ = _read_wav(gen)
    assert abs(out3.shape[-1] / sr - (7 + 1.0 + 12)) < 0.03


def test_small_reference_helpers():
    """model/utils.py: seed_everything, maybe_masked_mean, repetition_found; model/modules.py: get_pos_embed_indices, precompute_freqs_cis"""
    from eraxvif5tts_b200.model import utils as U
    from eraxvif5tts_b200.model import modules as M
    U.seed_everything(3)
    a = torch.rand(2)
    U.seed_everything(3)
    assert torch.equal(a, torch.rand(2))
    t = torch.arange(24, dtype=torch.float32).reshape(2, 4, 3)
    mask = torch.tensor([[True, True, False, False], [False, False, False, False]])
    assert torch.equal(U.maybe_masked_mean(t), t.mean(dim=1))
    mm = U.maybe_masked_mean(t, mask)
    assert torch.allclose(mm[0], t[0, :2].mean(dim=0)) and torch.equal(mm[1], torch.zeros(3))
    assert U.repetition_found("ab" * 12) and not U.repetition_found("the quick brown fox")
    assert M.get_pos_embed_indices(torch.tensor([0, 5]), 6, 8).tolist() == [[0, 1, 2, 3, 4, 5], [5, 6, 7, 7, 7, 7]]
    from eraxvif5tts_b200.model.backbones.dit import precompute_freqs_cis as eng_table
    assert torch.equal(M.precompute_freqs_cis(64, 128), eng_table(64, 128))


def test_plan_ragged_batches_budget_and_coverage():
    from eraxvif5tts_b200.infer.f5tts_wrapper import plan_ragged_batches
    import random
    rnd = random.Random(3)
    durs = [rnd.randint(300, 1900) for _ in range(57)] + [5000]
    batches = plan_ragged_batches(durs, 8000)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(len(durs)))                       # every item exactly once
    for b in batches:
        longest = max(durs[i] for i in b)
        assert len(b) == 1 or len(b) * longest <= 8000          # padded size within the budget
        assert durs[b[0]] == longest                            # sorted: the first item is the longest
    # bucketing bounds the padding waste: padded frames within 25 % of the real ones for this spread
    padded = sum(len(b) * max(durs[i] for i in b) for b in batches)
    assert padded <= 1.25 * sum(durs)
    assert plan_ragged_batches([], 100) == [] and plan_ragged_batches([7], 1) == [[0]]
