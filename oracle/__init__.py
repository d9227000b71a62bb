"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU (torch fp32) restatement of the reference's algorithm for the north-star path
(F5TTSWrapper.generate -> CFM.sample -> MelSpec / DiT / Euler+CFG -> Vocos.decode).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import anything from this package, and only as the checker or the reported CPU
baseline.  The product package `eraxvif5tts_b200` never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * in-tree reference code (cfm.py, backbones/dit.py, modules.py, utils.py): PINNED — the
    restatement is checked against the reference's own modules, loaded unmodified by path in the
    build container (`oracle/ref_shim.py`), and against golden vectors generated from them
    (`tests/golden/*.pt`, generator `oracle/gen_golden.py`).
  * third-party pieces the reference imports but does not vendor (torchdiffeq fixed-grid Euler /
    midpoint, x_transformers RotaryEmbedding / apply_rotary_pos_emb, vocos): restated from their
    published algorithms; the reference holds no test or golden vector for them ->
    "parity unpinned" for those three boundaries.
"""
