"""The alignment / duration-predictor oracle against outputs of the reference's own functions (tests/golden/align_golden.pt, made
by tests/golden/make_align_golden.py from /root/reference): bit-exact alignments, 1e-5 for the predictor."""
import os

import numpy as np
import torch

from oracle import align_oracle as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.pt")


def test_viterbi_and_window_oracle_match_reference_outputs():
    gold = torch.load(GOLD)
    assert len(gold["cases"]) >= 8
    for c in gold["cases"]:
        sim = c["sim"].numpy()
        al, dur = A.viterbi_alignment(sim)
        assert np.array_equal(al, c["viterbi"].numpy()), sim.shape
        assert np.array_equal(dur, c["viterbi"].sum(-1).long().numpy())
        if c["window"] is not None:
            aw, _ = A.windowed_alignment(sim)
            assert np.array_equal(aw, c["window"].numpy()), sim.shape


def test_duration_predictor_oracle_matches_reference_module():
    d = torch.load(GOLD)["dp"]
    out = A.duration_predictor(d["state_dict"], d["ids"], d["mask"], 1)
    assert float((out - d["out"]).abs().max()) <= 1e-5
    pout = A.duration_predictor(d["state_dict"], d["ids"] + 1, d["mask"], 0)
    assert float((pout - d["phoneme_out"]).abs().max()) <= 1e-5
    assert float(out[1, 0, 13:].abs().max()) == 0.0  # masked positions


def test_alignment_properties():
    """every frame up to the last belongs to exactly one token and segments are monotone (when the search does not stop early)"""
    rng = np.random.default_rng(0)
    sim = rng.standard_normal((2, 9, 70)).astype(np.float32)
    for al, dur in (A.viterbi_alignment(sim), A.windowed_alignment(sim)):
        assert set(np.unique(al)) <= {0.0, 1.0}
        assert (al.sum(1) <= 1).all()
        for b in range(2):
            rows, cols = np.nonzero(al[b])
            order = np.argsort(cols, kind="stable")
            assert (np.diff(rows[order]) >= 0).all()
        assert (dur.sum(-1) <= 70).all()


def test_progressive_alignment_matches_reference_golden():
    """progressive_monotonic_alignment (alignment_utils.py:260-334) runs on the host in the product (a sequential greedy sweep); it is
    bit-exact against the reference's own function on every golden case, including the batch-total score bookkeeping that only shows
    with b > 1 (the b = 1 twin of the same item refines differently)."""
    import os
    import torch
    from eraxvif5tts_b200.model.alignment_utils import monotonic_alignment_search, progressive_monotonic_alignment
    d = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.pt"), weights_only=False)
    twins_differ = 0
    for c in d["cases"]:
        got = progressive_monotonic_alignment(c["sim"])
        assert got.dtype == c["sim"].dtype and torch.equal(got, c["progressive"]), tuple(c["sim"].shape)
        assert torch.equal(monotonic_alignment_search(c["sim"], "progressive"), c["progressive"])
        if "progressive_item0" in c:
            one = progressive_monotonic_alignment(c["sim"][:1])
            assert torch.equal(one, c["progressive_item0"])
            twins_differ += int(not torch.equal(one[0], c["progressive"][0]))
    assert twins_differ > 0
