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
