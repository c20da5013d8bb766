"""CPU: the oracle restatement (oracle/nmn_oracle.py) against the golden vectors that
oracle/make_golden.py recorded from the reference's own NeuralModuleNetwork (tests/golden/nmn_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import nmn_oracle
from probnmn_clevr_b200.synthetic import make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "nmn_golden.npz")


@pytest.fixture(scope="module")
def setup():
    vocab = Vocabulary.clevr()
    return vocab, make_nmn_state_dict(vocab, 0), np.load(GOLDEN)


@pytest.mark.parametrize("name", ["semantic", "sampled", "garbage"])
def test_oracle_matches_reference_golden(setup, name):
    vocab, sd, g = setup
    programs = torch.from_numpy(g[f"{name}.programs"])
    answers = torch.from_numpy(g[f"{name}.answers"])
    feats = make_features(programs.shape[0], 0)
    with torch.no_grad():
        out = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
        out_na = nmn_oracle.nmn_forward(sd, vocab, feats, programs)
    assert np.array_equal(out["valid"].numpy(), g[f"{name}.valid"])
    assert np.array_equal(out["predictions"].numpy(), g[f"{name}.predictions"])
    # fp32 CPU vs fp32 CPU of the same ops: 1e-5 covers thread-count dependent summation order
    np.testing.assert_allclose(out["logits"].numpy(), g[f"{name}.logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss"].numpy(), g[f"{name}.loss"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out_na["loss"].numpy(), g[f"{name}.loss_noanswer"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["final"].sum(dim=(1, 2, 3)).numpy(), g[f"{name}.final_sum"], rtol=1e-4, atol=1e-3)


def test_appendix_a_truth_table(setup):
    """interpreter semantics the reference exhibits, straight from the golden file"""
    _, _, g = setup
    assert g["semantic.valid"].tolist() == [1, 1, 1, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]
    assert g["garbage.valid"].sum() == 0
    # invalid rows: prediction = @@UNKNOWN@@ (28), loss = 3.33 (nmn.py:250-269)
    inv = g["semantic.valid"] == 0
    assert (g["semantic.predictions"][inv] == 28).all()
    np.testing.assert_allclose(g["semantic.loss"][inv], 3.33, rtol=1e-6)


def test_vocabulary_roundtrip(tmp_path):
    v = Vocabulary.clevr()
    assert v.get_vocab_size("programs") == 44 and v.get_vocab_size("answers") == 29
    assert v.get_token_index("@@PADDING@@", "programs") == 0 and v.get_token_index("@end@", "questions") == 3
    v.save_to_files(str(tmp_path / "vocab"))
    w = Vocabulary.from_files(str(tmp_path / "vocab"))
    for ns in ("programs", "questions", "answers"):
        assert w.get_token_to_index_vocabulary(ns) == v.get_token_to_index_vocabulary(ns)
