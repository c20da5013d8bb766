"""CPU checks of the seq2seq oracle (oracle/seq2seq_oracle.py): against golden vectors recorded from the reference's OWN
seq2seq_base.py / program_prior.py run verbatim over a shim of the absent AllenNLP 0.9.0 (second half of this file),
against the torch modules AllenNLP wraps (nn.LSTM over a packed sequence inside PytorchSeq2SeqWrapper, nn.LSTMCell),
and on hand-checked cases of the in-repo logic (seq2seq_base.py:203-293)."""
import torch
from torch import nn

from oracle import seq2seq_oracle as O
from probnmn_clevr_b200.synthetic import make_questions, make_seq2seq_state_dict


def test_boundary_tokens():
    t = torch.tensor([[5, 6, 7, 0, 0], [9, 0, 0, 0, 0], [0, 0, 0, 0, 0], [4, 5, 6, 7, 8]])
    out = O.add_sentence_boundary_token_ids(t)
    assert out.tolist() == [[2, 5, 6, 7, 3, 0, 0], [2, 9, 3, 0, 0, 0, 0], [2, 3, 0, 0, 0, 0, 0], [2, 4, 5, 6, 7, 8, 3]]


def test_trim_predictions():
    p = torch.tensor([[5, 6, 3, 7, 8], [3, 5, 6, 7, 8], [5, 6, 7, 8, 9], [5, 3, 3, 3, 3]])
    assert O.trim_predictions(p).tolist() == [[5, 6, 3, 0, 0], [0, 0, 0, 0, 0], [5, 6, 7, 8, 9], [5, 3, 0, 0, 0]]


def test_encoder_equals_packed_nn_lstm():
    sd = make_seq2seq_state_dict(seed=1)
    q = make_questions(9, 93, seed=2, max_length=17, min_length=1)
    source = O.add_sentence_boundary_token_ids(q)[:, 1:]
    enc, mask = O.encode(sd, source)
    lstm = nn.LSTM(256, 256, 2, batch_first=True)
    lstm.load_state_dict({k.replace("_encoder._module.", ""): v for k, v in sd.items() if k.startswith("_encoder._module.")})
    x = nn.functional.embedding(source, sd["_source_embedder.token_embedder_tokens.weight"])
    lengths = mask.sum(1)
    packed = nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True, enforce_sorted=False)
    with torch.no_grad():
        y, _ = lstm(packed)
    y, _ = nn.utils.rnn.pad_packed_sequence(y, batch_first=True, total_length=source.shape[1])
    assert float((enc - y).abs().max()) < 1e-5


def test_decoder_cell_equals_nn_lstmcell():
    sd = make_seq2seq_state_dict(seed=3)
    cell = nn.LSTMCell(512, 256)
    cell.load_state_dict({k.replace("_decoder_cell.", ""): v for k, v in sd.items() if k.startswith("_decoder_cell.")})
    g = torch.Generator().manual_seed(0)
    x, h, c = torch.randn(6, 512, generator=g), torch.randn(6, 256, generator=g), torch.randn(6, 256, generator=g)
    with torch.no_grad():
        h1, c1 = cell(x, (h, c))
    h2, c2 = O.lstm_cell(x, h, c, sd["_decoder_cell.weight_ih"], sd["_decoder_cell.weight_hh"], sd["_decoder_cell.bias_ih"],
                         sd["_decoder_cell.bias_hh"])
    assert float((h1 - h2).abs().max()) < 1e-6 and float((c1 - c2).abs().max()) < 1e-6


def test_losses_on_a_hand_checked_case():
    sd = make_seq2seq_state_dict(seed=4)
    q = make_questions(4, 93, seed=5, max_length=9)
    p = torch.tensor([[10, 11, 12, 0], [13, 0, 0, 0], [14, 15, 16, 17], [0, 0, 0, 0]])
    out = O.seq2seq_forward(sd, q, p, decoding_strategy="greedy")
    assert out["logits"].shape == (4, 5, 44)
    logp = torch.log_softmax(out["logits"], -1)
    tgt = O.add_sentence_boundary_token_ids(p)[:, 1:]
    for b in range(4):
        n = int((tgt[b] != 0).sum())
        want = -sum(float(logp[b, t, tgt[b, t]]) for t in range(n)) / n
        assert abs(float(out["loss"][b]) - want) < 1e-5
    free = O.seq2seq_forward(sd, q, None, decoding_strategy="greedy", max_decoding_steps=7)
    assert free["predictions"].shape == (4, 7) and free["loss"].shape == (4,)


# ---- pinned against the reference's own files (tests/golden/seq2seq_golden.npz, oracle/make_seq2seq_golden.py) ---------------
import os  # noqa: E402

import numpy as np  # noqa: E402
import pytest  # noqa: E402

from oracle import prior_oracle  # noqa: E402
from probnmn_clevr_b200.synthetic import make_prior_state_dict  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "seq2seq_golden.npz")
# (case, model kind, weight seed, gain, free-running steps) -- as in oracle/make_seq2seq_golden.py CASES
GOLDEN_CASES = [("pg", "pg", 0, 1.0, 26), ("pg_sharp", "pg", 3, 4.0, 26), ("qr", "qr", 1, 1.0, 45), ("qr_sharp", "qr", 2, 4.0, 45)]


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12)) if a.numel() else 0.0


def _golden_state_dict(kind, seed, gain):
    vs, vt = (93, 44) if kind == "pg" else (44, 93)
    return make_seq2seq_state_dict(vs, vt, seed=seed, gain=gain)


def _check_grads(g, prefix, sd, tol):
    for k, v in sd.items():
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        norm = float(g[f"{prefix}.gradnorm.{k}"])
        assert abs(float(got.double().norm()) - norm) <= tol * max(norm, 1e-12), k
        if f"{prefix}.grad.{k}" in g:
            want = torch.from_numpy(g[f"{prefix}.grad.{k}"])
        else:
            want, got = torch.from_numpy(g[f"{prefix}.gradsub.{k}"]), got.reshape(-1)[::997]
        assert float((got - want).abs().max()) <= tol * (float(want.abs().max()) + 1e-12) + 1e-9, k


@pytest.mark.parametrize("case,kind,seed,gain,steps", GOLDEN_CASES)
def test_oracle_matches_reference_seq2seq_golden(case, kind, seed, gain, steps):
    """The restatement against what the reference's OWN seq2seq_base.py computed (verbatim, over the AllenNLP shim) on
    the same seeded weights and inputs: teacher-forced logits / loss / gradients, free-running greedy tokens / loss,
    and the REINFORCE loss / gradients on the reference's sampled tokens."""
    g = np.load(GOLDEN)
    sd = {k: v.requires_grad_(True) for k, v in _golden_state_dict(kind, seed, gain).items()}
    src, tgt = torch.from_numpy(g[f"{case}.source"]), torch.from_numpy(g[f"{case}.target"])
    w = torch.from_numpy(g[f"{case}.weights"])
    out = O.seq2seq_forward(sd, src, tgt, "greedy")
    assert _rel(out["logits"].detach(), torch.from_numpy(g[f"{case}.tf.logits"])) < 1e-5
    assert _rel(out["loss"].detach(), torch.from_numpy(g[f"{case}.tf.loss"])) < 1e-5
    assert torch.equal(out["predictions"], torch.from_numpy(g[f"{case}.tf.greedy_predictions"]))
    (out["loss"] * w).sum().backward()
    _check_grads(g, f"{case}.tf", sd, 1e-4)
    with torch.no_grad():
        out = O.seq2seq_forward(sd, src, None, "greedy", steps)
    assert torch.equal(out["raw_predictions"], torch.from_numpy(g[f"{case}.greedy.raw_predictions"]))
    assert torch.equal(out["predictions"], torch.from_numpy(g[f"{case}.greedy.predictions"]))
    assert _rel(out["loss"], torch.from_numpy(g[f"{case}.greedy.loss"])) < 1e-5
    assert _rel(out["logits"], torch.from_numpy(g[f"{case}.greedy.logits"])) < 1e-5
    for v in sd.values():
        v.grad = None
    raw = torch.from_numpy(g[f"{case}.sampled.raw_predictions"])
    out = O.seq2seq_forward(sd, src, None, "sampling", steps, forced_choices=raw)
    assert torch.equal(out["predictions"], torch.from_numpy(g[f"{case}.sampled.predictions"]))
    assert _rel(out["loss"].detach(), torch.from_numpy(g[f"{case}.sampled.loss"])) < 1e-5
    (out["loss"] * w).sum().backward()
    _check_grads(g, f"{case}.sampled", sd, 1e-4)


def test_oracle_matches_reference_golden_edge_and_prior():
    g = np.load(GOLDEN)
    sd = make_seq2seq_state_dict(93, 44, seed=4)
    sd["_output_projection_layer.bias"][3] = 50.0
    out = O.seq2seq_forward(sd, torch.from_numpy(g["end_first.source"]), None, "greedy", 26)
    assert torch.equal(out["predictions"], torch.from_numpy(g["end_first.predictions"])) and int(out["predictions"].abs().sum()) == 0
    assert torch.equal(out["loss"], torch.from_numpy(g["end_first.loss"]))
    sdp = make_prior_state_dict(44, hidden=256, seed=0)
    out = prior_oracle.prior_forward(sdp, torch.from_numpy(g["prior.programs"]))
    assert _rel(out["loss"], torch.from_numpy(g["prior.loss"])) < 1e-5
    assert list(out["predictions"].shape) == g["prior.predictions_shape"].tolist()


@pytest.mark.skipif(not os.path.exists("/root/reference/probnmn/modules/seq2seq_base.py"), reason="reference tree absent")
def test_reference_files_run_verbatim_over_the_shim():
    """Build container only: the reference's classes load the synthetic state dict strictly (same parameter names and
    shapes as the drop-ins) and reproduce one golden number live."""
    from oracle.ref_loader import load_reference_seq2seq
    from probnmn_clevr_b200.vocabulary import Vocabulary
    RefPG, RefQR, RefPrior = load_reference_seq2seq()
    g = np.load(GOLDEN)
    ref = RefPG(Vocabulary.clevr())
    ref.load_state_dict(_golden_state_dict("pg", 0, 1.0), strict=True)
    ref.eval()
    with torch.no_grad():
        out = ref(torch.from_numpy(g["pg.source"]), None, decoding_strategy="greedy")
    assert torch.equal(out["predictions"], torch.from_numpy(g["pg.greedy.predictions"]))
    assert _rel(out["loss"], torch.from_numpy(g["pg.greedy.loss"])) < 1e-6
