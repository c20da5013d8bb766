"""CPU checks of the seq2seq oracle (oracle/seq2seq_oracle.py).  AllenNLP 0.9.0 is absent, so the oracle's restatement
is PARITY UNPINNED against the reference itself; what CAN be pinned here is that its LSTM arithmetic equals the
torch modules AllenNLP wraps (nn.LSTM over a packed sequence inside PytorchSeq2SeqWrapper, nn.LSTMCell) and that the
in-repo logic (seq2seq_base.py:203-293) is restated faithfully on hand-checked cases."""
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
