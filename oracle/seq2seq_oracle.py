"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's ProgramGenerator
(``probnmn/models/program_generator.py`` -> ``probnmn/modules/seq2seq_base.py`` -> AllenNLP ``SimpleSeq2Seq``)
in plain fp32 PyTorch.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file.

PINNING.  Everything that lives in the reference repository -- ``Seq2SeqBase.forward`` / ``_forward_loop`` /
``_trim_predictions`` / ``_get_loss`` (seq2seq_base.py:101-341, cited line by line below) and the ProgramGenerator /
QuestionReconstructor wrappers -- is PINNED: ``oracle/make_seq2seq_golden.py`` imports those files VERBATIM from
/root/reference, runs them on seeded weights / inputs and records ``tests/golden/seq2seq_golden.npz``; this restatement
reproduces those vectors to <= 1e-6 (logits, both losses, greedy tokens, gradients; tests/test_seq2seq_oracle.py).
What those files build on lives in ``allennlp==0.9.0`` (``requirements.txt:1``), which is neither vendored under
``/root/reference`` nor installable here (no network); the golden run supplies it through ``oracle/ref_shim/allennlp``:
torch's own ``nn.LSTM`` on packed sequences and ``nn.LSTMCell`` (so the LSTM arithmetic is pinned to torch), plus ~60
lines restated from AllenNLP 0.9.0's published source, which therefore remain PARITY UNPINNED against AllenNLP itself
(SURVEY.md appendix C):

  * ``add_sentence_boundary_token_ids``: ``[@start@, w_1..w_n, @end@, 0...]``               (nn/util.py)
  * ``_encode``: mask = tokens != 0; embedding (padding_index 0); ``PytorchSeq2SeqWrapper(nn.LSTM)``: packed
    sequence semantics = state frozen and output zero beyond each row's length
  * ``_init_decoder_state``: h0 = encoder output at the last valid position, c0 = 0
  * ``_prepare_output_projections``: e = target_embedder[token]; DotProductAttention + ``masked_softmax``
    (p = softmax(s*m)*m; p /= (sum p + 1e-13)); a = sum_t p_t enc_t; (h,c) = LSTMCell(cat(a, e), (h,c));
    logits = W_o h + b_o
  * ``sequence_cross_entropy_with_logits(average=None)``: per row sum(nll*mask) / (sum(mask) + 1e-13)

State-dict keys are AllenNLP's: ``_source_embedder.token_embedder_tokens.weight``,
``_encoder._module.{weight,bias}_{ih,hh}_l{0,1}``, ``_target_embedder.weight``,
``_decoder_cell.{weight,bias}_{ih,hh}``, ``_output_projection_layer.{weight,bias}``.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

PAD, UNK, START, END = 0, 1, 2, 3  # identical in every padded namespace (seq2seq_base.py:61-65)


def add_sentence_boundary_token_ids(tokens: torch.Tensor) -> torch.Tensor:
    """(B,T) 0-padded -> (B,T+2): @start@ first, @end@ right after the last real token."""
    B, T = tokens.shape
    lengths = (tokens != PAD).sum(1)
    out = tokens.new_zeros(B, T + 2)
    out[:, 1:-1] = tokens
    out[:, 0] = START
    out[torch.arange(B), lengths + 1] = END
    return out


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=-1)  # PyTorch gate order
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def encode(sd: Dict[str, torch.Tensor], source: torch.Tensor, num_layers: int = 2):
    """source: (B,T) tokens (already boundary-processed).  Returns encoder outputs (B,T,H) and mask (B,T)."""
    mask = source != PAD
    x = F.embedding(source, sd["_source_embedder.token_embedder_tokens.weight"])
    B, T, _ = x.shape
    for layer in range(num_layers):
        w_ih, w_hh = sd[f"_encoder._module.weight_ih_l{layer}"], sd[f"_encoder._module.weight_hh_l{layer}"]
        b_ih, b_hh = sd[f"_encoder._module.bias_ih_l{layer}"], sd[f"_encoder._module.bias_hh_l{layer}"]
        H = w_hh.shape[1]
        h, c = x.new_zeros(B, H), x.new_zeros(B, H)
        outs = []
        for t in range(T):
            h2, c2 = lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            m = mask[:, t].unsqueeze(1).to(x.dtype)
            h, c = m * h2 + (1 - m) * h, m * c2 + (1 - m) * c  # packed sequence: frozen beyond the length
            outs.append(m * h2)                                 # ... and zero output there
        x = torch.stack(outs, 1)
    return x, mask


def decoder_step(sd, tokens, h, c, enc, mask):
    """AllenNLP ``_prepare_output_projections`` (called at seq2seq_base.py:201)."""
    e = F.embedding(tokens, sd["_target_embedder.weight"])
    scores = torch.bmm(enc, h.unsqueeze(2)).squeeze(2)  # DotProductAttention
    m = mask.to(enc.dtype)
    p = F.softmax(scores * m, dim=-1) * m               # masked_softmax (0.9.0)
    p = p / (p.sum(-1, keepdim=True) + 1e-13)
    attended = torch.bmm(p.unsqueeze(1), enc).squeeze(1)
    h, c = lstm_cell(torch.cat([attended, e], -1), h, c, sd["_decoder_cell.weight_ih"], sd["_decoder_cell.weight_hh"],
                     sd["_decoder_cell.bias_ih"], sd["_decoder_cell.bias_hh"])
    logits = F.linear(h, sd["_output_projection_layer.weight"], sd["_output_projection_layer.bias"])
    return logits, h, c


def trim_predictions(pred: torch.Tensor) -> torch.Tensor:
    """seq2seq_base.py:278-293: keep up to and including the first @end@, zero the rest; a row whose FIRST token
    is @end@ becomes all zeros; a row without @end@ is unchanged."""
    out = torch.zeros_like(pred)
    for i, row in enumerate(pred.tolist()):
        if END in row:
            k = row.index(END)
            if k > 0:
                out[i, : k + 1] = pred[i, : k + 1]
        else:
            out[i] = pred[i]
    return out


def seq2seq_forward(sd: Dict[str, torch.Tensor], source_tokens: torch.Tensor, target_tokens: Optional[torch.Tensor] = None,
                    decoding_strategy: str = "sampling", max_decoding_steps: int = 26,
                    generator: Optional[torch.Generator] = None, forced_choices: Optional[torch.Tensor] = None):
    """``Seq2SeqBase.forward`` + ``_forward_loop`` (seq2seq_base.py:101-276) without the metric objects.

    ``forced_choices`` (B,steps) replaces the multinomial draw (used to replay the CUDA path's samples so that
    log-probabilities and gradients can be compared; the two RNG streams cannot be matched)."""
    source = add_sentence_boundary_token_ids(source_tokens)[:, 1:]  # :128-139 (leading @start@ dropped)
    targets = add_sentence_boundary_token_ids(target_tokens) if target_tokens is not None else None
    enc, mask = encode(sd, source)
    B = source.shape[0]
    last = mask.sum(1) - 1                                           # _init_decoder_state
    h = enc[torch.arange(B), last]
    c = torch.zeros_like(h)
    steps = targets.shape[1] - 1 if targets is not None else max_decoding_steps  # :168-177
    last_pred = source.new_full((B,), START)
    step_logits, step_logprobs, step_preds = [], [], []
    for t in range(steps):
        inp = targets[:, t] if targets is not None else last_pred  # scheduled sampling ratio 0 (:188-198)
        logits, h, c = decoder_step(sd, inp, h, c, enc, mask)
        probs = F.softmax(logits, dim=-1)
        logprobs = F.log_softmax(logits, dim=-1)
        if decoding_strategy == "greedy":
            pred = probs.max(1)[1]                                   # :208-209
        else:
            p = probs.detach().clone()
            p[:, PAD] = 0; p[:, UNK] = 0; p[:, START] = 0            # :212-214
            pred = forced_choices[:, t] if forced_choices is not None else torch.multinomial(p, 1, generator=generator).squeeze(1)
        last_pred = pred
        step_preds.append(pred)
        step_logits.append(logits)
        step_logprobs.append(logprobs[torch.arange(B), pred])        # :220
    raw = torch.stack(step_preds, 1)
    predictions = trim_predictions(raw)                              # :230
    logprobs = torch.stack(step_logprobs, 1)
    pmask = (predictions != PAD).to(logprobs.dtype)
    seq_lp = (logprobs * pmask).sum(-1) / (pmask.sum(-1) + 1e-12)    # :235-244
    out = {"predictions": predictions, "loss": -seq_lp, "raw_predictions": raw, "logits": torch.stack(step_logits, 1)}
    if targets is not None:                                          # :247-254 + _get_loss :333-341
        rel_t, rel_m = targets[:, 1:], (targets != PAD)[:, 1:].to(logprobs.dtype)
        nll = -F.log_softmax(out["logits"], -1).gather(2, rel_t.unsqueeze(2)).squeeze(2)
        out["loss"] = (nll * rel_m).sum(1) / (rel_m.sum(1) + 1e-13)
    return out
