"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's ``ProgramPrior.forward``
(``probnmn/models/program_prior.py:80-155``) in plain fp32 PyTorch.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file.

PINNED against the reference's own ``program_prior.py``, imported verbatim by ``oracle/make_seq2seq_golden.py`` and run
over ``oracle/ref_shim/allennlp`` (``tests/golden/seq2seq_golden.npz``, key ``prior.*``; tests/test_seq2seq_oracle.py).
The shim wraps torch's own ``nn.LSTM`` on packed sequences; its boundary-token / sequence-cross-entropy helpers are
restated from AllenNLP 0.9.0's published source (``requirements.txt:1``, absent here) and stay unpinned against AllenNLP
itself.  The in-repo logic is cited line by line.

State-dict keys (AllenNLP names): ``_embedder.token_embedder_programs.weight`` (V,256, row 0 = padding),
``_encoder._module.{weight,bias}_{ih,hh}_l{0,1}``, ``_projection_layer.weight`` (256,256), ``_output_layer.weight``
(tied to the embedding, program_prior.py:59-62).
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .seq2seq_oracle import END, PAD, START, UNK, add_sentence_boundary_token_ids, lstm_cell


def prior_forward(sd: Dict[str, torch.Tensor], program_tokens: torch.Tensor, generator: Optional[torch.Generator] = None):
    tokens = add_sentence_boundary_token_ids(program_tokens)            # :104-107
    mask = tokens != PAD                                                 # :108
    emb_w = sd["_embedder.token_embedder_programs.weight"]
    x = F.embedding(tokens, emb_w)                                       # :113
    B, T, _ = x.shape
    for layer in range(2):                                               # :116 PytorchSeq2SeqWrapper(nn.LSTM): packed semantics
        w_ih, w_hh = sd[f"_encoder._module.weight_ih_l{layer}"], sd[f"_encoder._module.weight_hh_l{layer}"]
        b_ih, b_hh = sd[f"_encoder._module.bias_ih_l{layer}"], sd[f"_encoder._module.bias_hh_l{layer}"]
        h, c = x.new_zeros(B, w_hh.shape[1]), x.new_zeros(B, w_hh.shape[1])
        outs = []
        for t in range(T):
            h2, c2 = lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            m = mask[:, t].unsqueeze(1).to(x.dtype)
            h, c = m * h2 + (1 - m) * h, m * c2 + (1 - m) * c
            outs.append(m * h2)
        x = torch.stack(outs, 1)
    proj = F.linear(x, sd["_projection_layer.weight"])                   # :119
    logits = F.linear(proj, sd.get("_output_layer.weight", emb_w))       # :121 (tied weight)
    probs = F.softmax(logits, dim=-1).detach().clone()                   # :123-127
    probs[:, :, START] = 0; probs[:, :, PAD] = 0; probs[:, :, UNK] = 0
    preds = torch.stack([torch.multinomial(probs[b], 1, generator=generator).squeeze(1) for b in range(B)], 0)  # :129-137
    predictions = preds[:, :-1] * mask[:, 1:].long()                     # :139
    rel_t, rel_m = tokens[:, 1:], mask[:, 1:].to(logits.dtype)           # :142-147
    nll = -F.log_softmax(logits[:, :-1], -1).gather(2, rel_t.unsqueeze(2)).squeeze(2)
    loss = (nll * rel_m).sum(1) / (rel_m.sum(1) + 1e-13)
    return {"predictions": predictions, "loss": loss, "logits": logits[:, :-1]}
