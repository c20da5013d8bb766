"""ORACLE tooling -- ``allennlp.models.encoder_decoders.SimpleSeq2Seq`` as far as the reference uses it
(probnmn/modules/seq2seq_base.py:16,86-94,141-146,201).

AllenNLP 0.9.0 is the reference's pinned dependency (requirements.txt:1) and is neither vendored nor installable here.
This class restates, from the published 0.9.0 source (allennlp/models/encoder_decoders/simple_seq2seq.py), exactly the
members the reference's OWN ``Seq2SeqBase`` relies on, so that ``seq2seq_base.py`` can run VERBATIM for golden-vector
generation (oracle/make_seq2seq_golden.py):

  * ``__init__``: ``_start_index`` / ``_end_index`` from the target namespace, ``_max_decoding_steps``,
    ``_scheduled_sampling_ratio`` (0), ``_source_embedder``, ``_encoder``, ``_attention``, ``_target_embedder``
    (``Embedding(num_classes, source_embedder.get_output_dim())``), ``_decoder_cell`` (``LSTMCell(enc_dim + emb_dim,
    enc_dim)`` with attention), ``_output_projection_layer`` (``Linear(enc_dim, num_classes)``), ``_bleu``
  * ``_encode``, ``_init_decoder_state``, ``_prepare_output_projections``, ``_prepare_attended_input``

Beam search, ``forward``, ``_forward_loop`` and ``decode`` of the original are not restated: the reference overrides or
never calls them.  Parameter names (hence state-dict keys) are AllenNLP's.
"""
from typing import Dict, Tuple

import torch
from torch.nn.modules.linear import Linear
from torch.nn.modules.rnn import LSTMCell

from allennlp.modules.token_embedders import Embedding
from allennlp.nn import util
from allennlp.training.metrics import BLEU

START_SYMBOL, END_SYMBOL = "@start@", "@end@"


class SimpleSeq2Seq(torch.nn.Module):
    def __init__(self, vocab, source_embedder, encoder, max_decoding_steps: int, attention=None, attention_function=None,
                 beam_size: int = None, target_namespace: str = "tokens", target_embedding_dim: int = None,
                 scheduled_sampling_ratio: float = 0.0, use_bleu: bool = True) -> None:
        super().__init__()
        self.vocab = vocab
        self._target_namespace = target_namespace
        self._scheduled_sampling_ratio = scheduled_sampling_ratio
        self._start_index = self.vocab.get_token_index(START_SYMBOL, self._target_namespace)
        self._end_index = self.vocab.get_token_index(END_SYMBOL, self._target_namespace)
        if use_bleu:
            pad_index = self.vocab.get_token_index("@@PADDING@@", self._target_namespace)
            self._bleu = BLEU(exclude_indices={pad_index, self._end_index, self._start_index})
        else:
            self._bleu = None
        self._max_decoding_steps = max_decoding_steps
        self._source_embedder = source_embedder
        self._encoder = encoder
        num_classes = self.vocab.get_vocab_size(self._target_namespace)
        if attention is not None and attention_function is not None:
            raise ValueError("only one of attention / attention_function")
        self._attention = attention
        target_embedding_dim = target_embedding_dim or source_embedder.get_output_dim()
        self._target_embedder = Embedding(num_classes, target_embedding_dim)
        self._encoder_output_dim = self._encoder.get_output_dim()
        self._decoder_output_dim = self._encoder_output_dim
        if self._attention:
            self._decoder_input_dim = self._decoder_output_dim + target_embedding_dim
        else:
            self._decoder_input_dim = target_embedding_dim
        self._decoder_cell = LSTMCell(self._decoder_input_dim, self._decoder_output_dim)
        self._output_projection_layer = Linear(self._decoder_output_dim, num_classes)

    def _encode(self, source_tokens: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        embedded_input = self._source_embedder(source_tokens)
        source_mask = util.get_text_field_mask(source_tokens)
        encoder_outputs = self._encoder(embedded_input, source_mask)
        return {"source_mask": source_mask, "encoder_outputs": encoder_outputs}

    def _init_decoder_state(self, state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        batch_size = state["source_mask"].size(0)
        final_encoder_output = util.get_final_encoder_states(state["encoder_outputs"], state["source_mask"],
                                                             self._encoder.is_bidirectional())
        state["decoder_hidden"] = final_encoder_output
        state["decoder_context"] = state["encoder_outputs"].new_zeros(batch_size, self._decoder_output_dim)
        return state

    def _prepare_output_projections(self, last_predictions: torch.Tensor,
                                    state: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        encoder_outputs = state["encoder_outputs"]
        source_mask = state["source_mask"]
        decoder_hidden = state["decoder_hidden"]
        decoder_context = state["decoder_context"]
        embedded_input = self._target_embedder(last_predictions)
        if self._attention:
            attended_input = self._prepare_attended_input(decoder_hidden, encoder_outputs, source_mask)
            decoder_input = torch.cat((attended_input, embedded_input), -1)
        else:
            decoder_input = embedded_input
        decoder_hidden, decoder_context = self._decoder_cell(decoder_input, (decoder_hidden, decoder_context))
        state["decoder_hidden"] = decoder_hidden
        state["decoder_context"] = decoder_context
        output_projections = self._output_projection_layer(decoder_hidden)
        return output_projections, state

    def _prepare_attended_input(self, decoder_hidden_state, encoder_outputs, encoder_outputs_mask) -> torch.Tensor:
        encoder_outputs_mask = encoder_outputs_mask.float()
        input_weights = self._attention(decoder_hidden_state, encoder_outputs, encoder_outputs_mask)
        attended_input = util.weighted_sum(encoder_outputs, input_weights)
        return attended_input
