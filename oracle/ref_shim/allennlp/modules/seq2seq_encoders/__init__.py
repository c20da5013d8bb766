"""ORACLE tooling -- ``allennlp.modules.seq2seq_encoders.PytorchSeq2SeqWrapper`` (0.9.0, non-stateful use):
rows are sorted by length, zero-length rows dropped, the rest packed (``pack_padded_sequence``) and run through the
wrapped ``torch.nn.LSTM`` itself; the output is unpacked, padded back to the input's time length with zeros and restored
to the original row order.  Parameter prefix ``_module`` as in AllenNLP."""
import torch
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence


class PytorchSeq2SeqWrapper(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, stateful: bool = False) -> None:
        super().__init__()
        if stateful:
            raise NotImplementedError("stateful encoders are not used by the reference")
        self._module = module
        if not getattr(self._module, "batch_first", True):
            raise ValueError("Our encoder semantics assumes batch is always first!")
        self._is_bidirectional = bool(getattr(self._module, "bidirectional", False))
        self._num_directions = 2 if self._is_bidirectional else 1

    def get_input_dim(self) -> int:
        return self._module.input_size

    def get_output_dim(self) -> int:
        return self._module.hidden_size * self._num_directions

    def is_bidirectional(self) -> bool:
        return self._is_bidirectional

    def forward(self, inputs: torch.Tensor, mask: torch.Tensor, hidden_state: torch.Tensor = None) -> torch.Tensor:
        if mask is None:
            return self._module(inputs, hidden_state)[0]
        batch_size, total_sequence_length = mask.size()
        sequence_lengths = mask.long().sum(-1)
        num_valid = int(torch.sum(mask[:, 0]).item())
        sorted_lengths, permutation = sequence_lengths.sort(0, descending=True)
        sorted_inputs = inputs.index_select(0, permutation)
        _, restoration_indices = permutation.sort(0, descending=False)
        packed = pack_padded_sequence(sorted_inputs[:num_valid], sorted_lengths[:num_valid].tolist(), batch_first=True)
        packed_output, _ = self._module(packed, hidden_state)
        unpacked, _ = pad_packed_sequence(packed_output, batch_first=True)
        if num_valid < batch_size:
            _, length, output_dim = unpacked.size()
            zeros = unpacked.new_zeros(batch_size - num_valid, length, output_dim)
            unpacked = torch.cat([unpacked, zeros], 0)
        difference = total_sequence_length - unpacked.size(1)
        if difference > 0:
            zeros = unpacked.new_zeros(batch_size, difference, unpacked.size(-1))
            unpacked = torch.cat([unpacked, zeros], 1)
        return unpacked.index_select(0, restoration_indices)
