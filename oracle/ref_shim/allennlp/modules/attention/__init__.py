"""ORACLE tooling -- ``allennlp.modules.attention.DotProductAttention`` (0.9.0: ``Attention.forward`` computes the
similarities and, with ``normalize=True`` (the default), applies ``masked_softmax``)."""
import torch

from allennlp.nn.util import masked_softmax


class DotProductAttention(torch.nn.Module):
    def __init__(self, normalize: bool = True) -> None:
        super().__init__()
        self._normalize = normalize

    def forward(self, vector: torch.Tensor, matrix: torch.Tensor, matrix_mask: torch.Tensor = None) -> torch.Tensor:
        similarities = matrix.bmm(vector.unsqueeze(-1)).squeeze(-1)
        if self._normalize:
            return masked_softmax(similarities, matrix_mask)
        return similarities
