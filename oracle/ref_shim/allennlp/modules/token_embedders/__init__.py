"""ORACLE tooling -- ``allennlp.modules.token_embedders.Embedding`` (0.9.0): a ``weight`` parameter
(xavier-uniform, the ``padding_index`` row zeroed) looked up with ``torch.nn.functional.embedding``."""
import torch
from torch.nn.functional import embedding


class Embedding(torch.nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, projection_dim: int = None, weight: torch.Tensor = None,
                 padding_index: int = None, trainable: bool = True, max_norm: float = None, norm_type: float = 2.0,
                 scale_grad_by_freq: bool = False, sparse: bool = False) -> None:
        super().__init__()
        if projection_dim is not None:
            raise NotImplementedError("projection is not used by the reference")
        self.num_embeddings, self.padding_index = num_embeddings, padding_index
        self.max_norm, self.norm_type, self.scale_grad_by_freq, self.sparse = max_norm, norm_type, scale_grad_by_freq, sparse
        self.output_dim = embedding_dim
        if weight is None:
            weight = torch.FloatTensor(num_embeddings, embedding_dim)
            self.weight = torch.nn.Parameter(weight, requires_grad=trainable)
            torch.nn.init.xavier_uniform_(self.weight)
        else:
            self.weight = torch.nn.Parameter(weight, requires_grad=trainable)
        if self.padding_index is not None:
            self.weight.data[self.padding_index].fill_(0)

    def get_output_dim(self) -> int:
        return self.output_dim

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return embedding(inputs, self.weight, padding_idx=self.padding_index, max_norm=self.max_norm,
                         norm_type=self.norm_type, scale_grad_by_freq=self.scale_grad_by_freq, sparse=self.sparse)
