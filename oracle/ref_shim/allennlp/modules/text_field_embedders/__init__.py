"""ORACLE tooling -- ``allennlp.modules.text_field_embedders.BasicTextFieldEmbedder`` (0.9.0): one token embedder per
key of the text field, registered as ``token_embedder_<key>``, outputs concatenated in sorted key order."""
from typing import Dict

import torch


class BasicTextFieldEmbedder(torch.nn.Module):
    def __init__(self, token_embedders: Dict[str, torch.nn.Module]) -> None:
        super().__init__()
        self._token_embedders = token_embedders
        for key, embedder in token_embedders.items():
            self.add_module("token_embedder_%s" % key, embedder)

    def get_output_dim(self) -> int:
        return sum(e.get_output_dim() for e in self._token_embedders.values())

    def forward(self, text_field_input: Dict[str, torch.Tensor], num_wrapping_dims: int = 0) -> torch.Tensor:
        if sorted(self._token_embedders.keys()) != sorted(text_field_input.keys()):
            raise ValueError("Mismatched token keys: %s and %s" % (self._token_embedders.keys(), text_field_input.keys()))
        outs = [getattr(self, "token_embedder_%s" % key)(text_field_input[key]) for key in sorted(self._token_embedders)]
        return torch.cat(outs, dim=-1)
