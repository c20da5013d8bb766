"""Minimal accumulators with the AllenNLP 0.9.0 call signatures used by probnmn/models/nmn.py:121-124,262-263,291-294."""
import torch


class Average:
    def __init__(self):
        self._total, self._count = 0.0, 0

    def __call__(self, value):
        self._total += float(value)
        self._count += 1

    def get_metric(self, reset: bool = False):
        avg = self._total / self._count if self._count > 0 else 0.0
        if reset:
            self._total, self._count = 0.0, 0
        return avg


class BooleanAccuracy:
    def __init__(self):
        self._correct, self._total = 0.0, 0.0

    def __call__(self, predictions: torch.Tensor, gold_labels: torch.Tensor, mask=None):
        predictions, gold_labels = predictions.detach().cpu(), gold_labels.detach().cpu()
        batch = predictions.size(0)
        correct = predictions.view(batch, -1).eq(gold_labels.view(batch, -1)).prod(dim=1).float()
        self._correct += correct.sum().item()
        self._total += batch

    def get_metric(self, reset: bool = False):
        acc = self._correct / self._total if self._total > 0 else 0.0
        if reset:
            self._correct, self._total = 0.0, 0.0
        return acc
