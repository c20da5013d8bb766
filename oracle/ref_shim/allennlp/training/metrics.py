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


# ---- sequence metrics used by probnmn/modules/seq2seq_base.py:95-99,256-274 and probnmn/utils/metrics.py:9 (0.9.0) ----
import math  # noqa: E402
import sys  # noqa: E402
from collections import Counter  # noqa: E402


class SequenceAccuracy:
    def __init__(self) -> None:
        self.correct_count, self.total_count = 0.0, 0.0

    def __call__(self, predictions: torch.Tensor, gold_labels: torch.Tensor, mask=None):
        predictions, gold_labels = predictions.detach().cpu(), gold_labels.detach().cpu()
        mask = mask.detach().cpu() if mask is not None else None
        k = predictions.size()[1]
        expanded_size = list(gold_labels.size())
        expanded_size.insert(1, k)
        expanded_gold = gold_labels.unsqueeze(1).expand(expanded_size)
        if mask is not None:
            expanded_mask = mask.unsqueeze(1).expand(expanded_size)
            masked_gold = expanded_mask * expanded_gold
            masked_predictions = expanded_mask * predictions
        else:
            masked_gold, masked_predictions = expanded_gold, predictions
        eqs = masked_gold.eq(masked_predictions)
        matches_per_question = eqs.min(dim=2)[0]
        some_match = matches_per_question.max(dim=1)[0]
        self.total_count += predictions.size()[0]
        self.correct_count += some_match.sum().item()

    def get_metric(self, reset: bool = False):
        accuracy = self.correct_count / self.total_count if self.total_count > 0 else 0
        if reset:
            self.reset()
        return accuracy

    def reset(self):
        self.correct_count, self.total_count = 0.0, 0.0


class UnigramRecall:
    def __init__(self) -> None:
        self.correct_count, self.total_count = 0.0, 0.0

    def __call__(self, predictions: torch.Tensor, gold_labels: torch.Tensor, mask=None, end_index: int = sys.maxsize):
        predictions, gold_labels = predictions.detach().cpu(), gold_labels.detach().cpu()
        mask = mask.detach().cpu() if mask is not None else None
        correct = 0.0
        for i in range(predictions.size()[0]):
            beams = predictions[i]
            cur_gold = gold_labels[i]
            masked_gold = cur_gold * mask[i] if mask is not None else cur_gold
            cleaned_gold = [x for x in masked_gold if x != 0 and x != end_index]
            retval = 0.0
            for word in cleaned_gold:
                stillsearch = True
                for beam in beams:
                    if stillsearch and (word in beam):
                        retval += 1.0 / float(len(cleaned_gold))
                        stillsearch = False
            correct += retval
        self.correct_count += correct
        self.total_count += predictions.size()[0]

    def get_metric(self, reset: bool = False):
        recall = self.correct_count / self.total_count if self.total_count > 0 else 0
        if reset:
            self.reset()
        return recall

    def reset(self):
        self.correct_count, self.total_count = 0.0, 0.0


class BLEU:
    def __init__(self, ngram_weights=(0.25, 0.25, 0.25, 0.25), exclude_indices=None) -> None:
        self._ngram_weights = ngram_weights
        self._exclude_indices = exclude_indices or set()
        self.reset()

    def reset(self):
        self._precision_matches, self._precision_totals = Counter(), Counter()
        self._prediction_lengths, self._reference_lengths = 0, 0

    def _ngrams(self, tensor: torch.Tensor, ngram_size: int):
        ngram_counts = Counter()
        if ngram_size > tensor.size(-1):
            return ngram_counts
        for start_position in range(ngram_size):
            for tensor_slice in tensor[start_position:].split(ngram_size, dim=-1):
                if tensor_slice.size(-1) < ngram_size:
                    break
                ngram = tuple(x.item() for x in tensor_slice)
                if any(x in self._exclude_indices for x in ngram):
                    continue
                ngram_counts[ngram] += 1
        return ngram_counts

    def _get_modified_precision_counts(self, predicted_tokens, reference_tokens, ngram_size):
        clipped_matches, total_predicted = 0, 0
        for batch_num in range(predicted_tokens.size(0)):
            predicted_ngram_counts = self._ngrams(predicted_tokens[batch_num, :], ngram_size)
            reference_ngram_counts = self._ngrams(reference_tokens[batch_num, :], ngram_size)
            for ngram, count in predicted_ngram_counts.items():
                clipped_matches += min(count, reference_ngram_counts[ngram])
                total_predicted += count
        return clipped_matches, total_predicted

    def _get_valid_tokens_mask(self, tensor: torch.Tensor):
        valid = torch.ones(tensor.size(), dtype=torch.bool)
        for index in self._exclude_indices:
            valid = valid & (tensor != index)
        return valid

    def _get_brevity_penalty(self) -> float:
        if self._prediction_lengths > self._reference_lengths:
            return 1.0
        if self._reference_lengths == 0 or self._prediction_lengths == 0:
            return 0.0
        return math.exp(1.0 - self._reference_lengths / self._prediction_lengths)

    def __call__(self, predictions: torch.Tensor, gold_targets: torch.Tensor) -> None:
        predictions, gold_targets = predictions.detach().cpu(), gold_targets.detach().cpu()
        for ngram_size, _ in enumerate(self._ngram_weights, start=1):
            m, t = self._get_modified_precision_counts(predictions, gold_targets, ngram_size)
            self._precision_matches[ngram_size] += m
            self._precision_totals[ngram_size] += t
        if not self._exclude_indices:
            self._prediction_lengths += predictions.size(0) * predictions.size(1)
            self._reference_lengths += gold_targets.size(0) * gold_targets.size(1)
        else:
            self._prediction_lengths += self._get_valid_tokens_mask(predictions).sum().item()
            self._reference_lengths += self._get_valid_tokens_mask(gold_targets).sum().item()

    def get_metric(self, reset: bool = False):
        brevity_penalty = self._get_brevity_penalty()
        ngram_scores = (weight * (math.log(self._precision_matches[n] + 1e-13) - math.log(self._precision_totals[n] + 1e-13))
                        for n, weight in enumerate(self._ngram_weights, start=1))
        bleu = brevity_penalty * math.exp(sum(ngram_scores))
        if reset:
            self.reset()
        return {"BLEU": bleu}
