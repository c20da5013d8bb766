"""Seeded synthetic inputs and weights of the benchmark / parity workloads.

Everything is generated with ``numpy.random.default_rng(seed)`` (PCG64) so that the same arrays can be
regenerated on the GPU box from a seed instead of shipping 257 MB of weights in fixtures.
Formats follow the reference's data pipeline: features are post-ReLU ResNet-101 stage-3 maps
``(B,1024,14,14)`` fp32 (scripts/preprocess/extract_features.py:99-105,131; probnmn/data/datasets.py:140),
programs are prefix-order token ids, 0-padded (scripts/preprocess/preprocess_questions.py:51-74),
answers are ids in [0,28) (build_vocabulary.py:124-125).
"""
from collections import OrderedDict
from typing import List, Optional, Tuple

import numpy as np
import torch

from .nmn import module_class_of
from .vocabulary import SPECIAL_TOKENS, Vocabulary


def module_param_shapes(cls: str, dim: int = 128) -> List[Tuple[str, Tuple[int, ...]]]:
    """(state-dict suffix, shape) of one module, in nn.Module registration order
    (probnmn/modules/nmn_modules.py:71-79,110-116,144-157,194-197,231-237)."""
    c3 = lambda n: [(f"{n}.weight", (dim, dim, 3, 3)), (f"{n}.bias", (dim,))]
    if cls == "attention":
        return c3("conv1") + c3("conv2") + [("conv3.weight", (1, dim, 1, 1)), ("conv3.bias", (1,))]
    if cls == "query":
        return c3("conv1") + c3("conv2")
    if cls == "relate":
        out = []
        for i in range(1, 6):
            out += c3(f"conv{i}")
        return out + [("conv6.weight", (1, dim, 1, 1)), ("conv6.bias", (1,))]
    if cls == "same":
        return [("conv.weight", (1, dim + 1, 1, 1)), ("conv.bias", (1,))]
    if cls == "comparison":
        return [("projection.weight", (dim, 2 * dim, 1, 1)), ("projection.bias", (dim,))] + c3("conv1") + c3("conv2")
    return []


def nmn_param_shapes(vocabulary, image_feature_size=(1024, 14, 14), module_channels=128,
                     class_projection_channels=1024, classifier_linear_size=1024):
    """Ordered (name, shape) list with the reference's state-dict names."""
    cin, h, w = image_feature_size
    n_ans = len(vocabulary.get_index_to_token_vocabulary(namespace="answers")) - 1
    d = module_channels
    shapes = [
        ("stem.0.weight", (d, cin, 3, 3)), ("stem.0.bias", (d,)),
        ("stem.2.weight", (d, d, 3, 3)), ("stem.2.bias", (d,)),
        ("classifier.0.weight", (class_projection_channels, d, 1, 1)), ("classifier.0.bias", (class_projection_channels,)),
        ("classifier.4.weight", (classifier_linear_size, class_projection_channels * h * w // 4)),
        ("classifier.4.bias", (classifier_linear_size,)),
        ("classifier.6.weight", (n_ans, classifier_linear_size)), ("classifier.6.bias", (n_ans,)),
    ]
    for token in vocabulary.get_token_to_index_vocabulary("programs"):
        cls = module_class_of(token)
        if cls in (None, "scene", "and", "or"):
            continue
        shapes += [(f"{token}.{s}", shp) for s, shp in module_param_shapes(cls, d)]
    return shapes


def make_nmn_state_dict(vocabulary, seed: int = 0, **kw) -> "OrderedDict[str, torch.Tensor]":
    """He-normal weights (std = sqrt(2/fan_in)), uniform(+-1/sqrt(fan_in)) biases, fp32."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    last_fan_in = 1
    for name, shape in nmn_param_shapes(vocabulary, **kw):
        if name.endswith("weight"):
            fan_in = int(np.prod(shape[1:]))
            last_fan_in = fan_in
            arr = rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(2.0 / fan_in))
        else:
            b = 1.0 / np.sqrt(last_fan_in)
            arr = rng.uniform(-b, b, size=shape).astype(np.float32)
        sd[name] = torch.from_numpy(arr)
    return sd


def make_features(batch: int, seed: int = 0, channels: int = 1024) -> torch.Tensor:
    rng = np.random.default_rng(1000 + seed)
    x = rng.standard_normal((batch, channels, 14, 14), dtype=np.float32)
    return torch.from_numpy(np.maximum(x, 0) * np.float32(0.5))


def make_answers(batch: int, seed: int = 0, num_answers: int = 28) -> torch.Tensor:
    rng = np.random.default_rng(2000 + seed)
    return torch.from_numpy(rng.integers(0, num_answers, size=batch, dtype=np.int64))


# ------------------------------------------------------------------------------------------------
# programs
# ------------------------------------------------------------------------------------------------
_ATTRS = {
    "color": ["blue", "brown", "cyan", "gray", "green", "purple", "red", "yellow"],
    "material": ["metal", "rubber"],
    "shape": ["cube", "cylinder", "sphere"],
    "size": ["large", "small"],
}
_RELS = ["behind", "front", "left", "right"]


class ProgramSampler:
    """Template grammar over the CLEVR program vocabulary (prefix order, like the reference's H5 files)."""

    def __init__(self, vocabulary, seed: int = 0):
        self.vocab = vocabulary
        self.rng = np.random.default_rng(3000 + seed)

    def _filters(self, n: int) -> List[str]:
        attrs = list(self.rng.permutation(list(_ATTRS.keys()))[:n])
        return [f"filter_{a}[{self.rng.choice(_ATTRS[a])}]" for a in attrs]

    def _chain(self) -> List[str]:
        """application order (innermost first): scene, filters, [unique, relate|same, filters]*"""
        toks = ["scene"] + self._filters(int(self.rng.integers(1, 4)))
        for _ in range(int(self.rng.choice([0, 0, 1, 1, 2]))):
            toks.append("unique")
            if self.rng.random() < 0.75:
                toks.append(f"relate[{self.rng.choice(_RELS)}]")
            else:
                toks.append(f"same_{self.rng.choice(list(_ATTRS.keys()))}")
            toks += self._filters(int(self.rng.integers(0, 3)))
        return toks

    def sample_tokens(self) -> List[str]:
        r = self.rng.random()
        pre = lambda chain: list(reversed(chain))  # prefix order = outermost first
        if r < 0.45:
            c = self._chain()
            if self.rng.random() < 0.5:
                return [str(self.rng.choice(["count", "exist"]))] + pre(c)
            return [f"query_{self.rng.choice(list(_ATTRS.keys()))}", "unique"] + pre(c)
        if r < 0.65:
            op = str(self.rng.choice(["equal_integer", "less_than", "greater_than"]))
            return [op, "count"] + pre(self._chain()) + ["count"] + pre(self._chain())
        if r < 0.85:
            a = str(self.rng.choice(list(_ATTRS.keys())))
            return [f"equal_{a}", f"query_{a}", "unique"] + pre(self._chain()) + [f"query_{a}", "unique"] + pre(self._chain())
        op = str(self.rng.choice(["union", "intersect"]))
        return [str(self.rng.choice(["count", "exist"])), op] + pre(self._chain()) + pre(self._chain())

    def sample(self, batch: int, length: int = 26) -> torch.Tensor:
        out = np.zeros((batch, length), dtype=np.int64)
        for b in range(batch):
            while True:
                toks = self.sample_tokens()
                if len(toks) <= length:
                    break
            out[b, : len(toks)] = [self.vocab.get_token_index(t, "programs") for t in toks]
        return torch.from_numpy(out)

    def garbage(self, batch: int, length: int = 26) -> torch.Tensor:
        """uniform random token ids: mostly invalid programs (early-REINFORCE regime, nmn.py:235-238)"""
        v = self.vocab.get_vocab_size("programs")
        return torch.from_numpy(self.rng.integers(0, v, size=(batch, length), dtype=np.int64))


def programs_from_tokens(vocabulary, token_lists: List[List[str]], length: Optional[int] = None) -> torch.Tensor:
    length = length or max(1, max(len(t) for t in token_lists))
    out = torch.zeros(len(token_lists), length, dtype=torch.int64)
    for i, toks in enumerate(token_lists):
        for j, t in enumerate(toks):
            out[i, j] = vocabulary.get_token_index(t, "programs")
    return out


def make_questions(batch: int, vocab_size: int, seed: int = 0, max_length: int = 40, min_length: int = 5) -> torch.Tensor:
    rng = np.random.default_rng(4000 + seed)
    out = np.zeros((batch, max_length), dtype=np.int64)
    for b in range(batch):
        n = int(rng.integers(min_length, max_length + 1))
        out[b, :n] = rng.integers(len(SPECIAL_TOKENS), vocab_size, size=n)
    return torch.from_numpy(out)


def make_seq2seq_state_dict(vocab_src: int = 93, vocab_tgt: int = 44, hidden: int = 256, seed: int = 0,
                            gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Reference-shaped ``ProgramGenerator`` / ``QuestionReconstructor`` parameters (AllenNLP ``SimpleSeq2Seq`` key
    names) with PyTorch's default initialisations; ``gain`` > 1 scales the output projection
    so that token decisions have realistic margins."""
    g = torch.Generator().manual_seed(7000 + seed)
    k = 1.0 / hidden ** 0.5
    u = lambda *shape: (torch.rand(*shape, generator=g) * 2 - 1) * k
    sd = OrderedDict()
    emb = torch.randn(vocab_src, hidden, generator=g) * (2.0 / (vocab_src + hidden)) ** 0.5
    emb[0] = 0  # padding_index
    sd["_source_embedder.token_embedder_tokens.weight"] = emb
    for layer in range(2):
        sd[f"_encoder._module.weight_ih_l{layer}"] = u(4 * hidden, hidden)
        sd[f"_encoder._module.weight_hh_l{layer}"] = u(4 * hidden, hidden)
        sd[f"_encoder._module.bias_ih_l{layer}"] = u(4 * hidden)
        sd[f"_encoder._module.bias_hh_l{layer}"] = u(4 * hidden)
    sd["_target_embedder.weight"] = torch.randn(vocab_tgt, hidden, generator=g) * (2.0 / (vocab_tgt + hidden)) ** 0.5
    sd["_decoder_cell.weight_ih"] = u(4 * hidden, 2 * hidden)
    sd["_decoder_cell.weight_hh"] = u(4 * hidden, hidden)
    sd["_decoder_cell.bias_ih"] = u(4 * hidden)
    sd["_decoder_cell.bias_hh"] = u(4 * hidden)
    sd["_output_projection_layer.weight"] = u(vocab_tgt, hidden) * gain
    sd["_output_projection_layer.bias"] = u(vocab_tgt) * gain
    return sd


# ------------------------------------------------------------------------------------------------
# joint-training batches: questions that determine their programs, supervision flags, remaining weights
# ------------------------------------------------------------------------------------------------
def questions_for_programs(programs: torch.Tensor, vocab_size: int, seed: int = 0, max_length: int = 40) -> torch.Tensor:
    """Synthetic questions ``(B, max_length)`` from which the program is recoverable, standing in for CLEVR's
    question -> program mapping (no dataset in this environment): program token p at prefix position i becomes the
    question word ``4 + 2 * (p - 4) + (i % 2)``, followed by 0..14 filler words from the top of the vocabulary (real
    questions are longer than their programs' 'content').  Token ids stay below ``vocab_size``; 0 pads."""
    rng = np.random.default_rng(5000 + seed)
    prog = programs.numpy()
    B = prog.shape[0]
    content_top = 4 + 2 * 40
    assert vocab_size > content_top + 1, "question vocabulary too small for the synthetic mapping"
    out = np.zeros((B, max_length), dtype=np.int64)
    for b in range(B):
        toks = [int(p) for p in prog[b] if p != 0]
        q = [4 + 2 * (p - 4) + (i % 2) for i, p in enumerate(toks)]
        q += list(rng.integers(content_top, vocab_size, size=int(rng.integers(0, 15))))
        q = q[:max_length]
        out[b, : len(q)] = q
    return torch.from_numpy(out)


def make_joint_batch(vocabulary, batch: int, seed: int = 0, program_length: int = 26, question_length: int = 40,
                     supervised_fraction: float = 0.5, with_images: bool = True) -> dict:
    """A ``JointTrainingDataset`` batch (probnmn/data/datasets.py:209-228) of synthetic tensors on the host: ``question``
    (B, 40), ``program`` (B, 26), ``answer`` (B,), ``image`` (B, 1024, 14, 14) fp32, ``supervision`` (B,) with
    Bernoulli(0.5) flags (``SupervisionWeightedRandomSampler`` balances the two kinds, data/samplers.py:17-26)."""
    programs = ProgramSampler(vocabulary, seed=seed).sample(batch, program_length)
    rng = np.random.default_rng(6000 + seed)
    return {
        "question": questions_for_programs(programs, vocabulary.get_vocab_size("questions"), seed, question_length),
        "program": programs,
        "answer": make_answers(batch, seed),
        "image": make_features(batch, seed) if with_images else None,
        "supervision": torch.from_numpy((rng.random(batch) < supervised_fraction).astype(np.int64)),
    }


def make_prior_state_dict(vocab: int = 44, hidden: int = 256, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Reference-shaped ``ProgramPrior`` parameters (AllenNLP key names; the output layer is tied to the embedding,
    program_prior.py:59-62) with PyTorch's default initialisations."""
    g = torch.Generator().manual_seed(8000 + seed)
    k = 1.0 / hidden ** 0.5
    u = lambda *shape: (torch.rand(*shape, generator=g) * 2 - 1) * k
    sd = OrderedDict()
    emb = torch.randn(vocab, hidden, generator=g) * (2.0 / (vocab + hidden)) ** 0.5
    emb[0] = 0
    sd["_embedder.token_embedder_programs.weight"] = emb
    for layer in range(2):
        sd[f"_encoder._module.weight_ih_l{layer}"] = u(4 * hidden, hidden)
        sd[f"_encoder._module.weight_hh_l{layer}"] = u(4 * hidden, hidden)
        sd[f"_encoder._module.bias_ih_l{layer}"] = u(4 * hidden)
        sd[f"_encoder._module.bias_hh_l{layer}"] = u(4 * hidden)
    sd["_projection_layer.weight"] = u(hidden, hidden)
    sd["_output_layer.weight"] = emb
    return sd
