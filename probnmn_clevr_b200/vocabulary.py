"""A dependency-free stand-in for ``allennlp.data.Vocabulary`` (allennlp==0.9.0 is the reference's
pinned dependency, ``requirements.txt:1``; it is not installable here).

Only the methods the hot path touches are provided, with AllenNLP's signatures:
``get_token_index``, ``get_token_from_index``, ``get_vocab_size``,
``get_index_to_token_vocabulary``, ``get_token_to_index_vocabulary``, ``from_files``
(call sites: probnmn/models/nmn.py:63,87,132,204,251; probnmn/modules/seq2seq_base.py:61-69).
Any object with these methods (including a real AllenNLP ``Vocabulary``) can be passed to the models.

The on-disk format is the one ``scripts/preprocess/build_vocabulary.py:128-149`` writes: one token per
line per namespace, ``@@PADDING@@`` implicit at index 0 for padded namespaces, and the names of
non-padded namespaces in ``non_padded_namespaces.txt``.
"""
import os
from typing import Dict, Iterable, List, Optional

PADDING, UNKNOWN, START, END = "@@PADDING@@", "@@UNKNOWN@@", "@start@", "@end@"
SPECIAL_TOKENS = [PADDING, UNKNOWN, START, END]

_COLORS = ["blue", "brown", "cyan", "gray", "green", "purple", "red", "yellow"]
_MATERIALS = ["metal", "rubber"]
_SHAPES = ["cube", "cylinder", "sphere"]
_SIZES = ["large", "small"]
_RELATIONS = ["behind", "front", "left", "right"]

# The 40 CLEVR v1.0 program tokens (function name + "[value]"), as build_vocabulary.py:84-103 forms them.
CLEVR_PROGRAM_TOKENS: List[str] = sorted(
    ["count", "exist", "greater_than", "intersect", "less_than", "scene", "union", "unique"]
    + [f"equal_{a}" for a in ("color", "integer", "material", "shape", "size")]
    + [f"filter_color[{v}]" for v in _COLORS]
    + [f"filter_material[{v}]" for v in _MATERIALS]
    + [f"filter_shape[{v}]" for v in _SHAPES]
    + [f"filter_size[{v}]" for v in _SIZES]
    + [f"query_{a}" for a in ("color", "material", "shape", "size")]
    + [f"relate[{v}]" for v in _RELATIONS]
    + [f"same_{a}" for a in ("color", "material", "shape", "size")]
)

# The 28 CLEVR answers, sorted as strings (build_vocabulary.py:124), then @@UNKNOWN@@ (index 28).
CLEVR_ANSWER_TOKENS: List[str] = sorted(
    [str(i) for i in range(11)] + _COLORS + _MATERIALS + _SHAPES + _SIZES + ["yes", "no"]
)


class Vocabulary:
    def __init__(self, namespaces: Dict[str, List[str]], non_padded: Iterable[str] = ("answers",)):
        self._non_padded = set(non_padded)
        self._index_to_token: Dict[str, Dict[int, str]] = {}
        self._token_to_index: Dict[str, Dict[str, int]] = {}
        for ns, tokens in namespaces.items():
            self._index_to_token[ns] = dict(enumerate(tokens))
            self._token_to_index[ns] = {t: i for i, t in enumerate(tokens)}

    # ---- AllenNLP-compatible surface -----------------------------------------------------------
    def get_token_index(self, token: str, namespace: str = "tokens") -> int:
        t2i = self._token_to_index[namespace]
        if token in t2i:
            return t2i[token]
        return t2i[UNKNOWN]  # AllenNLP: OOV maps to the namespace's @@UNKNOWN@@ (KeyError if none)

    def get_token_from_index(self, index: int, namespace: str = "tokens") -> str:
        return self._index_to_token[namespace][int(index)]

    def get_vocab_size(self, namespace: str = "tokens") -> int:
        return len(self._token_to_index[namespace])

    def get_index_to_token_vocabulary(self, namespace: str = "tokens") -> Dict[int, str]:
        return self._index_to_token[namespace]

    def get_token_to_index_vocabulary(self, namespace: str = "tokens") -> Dict[str, int]:
        return self._token_to_index[namespace]

    @classmethod
    def from_files(cls, directory: str) -> "Vocabulary":
        with open(os.path.join(directory, "non_padded_namespaces.txt")) as f:
            non_padded = [line.strip() for line in f if line.strip()]
        namespaces: Dict[str, List[str]] = {}
        for fname in sorted(os.listdir(directory)):
            if not fname.endswith(".txt") or fname == "non_padded_namespaces.txt":
                continue
            ns = fname[:-4]
            with open(os.path.join(directory, fname)) as f:
                tokens = [line.rstrip("\n") for line in f if line.rstrip("\n") != ""]
            namespaces[ns] = tokens if ns in non_padded else [PADDING] + tokens
        return cls(namespaces, non_padded)

    def save_to_files(self, directory: str) -> None:
        os.makedirs(directory, exist_ok=True)
        with open(os.path.join(directory, "non_padded_namespaces.txt"), "w") as f:
            f.write("\n".join(sorted(self._non_padded)))
        for ns, i2t in self._index_to_token.items():
            tokens = [i2t[i] for i in range(len(i2t))]
            if ns not in self._non_padded:
                tokens = tokens[1:]
            with open(os.path.join(directory, ns + ".txt"), "w") as f:
                f.write("".join(t + "\n" for t in tokens))

    # ---- synthetic CLEVR vocabulary --------------------------------------------------------------
    @classmethod
    def clevr(cls, num_question_tokens: int = 93, question_tokens: Optional[List[str]] = None) -> "Vocabulary":
        """programs: 4 specials + the 40 CLEVR program tokens (44); answers: 28 + @@UNKNOWN@@;
        questions: 4 specials + (num_question_tokens - 4) tokens (the real set depends on the CLEVR
        json, build_vocabulary.py:65-80; placeholders ``w000..`` are used when none is given)."""
        if question_tokens is None:
            question_tokens = [f"w{i:03d}" for i in range(num_question_tokens - 4)]
        return cls(
            {
                "programs": SPECIAL_TOKENS + CLEVR_PROGRAM_TOKENS,
                "questions": SPECIAL_TOKENS + sorted(question_tokens),
                "answers": CLEVR_ANSWER_TOKENS + [UNKNOWN],
            }
        )
