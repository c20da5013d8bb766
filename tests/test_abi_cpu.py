"""CPU: the C-ABI library loads, exports every symbol include/pnmn.h declares, and its host-only entry
points (model description + program compiler) agree with the reference's interpreter semantics.
No compute entry point is called here (there is no GPU in this container and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler
from probnmn_clevr_b200.vocabulary import Vocabulary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "nmn_golden.npz")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        L.build()
    return L.lib()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "pnmn.h")).read()
    declared = set(re.findall(r"\b(pnmn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/pnmn.h but not exported"
    assert set(L.EXPORTS) <= declared
    assert lib.pnmn_version() == 1


def _plan_valid(model, programs, need_grad=1):
    lib = L.lib()
    programs = programs.contiguous()
    plan = lib.pnmn_plan_create(model._model_handle, ctypes.cast(programs.data_ptr(), ctypes.POINTER(ctypes.c_int64)),
                                programs.shape[0], programs.shape[1], need_grad)
    assert plan
    valid = torch.empty(programs.shape[0], dtype=torch.uint8)
    lib.pnmn_plan_valid(plan, ctypes.cast(valid.data_ptr(), ctypes.POINTER(ctypes.c_uint8)))
    sizes = (ctypes.c_int64 * L.SZ_COUNT)()
    lib.pnmn_plan_sizes(plan, sizes)
    stats = (ctypes.c_int64 * 16)()
    lib.pnmn_plan_stats(plan, stats)
    lib.pnmn_plan_destroy(plan)
    return valid.numpy(), list(sizes), list(stats)


@pytest.fixture(scope="module")
def model(lib):
    m = NeuralModuleNetwork(Vocabulary.clevr(), class_projection_channels=8, classifier_linear_size=8)
    m._ensure_flat()  # host-side only: builds the flat parameter buffer and the model description
    return m


@pytest.mark.parametrize("name", ["semantic", "sampled", "garbage"])
def test_program_compiler_validity_matches_reference(model, name):
    g = np.load(GOLDEN)
    valid, sizes, stats = _plan_valid(model, torch.from_numpy(g[f"{name}.programs"]))
    assert valid.tolist() == g[f"{name}.valid"].tolist()
    assert stats[0] == int(g[f"{name}.valid"].sum())
    assert all(s >= 0 for s in sizes)


def test_program_compiler_edge_cases(model):
    vocab = model.vocabulary
    # empty batch row / all padding -> valid (classifier sees the raw stem features); out-of-vocabulary id -> invalid
    progs = torch.zeros(3, 5, dtype=torch.int64)
    progs[1, 0] = 999
    progs[2, :2] = torch.tensor([vocab.get_token_index("count", "programs"), vocab.get_token_index("scene", "programs")])
    valid, _, stats = _plan_valid(model, progs)
    assert valid.tolist() == [1, 0, 1]
    assert stats[1] == 2  # one Query module = two 3x3 convs
    # maximum-length programs from the grammar are all valid, with and without gradient planning
    big = ProgramSampler(vocab, seed=5).sample(64, 40)
    for ng in (0, 1):
        valid, sizes, stats = _plan_valid(model, big, ng)
        assert valid.all()
        assert stats[1] > 0 and sizes[L.SZ_ARENA16] > 0


def test_lazy_metrics_is_a_dict_that_reads_late():
    """The training-mode "metrics" entry (nmn.py:273-274) must stay a dict for the reference's trainer
    (trainers/_trainer.py:197) but must not synchronise the device until a value is read."""
    import json

    import torch

    from probnmn_clevr_b200.nmn import _Accuracy, _LazyMetrics

    calls = []
    acc = _Accuracy()
    acc(torch.tensor(3), 4)
    acc(1, 4)
    state = acc.snapshot(reset=True)
    assert acc.get_metric() == 0.0  # reset happened at snapshot time

    def read():
        calls.append(1)
        return _Accuracy.value(state)

    m = _LazyMetrics({"answer_accuracy": read, "average_invalid": lambda: 0.25})
    assert isinstance(m, dict) and len(m) == 2 and "answer_accuracy" in m and list(m) == ["answer_accuracy", "average_invalid"]
    assert calls == []  # nothing evaluated so far
    assert m["answer_accuracy"] == 0.5 and calls == [1]
    assert dict(m) == {"answer_accuracy": 0.5, "average_invalid": 0.25} == {**m}
    assert json.loads(json.dumps(m)) == {"answer_accuracy": 0.5, "average_invalid": 0.25}
    assert calls == [1]  # evaluated once
