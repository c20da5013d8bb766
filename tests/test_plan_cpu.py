"""CPU: structure of the task lists the program compiler emits for the persistent executor (no device needed:
pnmn_plan_create runs on the host, pnmn_debug_plan_meta exports the dependency lists).

The reference walks each program token by token inside forward (probnmn/models/nmn.py:191-238); here every sample becomes
a set of dependent tasks, and the executor's in-order fetch is only deadlock-free if every task's producers sit in front
of it in the list."""
import ctypes

import numpy as np
import pytest
import torch

from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, programs_from_tokens
from probnmn_clevr_b200.vocabulary import Vocabulary


@pytest.fixture(scope="module")
def model():
    m = NeuralModuleNetwork(Vocabulary.clevr())
    m._ensure_flat()
    return m


def _metas(model, programs, need_grad=True):
    lib = L.lib()
    plan = model._compile(programs.contiguous(), need_grad, None)
    try:
        out = []
        for p in (0, 1):
            n = lib.pnmn_debug_plan_meta(plan, p, None, 0)
            buf = np.zeros((n, 16), dtype=np.int32)
            if n:
                lib.pnmn_debug_plan_meta(plan, p, buf.ctypes.data, n)
            out.append(buf)
        valid = np.zeros(programs.shape[0], dtype=np.uint8)
        lib.pnmn_plan_valid(plan, valid.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return out, valid
    finally:
        lib.pnmn_plan_destroy(plan)


def _depth(meta):
    depth = np.zeros(len(meta), dtype=int)
    for i, row in enumerate(meta):
        deps = row[2:2 + row[1]]
        if len(deps):
            depth[i] = depth[deps].max() + 1
    return int(depth.max()) + 1 if len(meta) else 0


def _check_list(meta):
    for i, row in enumerate(meta):
        assert row[0] in (0, 1)                       # TASK_CONV / TASK_ELT
        n = row[1]
        assert 0 <= n <= 10
        deps = row[2:2 + n]
        assert (deps >= 0).all() and (deps < i).all(), f"task {i} waits for a task that is not in front of it: {deps}"
        assert len(set(deps.tolist())) == n           # no duplicate producers
        assert (row[2 + n:12] == -1).all()


@pytest.mark.parametrize("seed,length", [(0, 40), (3, 26)])
def test_task_lists_are_topologically_ordered(model, seed, length):
    vocab = model.vocabulary
    sampler = ProgramSampler(vocab, seed=seed)
    programs = torch.cat([sampler.sample(48, length), sampler.garbage(16, length)])
    (fwd, bwd), valid = _metas(model, programs)
    assert valid[:48].all()
    assert len(fwd) > 0 and len(bwd) > len(fwd)
    _check_list(fwd)
    _check_list(bwd)
    # the forward-only plan (evaluation) has no backward list
    (fwd2, bwd2), _ = _metas(model, programs, need_grad=False)
    assert len(fwd2) == len(fwd) and len(bwd2) == 0


def test_independent_branches_run_as_concurrent_strands(model):
    """equal_integer(count(A), count(B)): the two chains between a `scene` and the comparison are independent
    (nmn.py:216-222), so the dependency depth of the two-branch program is that of its LONGER branch plus the comparison,
    not the sum of both."""
    vocab = model.vocabulary
    a = ["filter_color[red]", "filter_shape[cube]", "relate[left]", "unique", "filter_size[large]", "scene"]
    b = ["filter_material[metal]", "scene"]
    both = [["equal_integer", "count"] + a + ["count"] + b]
    only_a = [["count"] + a]
    only_b = [["count"] + b]
    d = {}
    for name, toks in (("both", both), ("a", only_a), ("b", only_b)):
        (fwd, bwd), valid = _metas(model, programs_from_tokens(vocab, toks, 26))
        assert valid.all()
        _check_list(fwd)
        _check_list(bwd)
        d[name] = (_depth(fwd), _depth(bwd))
    stem_and_gather = 3        # two stem convs + the final gather, present in every program
    compare = 3                # projection + two 3x3 convs
    assert d["both"][0] == max(d["a"][0], d["b"][0]) + compare
    assert d["both"][0] < d["a"][0] + d["b"][0] - stem_and_gather
    assert d["both"][1] < d["a"][1] + d["b"][1]


def test_malformed_binary_programs_fall_back_to_one_chain(model):
    """three `scene` tokens / two binary modules: compiled without strands (a value may then have two consumers);
    still valid where the reference executes them, and still a well-formed list."""
    vocab = model.vocabulary
    toks = [["count", "union", "filter_color[red]", "scene", "intersect", "filter_shape[cube]", "scene",
             "filter_size[small]", "scene"]]
    (fwd, bwd), valid = _metas(model, programs_from_tokens(vocab, toks, 26))
    _check_list(fwd)
    _check_list(bwd)
    assert valid.all()
