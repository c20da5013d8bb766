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


def _records(plan, which, width):
    lib = L.lib()
    n = lib.pnmn_debug_plan_records(plan, which, None, 0)
    buf = np.zeros((n, width), dtype=np.uint8)
    if n:
        lib.pnmn_debug_plan_records(plan, which, buf.ctypes.data, n)
    return buf


@pytest.mark.parametrize("flags", [0, L.PLAN_INPUT_BY_ROW])
def test_forward_half_plan_is_the_forward_pass_of_the_full_plan(model, flags):
    """PNMN_PLAN_FORWARD_HALF (include/pnmn.h): the need_grad = 0 stand-in must agree with the full plan of the same
    programs record for record -- same forward task list (same arena units, same configuration ids, same dependencies),
    same validity -- and its arena sizes must cover what the full plan's backward pass allocates."""
    lib = L.lib()
    vocab = model.vocabulary
    for seed in range(12):
        sampler = ProgramSampler(vocab, seed=seed)
        programs = torch.cat([sampler.sample(40 + 9 * seed, 26), sampler.garbage(6, 26)]).contiguous()
        B, Lp = programs.shape
        ptr = ctypes.cast(programs.data_ptr(), ctypes.POINTER(ctypes.c_int64))
        half = lib.pnmn_plan_create_ex(model._model_handle, ptr, B, Lp, 0, flags | L.PLAN_FORWARD_HALF)
        full = lib.pnmn_plan_create_ex(model._model_handle, ptr, B, Lp, 1, flags)
        try:
            rec_h, rec_f = _records(half, 0, 128), _records(full, 0, 128)
            assert rec_h.shape == rec_f.shape and (rec_h == rec_f).all()
            cfg_h, cfg_f = _records(half, 2, 48), _records(full, 2, 48)
            assert len(cfg_h) <= len(cfg_f) and (cfg_h == cfg_f[: len(cfg_h)]).all()
            sizes_h, sizes_f = (ctypes.c_int64 * L.SZ_COUNT)(), (ctypes.c_int64 * L.SZ_COUNT)()
            lib.pnmn_plan_sizes(half, sizes_h)
            lib.pnmn_plan_sizes(full, sizes_f)
            for slot in (L.SZ_ARENA16, L.SZ_ARENA18, L.SZ_ARENA22, L.SZ_MAPS, L.SZ_DMAPS, L.SZ_IDX, L.SZ_AIN):
                assert sizes_h[slot] >= sizes_f[slot], (seed, slot)
            assert sizes_h[L.SZ_ARENA16] < 2 * sizes_f[L.SZ_ARENA16]      # ... without doubling the footprint
            valid_h, valid_f = np.zeros(B, np.uint8), np.zeros(B, np.uint8)
            lib.pnmn_plan_valid(half, valid_h.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
            lib.pnmn_plan_valid(full, valid_f.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
            assert (valid_h == valid_f).all()
            assert lib.pnmn_debug_plan_records(half, 1, None, 0) == 0
        finally:
            lib.pnmn_plan_destroy(half)
            lib.pnmn_plan_destroy(full)
