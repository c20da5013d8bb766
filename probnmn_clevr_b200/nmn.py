"""Drop-in ``NeuralModuleNetwork`` (reference: probnmn/models/nmn.py) whose stem + module executor run in
hand-written sm_100a CUDA behind the C ABI of ``include/pnmn.h``.

Same constructor, ``from_config``, ``forward(features, programs, answers=None)`` -> ``{"predictions",
"loss"[, "metrics"]}``, ``get_metrics`` and state-dict keys / shapes as the reference
(SURVEY.md §8b, appendix B), so ``JointTrainingTrainer`` / ``ModuleTrainingTrainer``,
``CheckpointManager.load`` and Adam work unchanged.  What differs is how the work is done:

* the reference interprets every sample's program in python at batch size 1 with a device->host sync
  per sample (nmn.py:191-238); here ONE copy of the token matrix goes to the host, the C++ program
  compiler turns the batch into task tables, and the tensor-core executor runs them level by level;
* all parameters are views into one flat fp32 buffer (reference names and shapes are kept), so the
  library sees a base pointer + offsets and the gradient all-reduce is a single NCCL call.

There is no CPU or eager fallback: without the CUDA library / a CUDA device ``forward`` raises.
"""
import ctypes
import threading
import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib as L
from .vocabulary import Vocabulary

SKIP_TOKENS = ("@@PADDING@@", "@@UNKNOWN@@", "@start@", "@end@", "unique")


def module_class_of(token: str) -> Optional[str]:
    """token -> module family; the substring rules (and their order) of nmn.py:90-111."""
    if token in SKIP_TOKENS:
        return None
    if token == "scene":
        return "scene"
    if token == "intersect":
        return "and"
    if token == "union":
        return "or"
    if "equal" in token or token in ("less_than", "greater_than"):
        return "comparison"
    if "query" in token or token in ("exist", "count"):
        return "query"
    if "relate" in token:
        return "relate"
    if "same" in token:
        return "same"
    return "attention"


_KIND = {None: L.TOK_SKIP, "scene": L.TOK_SCENE, "and": L.TOK_AND, "or": L.TOK_OR, "comparison": L.TOK_COMPARE,
         "query": L.TOK_QUERY, "relate": L.TOK_RELATE, "same": L.TOK_SAME, "attention": L.TOK_ATTENTION}

# (attribute name, out_ch, in_ch, kernel, kaiming-normal init?) per family, in forward order.
# Shapes / names: nmn_modules.py:71-79 (Attention), 110-116 (Query), 144-157 (Relate), 194-197 (Same),
# 231-237 (Comparison: `projection` keeps nn.Conv2d's default init).
def _family_spec(family: str, d: int):
    c3 = lambda n: (n, d, d, 3, True)
    return {
        "attention": [c3("conv1"), c3("conv2"), ("conv3", 1, d, 1, True)],
        "query": [c3("conv1"), c3("conv2")],
        "relate": [c3(f"conv{i}") for i in range(1, 6)] + [("conv6", 1, d, 1, True)],
        "same": [("conv", 1, d + 1, 1, True)],
        "comparison": [("projection", d, 2 * d, 1, False), c3("conv1"), c3("conv2")],
    }[family]


class _ModuleParameters(nn.Module):
    """Parameter holder of one neural module; the math runs in the CUDA executor, not in ``forward``."""

    def __init__(self, family: str, dim: int):
        super().__init__()
        self.family = family
        for name, cout, cin, k, kaiming in _family_spec(family, dim):
            conv = nn.Conv2d(cin, cout, kernel_size=k)
            if kaiming:
                nn.init.kaiming_normal_(conv.weight)
            setattr(self, name, conv)

    def forward(self, *args):  # pragma: no cover
        raise RuntimeError("neural modules are executed by the CUDA program executor (NeuralModuleNetwork.forward)")


class _Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


class _Average:
    def __init__(self):
        self.total, self.count = 0.0, 0

    def __call__(self, v):
        self.total += float(v)
        self.count += 1

    def get_metric(self, reset=False):
        out = self.total / self.count if self.count else 0.0
        if reset:
            self.total, self.count = 0.0, 0
        return out


class _Accuracy:
    """BooleanAccuracy (nmn.py:120-121).  Per-batch correct counts may arrive as DEVICE scalars: they are only read
    (one device -> host sync) when somebody asks for the value, so that ``forward`` never waits for the device."""

    def __init__(self):
        self.total, self.count, self.pending = 0.0, 0, []

    def __call__(self, correct, total):
        if torch.is_tensor(correct):
            self.pending.append(correct)
        else:
            self.total += float(correct)
        self.count += int(total)

    def snapshot(self, reset=False):
        state = (self.total, self.count, list(self.pending))
        if reset:
            self.total, self.count, self.pending = 0.0, 0, []
        return state

    @staticmethod
    def value(state):
        total, count, pending = state
        total += sum(float(t.item()) for t in pending)
        return total / count if count else 0.0

    def get_metric(self, reset=False):
        return self.value(self.snapshot(reset))


class _LazyMetrics(dict):
    """The ``"metrics"`` entry of the training-mode output (nmn.py:273-274): a real ``dict`` (the reference's trainer
    tests ``isinstance(..., dict)``, trainers/_trainer.py:197) whose values are computed -- and the device synchronised --
    the first time one of them is read.  Keys, length and membership never synchronise."""

    def __init__(self, thunks):
        super().__init__({k: None for k in thunks})
        self._thunks = dict(thunks)

    def _force(self):
        if self._thunks is not None:
            thunks, self._thunks = self._thunks, None
            for k, f in thunks.items():
                super().__setitem__(k, f())
        return self

    def __getitem__(self, k):
        self._force()
        return super().__getitem__(k)

    def get(self, k, default=None):
        self._force()
        return super().get(k, default)

    def __iter__(self):  # (also keeps dict(...) / {**...} off the C fast path that would copy the placeholders)
        return iter(list(super().keys()))

    def items(self):
        self._force()
        return super().items()

    def values(self):
        self._force()
        return super().values()

    def copy(self):
        return dict(self._force())

    def __eq__(self, other):
        self._force()
        if isinstance(other, _LazyMetrics):
            other._force()
        return super().__eq__(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._force()
        return super().__repr__()


# ------------------------------------------------------------------------------------------------
# workspaces: persistent, zero-initialised arenas (the executor relies on permanent zero padding)
# ------------------------------------------------------------------------------------------------
class _Workspace:
    ARENAS = {L.SZ_ARENA16: "arena16", L.SZ_ARENA18: "arena18", L.SZ_ARENA22: "arena22", L.SZ_MAPS: "maps",
              L.SZ_DMAPS: "dmaps", L.SZ_AIN: "ain"}

    def __init__(self, device):
        self.device = device
        self.t: Dict[str, torch.Tensor] = {}
        self.scratch = torch.zeros(64, device=device)

    def ensure(self, sizes):
        for slot, name in self.ARENAS.items():
            need = int(sizes[slot])
            cur = self.t.get(name)
            if cur is None or cur.numel() < need:
                self.t[name] = None  # free first
                self.t[name] = torch.zeros(int(need * 1.25) + 1024, dtype=torch.float32, device=self.device)
        need = int(sizes[L.SZ_IDX])
        if self.t.get("idx") is None or self.t["idx"].numel() < need:
            self.t["idx"] = torch.zeros(need * 2 + 16, dtype=torch.int32, device=self.device)
        need = int(sizes[L.SZ_BLOB])
        if self.t.get("blob") is None or self.t["blob"].numel() < need:
            self.t["blob"] = torch.zeros(int(need * 1.5) + 4096, dtype=torch.uint8, device=self.device)


class _WorkspacePool:
    def __init__(self):
        self.free: Dict[torch.device, List[_Workspace]] = {}

    def acquire(self, device) -> _Workspace:
        lst = self.free.setdefault(device, [])
        return lst.pop() if lst else _Workspace(device)

    def release(self, ws: _Workspace):
        self.free.setdefault(ws.device, []).append(ws)


_POOL = _WorkspacePool()


class _BlobPool:
    """Device buffers for task tables uploaded ahead of time (``precompile``).  They are kept and recycled instead of being
    allocated per plan: a buffer that is filled on the upload stream and read on the compute stream can only go back to
    the caching allocator with deferred frees, and the ``cudaMalloc`` calls that replace it (every few steps) stall the
    device.  A buffer is reused once the compute stream has passed the point where its last user released it."""

    def __init__(self):
        self.free: List[Tuple[torch.Tensor, Optional[torch.cuda.Event]]] = []
        self.lock = threading.Lock()

    def acquire(self, nbytes: int, device, stream) -> torch.Tensor:
        with self.lock:
            for k, (t, ev) in enumerate(self.free):
                if t.device == device and t.numel() >= nbytes and (ev is None or ev.query()):
                    del self.free[k]
                    return t
        with torch.cuda.stream(stream):
            return torch.empty(int(nbytes * 1.5) + 4096, dtype=torch.uint8, device=device)

    def release(self, t: torch.Tensor) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
        with self.lock:
            self.free.append((t, ev))


_BLOBS = _BlobPool()


class _Run:
    """One forward's plan + workspace; released after backward (or when dropped)."""

    def __init__(self, plan, ws, bufs, gflat_box=None, lo=0, hi=0, stream=None, staged=False, extra_ws=None):
        self.plan, self.ws, self.bufs, self.gflat_box = plan, ws, bufs, gflat_box
        self.lo, self.hi, self.stream, self.staged = lo, hi, stream, staged   # rows of the batch, side stream, prestaged inputs
        self.extra_ws = extra_ws    # workspace that holds the prestaged inputs of ALL chunks (released with this run)
        self.blob = None
        self.xin_off = 0
        # split compile (NeuralModuleNetwork._single_run): `plan` is the forward half, `full` the future of
        # (full plan, its uploaded task tables, upload event) that the backward pass runs from
        self.full = None

    def take_full(self):
        """(full plan, device blob, upload event) of a split compile; the run owns them from here on."""
        plan, blob, event = self.full.result()
        self.full = None
        if self.plan is not None:
            L.lib().pnmn_plan_destroy(self.plan)
        self.plan = plan
        if getattr(self, "pool_blob", None) is not None:
            _BLOBS.release(self.pool_blob)
        self.pool_blob = blob
        return plan, blob, event

    def close(self):
        if self.full is not None:    # the backward pass never ran: the full plan still has to be collected
            try:
                self.take_full()
            except Exception:
                self.full = None
        if self.plan is not None:
            L.lib().pnmn_plan_destroy(self.plan)
            self.plan = None
        if getattr(self, "pool_blob", None) is not None:
            _BLOBS.release(self.pool_blob)
            self.pool_blob = None
        if self.ws is not None:
            _POOL.release(self.ws)
            self.ws = None
        if self.extra_ws is not None:
            _POOL.release(self.extra_ws)
            self.extra_ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ExecutorFn(torch.autograd.Function):
    """final_module_outputs = executor(features, programs; stem + module parameters).

    The 222 stem / module parameters do NOT travel through autograd one by one (that cost ~2.5 ms of host time per
    step): the backward pass accumulates straight into the model's persistent flat gradient buffer, whose per-parameter
    views are attached as ``.grad`` (``NeuralModuleNetwork._attach_grads``).  ``anchor`` is a dummy differentiable input
    that makes autograd call ``backward``."""

    @staticmethod
    def forward(ctx, features, anchor, runs, model):
        B = features.shape[0]
        dev = features.device
        final = torch.empty(B, 128, 14, 14, dtype=torch.float32, device=dev)
        entry = L.lib().pnmn_nmn_forward_f16 if features.dtype == torch.float16 else L.lib().pnmn_nmn_forward
        row_bytes = features[0].numel() * features.element_size() if B else 0
        current = torch.cuda.current_stream(dev)
        fork = None
        for run in runs:
            # a batch compiled as several plans (NeuralModuleNetwork.compile_chunks): every plan runs on its own stream, from
            # its own rows of `features` into its own rows of `final`
            st = current
            if run.stream is not None:
                if fork is None:
                    fork = torch.cuda.Event()
                    fork.record(current)
                run.stream.wait_event(fork)
                st = run.stream
            # (features = NULL: NeuralModuleNetwork.prestage already packed the weights and laid the features out)
            fptr = None if run.staged else ctypes.c_void_p(features.data_ptr() + run.lo * row_bytes)
            L.check(entry(run.plan, ctypes.byref(run.bufs), fptr, ctypes.c_void_p(final.data_ptr() + run.lo * 128 * 196 * 4),
                          ctypes.c_void_p(st.cuda_stream)), "pnmn_nmn_forward")
            if run.stream is not None:
                done = torch.cuda.Event()
                done.record(run.stream)
                current.wait_event(done)
        ctx.runs, ctx.model = runs, model
        return final

    @staticmethod
    def backward(ctx, grad_final):
        runs, model = ctx.runs, ctx.model
        if any(run.plan is None for run in runs):
            raise RuntimeError("NeuralModuleNetwork backward called twice (the plan was already released)")
        grad_final = grad_final.contiguous()
        dev = grad_final.device
        target, finish = model._attach_grads()
        with torch.cuda.device(dev):  # (the autograd thread's current device is not necessarily the tensors')
            current = torch.cuda.current_stream(dev)
            fork = None
            for run in runs:
                run.bufs.grads = target.data_ptr()
                split_stats = None
                if run.full is not None:
                    # split compile: the forward pass ran from the forward-half plan; the full plan of the same programs
                    # (same arena layout) was compiled and uploaded meanwhile on a helper thread
                    plan, blob, event = run.take_full()
                    current.wait_event(event)
                    run.bufs.blob = blob.data_ptr()
                    _, sizes, model.last_plan_stats = model._plan_info(plan, run.hi - run.lo)
                    split_stats = model.last_plan_stats
                    # (the forward half's arena sizes are an upper bound of the full plan's by construction -- every
                    # backward allocation site of the compiler is counted, tests/test_plan_cpu.py; never write past an arena)
                    for slot, name in _Workspace.ARENAS.items():
                        if name != "ain" and int(sizes[slot]) > run.ws.t[name].numel():
                            raise RuntimeError(f"internal error: the full plan needs a larger {name} arena than its forward half reserved")
                st = current
                if run.stream is not None:
                    if fork is None:
                        fork = torch.cuda.Event()
                        fork.record(current)
                    run.stream.wait_event(fork)
                    st = run.stream
                if model.backward_stats_hook is not None:   # (benchmarks: FLOP accounting of the pass being launched)
                    model.backward_stats_hook(split_stats if split_stats is not None
                                              else model._plan_info(run.plan, run.hi - run.lo)[2])
                L.check(L.lib().pnmn_nmn_backward(run.plan, ctypes.byref(run.bufs),
                                                  ctypes.c_void_p(grad_final.data_ptr() + run.lo * 128 * 196 * 4),
                                                  ctypes.c_void_p(st.cuda_stream)), "pnmn_nmn_backward")
                if run.stream is not None:
                    done = torch.cuda.Event()
                    done.record(run.stream)
                    current.wait_event(done)
        if finish is not None:
            finish()
        for run in runs:
            run.close()
        return None, None, None, None


class NeuralModuleNetwork(nn.Module):
    r"""
    Holds a stem, one neural module per program token and a classifier, and answers a batch of questions
    by executing their programs over image features (reference: probnmn/models/nmn.py:25-124).

    Parameters
    ----------
    vocabulary:
        Any object with AllenNLP's ``Vocabulary`` lookup methods (see ``vocabulary.py``), namespaces
        "programs" and "answers".
    image_feature_size: tuple (K, R, C), optional (default = (1024, 14, 14))
    module_channels: int, optional (default = 128) -- the CUDA executor is specialised for 128
    class_projection_channels: int, optional (default = 1024)
    classifier_linear_size: int, optional (default = 1024)
    """

    def __init__(
        self,
        vocabulary,
        image_feature_size: Tuple[int, int, int] = (1024, 14, 14),
        module_channels: int = 128,
        class_projection_channels: int = 1024,
        classifier_linear_size: int = 1024,
    ):
        super().__init__()
        if module_channels != 128 or tuple(image_feature_size[1:]) != (14, 14) or image_feature_size[0] % 128:
            raise ValueError("the B200 executor is built for module_channels=128 and (128k, 14, 14) features")
        self.vocabulary = vocabulary
        channels, height, width = image_feature_size
        self._in_channels = channels
        # "@@UNKNOWN@@" is never produced by the classifier (nmn.py:57-63)
        num_answers = len(vocabulary.get_index_to_token_vocabulary(namespace="answers")) - 1
        self._unknown_answer = vocabulary.get_token_index("@@UNKNOWN@@", namespace="answers")

        self.stem = nn.Sequential(
            nn.Conv2d(channels, module_channels, kernel_size=3, padding=1), nn.ReLU(),
            nn.Conv2d(module_channels, module_channels, kernel_size=3, padding=1), nn.ReLU(),
        )
        self.classifier = nn.Sequential(
            nn.Conv2d(module_channels, class_projection_channels, kernel_size=1), nn.ReLU(),
            nn.MaxPool2d(kernel_size=2, stride=2), _Flatten(),
            nn.Linear(class_projection_channels * height * width // 4, classifier_linear_size), nn.ReLU(),
            nn.Linear(classifier_linear_size, num_answers),
        )
        # one parameter holder per program token, registered under the token's name (nmn.py:86-115)
        self._token_family: Dict[str, Optional[str]] = {}
        for token in vocabulary.get_token_to_index_vocabulary("programs"):
            family = module_class_of(token)
            self._token_family[token] = family
            if family in ("attention", "query", "relate", "same", "comparison"):
                self.add_module(token, _ModuleParameters(family, module_channels))

        self._answer_accuracy = _Accuracy()
        self._average_invalid_programs = _Average()
        self._flat: Optional[torch.Tensor] = None
        self._layout: Optional[List[Tuple[str, int, int, torch.Size]]] = None
        self._exec_params: Optional[List[nn.Parameter]] = None
        self._gflat: Optional[torch.Tensor] = None
        self._gviews: Optional[List[torch.Tensor]] = None
        self._anchor: Optional[torch.Tensor] = None
        self._model_handle = None
        self._packed: Optional[torch.Tensor] = None
        self.last_plan_stats: Optional[List[int]] = None
        self._gflat_box: Dict[str, torch.Tensor] = {}
        self._grad_overlap = None
        self._upload_stream = None
        self._precompiled: list = []  # pending (programs, need_grad, future of a plan) entries, see precompile()
        self._prestaged = None        # (features tensor, workspace) of a prestage() call the next forward may use
        # > 1: a batch's programs are compiled as this many independent plans on as many host threads (_chunked_runs)
        self.compile_chunks = int(os.environ.get("PNMN_COMPILE_CHUNKS", "1"))
        self._chunk_streams = None
        # a training-mode forward whose programs were not compiled ahead (precompile) launches from the forward-half plan and
        # takes its backward pass from the full plan compiled meanwhile (_single_run); PNMN_SPLIT_COMPILE=0: one plan, inline
        self.split_compile = os.environ.get("PNMN_SPLIT_COMPILE", "1") != "0"
        self.backward_stats_hook = None   # optional callable(plan stats) invoked when a backward pass is launched
        self._pack_table: Optional[torch.Tensor] = None
        # parity tests: keep every 1-channel module output (attention map) of the last forward, see _read_attention_maps
        self.capture_attention_maps = False
        self.last_attention_maps = None
        # classifier products (nmn.py:75-83): "split" (default) = pnmn_gemm_split, the repo's tcgen05 GEMM over bf16 (hi, lo)
        # split operands (~16 mantissa bits per operand, fp32 accumulation).  "ieee" / "tf32" run the plain nn.Sequential on
        # cuBLAS / cuDNN instead -- comparison runs only (scripts/classifier_bench.py): IEEE is 3.6x slower, TF32 misses the
        # gradient parity bar
        self.classifier_math = os.environ.get("PNMN_CLASSIFIER", "tf32" if os.environ.get("PNMN_CLASSIFIER_TF32") == "1" else "split")
        if self.classifier_math not in ("split", "ieee", "tf32"):
            raise ValueError("PNMN_CLASSIFIER must be split, ieee or tf32")
        self.classifier_tf32 = self.classifier_math == "tf32"

    @classmethod
    def from_config(cls, config):
        r"""Instantiate from a reference-style ``Config`` (duck-typed: ``_C.DATA.VOCABULARY``, ``_C.NMN.*``;
        nmn.py:126-137)."""
        _C = config
        return cls(
            vocabulary=Vocabulary.from_files(_C.DATA.VOCABULARY),
            image_feature_size=tuple(_C.NMN.IMAGE_FEATURE_SIZE),
            module_channels=_C.NMN.MODULE_CHANNELS,
            class_projection_channels=_C.NMN.CLASS_PROJECTION_CHANNELS,
            classifier_linear_size=_C.NMN.CLASSIFIER_LINEAR_SIZE,
        )

    # ---- flat parameter buffer ---------------------------------------------------------------------
    def _exec_named_parameters(self):
        """stem + module parameters, i.e. everything the CUDA executor reads (classifier excluded)."""
        return [(n, p) for n, p in self.named_parameters() if not n.startswith("classifier.")]

    def _ensure_flat(self):
        """All executor parameters are views into ONE flat fp32 buffer (reference names / shapes kept).  The check that
        they still are is O(1) per forward (first / last parameter); ``.to()`` / ``.cuda()`` re-create every ``.data``
        and are caught by it."""
        cached = self._exec_params
        if cached is not None and self._flat is not None:
            first, last = cached[0], cached[-1]
            base, (_, lo, _, _), (_, hi, _, _) = self._flat.data_ptr(), self._layout[0], self._layout[-1]
            if first.data_ptr() == base + 4 * lo and last.data_ptr() == base + 4 * hi and first.device == self._flat.device:
                return
        named = self._exec_named_parameters()
        dev = named[0][1].device
        layout, off = [], 0
        for name, p in named:
            layout.append((name, off, p.numel(), p.shape))
            off += (p.numel() + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for (name, o, n, shape), (_, p) in zip(layout, named):
            flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = flat[o:o + n].view(shape)
        self._flat, self._layout = flat, layout
        self._exec_params = [p for _, p in named]
        self._gflat, self._gviews = None, None
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._packed = None
        if self._model_handle is None:
            self._model_handle = self._create_model_handle()

    def _attach_grads(self):
        """Returns (flat buffer the CUDA backward accumulates into, optional fix-up callable).  Fast paths: every
        ``.grad`` is None (after ``zero_grad(set_to_none=True)``: zero the persistent buffer, attach its views) or every
        ``.grad`` already IS its view (``zero_grad(set_to_none=False)`` / gradient accumulation: add in place).  Anything
        else (foreign gradient tensors) goes through a temporary buffer."""
        params = self._exec_params
        if self._gflat is None or self._gflat.device != self._flat.device:
            self._gflat = torch.zeros_like(self._flat)
            self._gviews = [self._gflat[o:o + n].view(shape) for _, o, n, shape in self._layout]
        views = self._gviews
        self._gflat_box["gflat"] = self._gflat
        if all(p.grad is v for p, v in zip(params, views) if p.requires_grad):
            return self._gflat, None  # (the common case first: one pass over the 222 parameters)
        if all(p.grad is None for p in params):
            self._gflat.zero_()
            for p, v in zip(params, views):
                if p.requires_grad:
                    p.grad = v
            return self._gflat, None
        tmp = torch.zeros_like(self._flat)

        def finish():
            for p, (_, o, n, shape) in zip(params, self._layout):
                if not p.requires_grad:
                    continue
                g = tmp[o:o + n].view(shape)
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad = p.grad + g
        return tmp, finish

    def zero_grad(self, set_to_none: bool = True) -> None:
        """``nn.Module.zero_grad``.  When the executor parameters' gradients are the views of the flat gradient buffer (the
        state every backward pass leaves), they are cleared with ONE memset and stay attached -- zero tensors, which is what
        the reference's ``optimizer.zero_grad()`` produces under its pinned torch 1.4 (``_trainer.py:193``) -- instead of
        222 attribute writes here and 222 re-attachments in the next backward (~0.5 ms of host time per step)."""
        params, views = self._exec_params, self._gviews
        if (set_to_none and params and views is not None and self._gflat is not None
                and all(p.grad is v for p, v in zip(params, views) if p.requires_grad)):
            self._gflat.zero_()
            for p in self.classifier.parameters():
                p.grad = None
            return
        super().zero_grad(set_to_none=set_to_none)

    def _create_model_handle(self):
        offs = {name: o for name, o, _, _ in self._layout}
        tokens = self.vocabulary.get_index_to_token_vocabulary(namespace="programs")
        V = max(tokens.keys()) + 1
        kinds = (ctypes.c_int32 * V)()
        table = (ctypes.c_int64 * (V * L.MAX_MODULE_PARAMS))(*([-1] * (V * L.MAX_MODULE_PARAMS)))
        for idx, token in tokens.items():
            family = module_class_of(token)
            kinds[idx] = _KIND[family]
            if family in ("attention", "query", "relate", "same", "comparison"):
                for i, (name, *_rest) in enumerate(_family_spec(family, 128)):
                    table[idx * L.MAX_MODULE_PARAMS + 2 * i] = offs[f"{token}.{name}.weight"]
                    table[idx * L.MAX_MODULE_PARAMS + 2 * i + 1] = offs[f"{token}.{name}.bias"]
        stem = (ctypes.c_int64 * 4)(offs["stem.0.weight"], offs["stem.0.bias"], offs["stem.2.weight"], offs["stem.2.bias"])
        handle = L.lib().pnmn_model_create(V, kinds, table, stem, self._in_channels)
        if not handle:
            raise RuntimeError("pnmn_model_create failed: " + L.lib().pnmn_last_error().decode())
        return handle

    def __del__(self):
        try:
            if self._model_handle is not None:
                L.lib().pnmn_model_destroy(self._model_handle)
        except Exception:
            pass

    # ---- forward -------------------------------------------------------------------------------------
    def forward(self, features: torch.Tensor, programs: torch.Tensor, answers: Optional[torch.Tensor] = None):
        r"""
        Same contract as the reference (nmn.py:139-275): ``features`` (B, C, 14, 14) float, ``programs``
        (B, L) prefix-order token ids, optional ``answers`` (B,).  ``features`` may also be ``torch.float16`` when they come
        from ``feed.ImageFeatureCache`` (same results, half the bytes).  Returns ``{"predictions": (B,) int64,
        "loss": (B,) float32}`` plus ``"metrics"`` in training mode.  A program the reference could not
        execute never raises: its prediction is ``@@UNKNOWN@@`` and its loss the constant 3.33.
        """
        if not features.is_cuda:
            raise RuntimeError("NeuralModuleNetwork (B200) needs CUDA tensors; there is no CPU fallback")
        # the library launches on the calling thread's current device: make it the tensors' device (the reference's
        # trainer only does ``model.to(f"cuda:{gpu_ids[0]}")``, trainers/_trainer.py:92-95, never set_device)
        with torch.cuda.device(features.device):
            return self._forward(features, programs, answers)

    def _forward(self, features: torch.Tensor, programs: torch.Tensor, answers: Optional[torch.Tensor]):
        lib = L.lib()
        self._ensure_flat()
        # fp16 features are taken as they are: they come from a feature cache that already holds the executor's operand
        # values (feed.ImageFeatureCache / pnmn_round_features_f16); anything else is read as fp32 like the reference does
        features = features.contiguous() if features.dtype == torch.float16 else features.contiguous().float()
        B, Lp = programs.shape
        # The program compiler runs on the host: programs that are already host tensors cost no synchronisation
        # (the reference accepts them too, it calls ``programs[n].cpu()``, nmn.py:203); device tensors cost the
        # forward's single D2H copy.
        handed = getattr(programs, "_pnmn_host", None)
        if handed is not None and programs.is_cuda:
            # produced by a Seq2SeqBase forward on this device: its pinned host copy is already on its way (seq2seq._handover);
            # wait for THAT copy only instead of synchronising the compute stream (which may be busy with later work)
            handed[1].synchronize()
            programs_host = handed[0].clone()
        else:
            programs_host = programs.detach().to("cpu", torch.int64).contiguous()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._exec_params)
        # a prestage() of exactly these features (same memory, same weights buffer): skip the pack / layout kernels
        staged, self._prestaged = self._prestaged, None
        if staged is not None and not (staged[0].data_ptr() == features.data_ptr() and staged[0].shape == features.shape
                                       and staged[0].dtype == features.dtype and staged[2] == self._flat.data_ptr()):
            _POOL.release(staged[1])
            staged = None
        k = self.compile_chunks
        chunked = (k > 1 and B >= 16 * k and not self.capture_attention_maps
                   and not (staged is None and self._precompiled))
        if chunked:
            runs, valid_host, stats = self._chunked_runs(features, programs_host, need_grad, staged, k)
        else:
            runs, valid_host, stats = self._single_run(features, programs_host, need_grad, staged)
        self.last_plan_stats = stats
        if need_grad:
            final = _ExecutorFn.apply(features, self._anchor, runs, self)
        else:
            with torch.no_grad():
                final = _ExecutorFn.forward(_NullCtx(), features, None, runs, self)
        # The validity mask is read on the device from the plans' task tables (per-sample stem-input offset, < 0 = invalid)
        # rather than uploaded -- a host -> device copy queued here would wait behind whatever the copy engine is doing (the
        # next batch's 205 MB of features in a pipelined loop: 3.7 ms) and stall the whole stream.  The answer head reads it
        # BEFORE a run is closed: closing hands a pre-uploaded table buffer back to the pool, where a look-ahead compile may
        # overwrite it.
        if self.capture_attention_maps:
            self.last_attention_maps = self._read_attention_maps(runs[0].plan, runs[0].ws)

        # classifier (nmn.py:241-244)
        if self.classifier_math == "split":
            answer_logits = self._classifier_split(final)
        elif self.classifier_tf32:
            prev = torch.backends.cuda.matmul.allow_tf32
            final = _MatmulPrecision.apply(final, True, prev)
            answer_logits = _MatmulPrecision.apply(self.classifier(final), prev, True)
        else:
            answer_logits = self.classifier(final)
        # answer head (nmn.py:245-269): predictions, masking of invalid programs, per-row loss, correct count -- one kernel
        correct = torch.zeros((), dtype=torch.int64, device=features.device) if answers is not None else None
        tables = [(run.blob, run.xin_off, run.lo, run.hi) for run in runs]
        answer_predictions, loss = _AnswerLoss.apply(answer_logits, answers, tables, self._unknown_answer, correct)
        if not need_grad:
            for run in runs:
                run.close()
        if answers is not None:
            # the correct count stays on the device until a metric is read: no synchronisation inside forward
            self._answer_accuracy(correct, B)
            self._average_invalid_programs(int((valid_host == 0).sum()))

        output_dict = {"predictions": answer_predictions, "loss": loss}
        if self.training:
            # same values as the reference's ``self.get_metrics(reset=True)`` (nmn.py:274): the accumulators are
            # snapshotted and reset NOW, the numbers are produced when the dict is first read
            acc_state = self._answer_accuracy.snapshot(reset=True)
            invalid_now = self._average_invalid_programs.get_metric(reset=True)
            output_dict["metrics"] = _LazyMetrics({"answer_accuracy": lambda: _Accuracy.value(acc_state),
                                                   "average_invalid": lambda: invalid_now})
        return output_dict

    def _plan_info(self, plan, rows):
        lib = L.lib()
        valid = torch.empty(rows, dtype=torch.uint8)
        lib.pnmn_plan_valid(plan, ctypes.cast(valid.data_ptr(), ctypes.POINTER(ctypes.c_uint8)))
        sizes = (ctypes.c_int64 * L.SZ_COUNT)()
        lib.pnmn_plan_sizes(plan, sizes)
        stats = (ctypes.c_int64 * 16)()
        lib.pnmn_plan_stats(plan, stats)
        return valid, sizes, list(stats)

    def _single_run(self, features, programs_host, need_grad, staged):
        """One plan for the whole batch, on the caller's stream."""
        lib = L.lib()
        B = programs_host.shape[0]
        pre = self._take_precompiled(programs_host, need_grad) if staged is None else None
        pre_blob = pre_event = None
        full = None
        if pre is None:
            if need_grad and self.split_compile:
                # The programs have only just become known and the executor waits for this thread: compile the forward
                # half alone (a third of the time), let a helper thread compile + upload the full plan while the forward
                # pass runs, and take the backward pass from that one (PNMN_PLAN_FORWARD_HALF, include/pnmn.h).
                dev = features.device
                if self._upload_stream is None or self._upload_stream.device != dev:
                    self._upload_stream = torch.cuda.Stream(dev)
                # (the helper gets its own copy: `programs_host` may alias a host tensor the caller goes on to modify)
                full = _compile_pool().submit(self._compile_and_upload, programs_host.clone(), True, dev, staged is not None)
                plan = self._compile(programs_host, False, None, by_row=staged is not None, forward_half=True)
            else:
                plan = self._compile(programs_host, need_grad, None, by_row=staged is not None)
        else:
            plan, pre_blob, pre_event = pre
        valid_host, sizes, stats = self._plan_info(plan, B)
        ws = staged[1] if staged is not None else _POOL.acquire(features.device)
        ws.ensure(sizes)
        if self._packed is None or self._packed.device != features.device:
            self._packed = torch.empty(lib.pnmn_model_packed_floats(self._model_handle), dtype=torch.float32,
                                       device=features.device)
        blob = ws.t["blob"]
        if pre_blob is not None and pre_blob.device == features.device:
            # the look-ahead compile already uploaded the task tables on its own stream (pnmn_plan_upload)
            current = torch.cuda.current_stream(features.device)
            current.wait_event(pre_event)
            blob = pre_blob
        bufs = L.Buffers(ws.t["arena16"].data_ptr(), ws.t["arena18"].data_ptr(), ws.t["arena22"].data_ptr(),
                         ws.t["maps"].data_ptr(), ws.t["dmaps"].data_ptr(), ws.t["idx"].data_ptr(),
                         blob.data_ptr(), self._packed.data_ptr(), self._flat.data_ptr(), None,
                         ws.t["ain"].data_ptr(), ws.scratch.data_ptr())
        run = _Run(plan, ws, bufs, self._gflat_box, lo=0, hi=B, staged=staged is not None)
        run.blob = blob
        run.xin_off = int(stats[15])
        run.pool_blob = pre_blob  # goes back to the pool when the run is closed (after the backward pass)
        run.full = full
        return [run], valid_host, stats

    def _chunked_runs(self, features, programs_host, need_grad, staged, k):
        """The batch as ``k`` independent plans over contiguous row ranges, compiled on ``k`` host threads (ctypes releases
        the GIL around the C++ compiler) and executed side by side on ``k`` streams, each with 1/k of the executor's CTA slots.
        What it buys is TIME TO LAUNCH: the compiler is sequential per plan (~12 us per program), and when the programs only
        become known inside the step (joint training: they are sampled by the generator, modules/elbo.py:230-239) that time
        sits on the step's critical path; the executor itself is bound by the length of its dependency chains, not by the
        number of rows, so k smaller passes in parallel take about as long as one.  Weight packing and feature layout are
        shared (``prestage``); results are those of the single plan up to the summation order of the weight gradients."""
        lib = L.lib()
        dev = features.device
        B = programs_host.shape[0]
        if staged is None:
            self.prestage(features)
            staged, self._prestaged = self._prestaged, None
        ws_main = staged[1]
        bounds = [B * i // k for i in range(k + 1)]
        pool = _compile_pool()
        futures = [pool.submit(self._compile, programs_host[lo:hi], need_grad, dev, True)
                   for lo, hi in zip(bounds[:-1], bounds[1:])]
        if self._chunk_streams is None or len(self._chunk_streams) < k or self._chunk_streams[0].device != dev:
            self._chunk_streams = [torch.cuda.Stream(dev) for _ in range(k)]
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        unit_bytes = self._in_channels // 4 * 256 * 16 * 3 // 2
        runs, valids, total = [], [], [0] * 16
        for i, fut in enumerate(futures):
            lo, hi = bounds[i], bounds[i + 1]
            plan = fut.result()
            lib.pnmn_plan_set_exec_ctas(plan, max(2 * sms // k, 16))
            valid, sizes, stats = self._plan_info(plan, hi - lo)
            sizes[L.SZ_AIN] = 0                      # the stem inputs of every chunk live in the prestaged workspace
            ws = _POOL.acquire(dev)
            ws.ensure(sizes)
            bufs = L.Buffers(ws.t["arena16"].data_ptr(), ws.t["arena18"].data_ptr(), ws.t["arena22"].data_ptr(),
                             ws.t["maps"].data_ptr(), ws.t["dmaps"].data_ptr(), ws.t["idx"].data_ptr(),
                             ws.t["blob"].data_ptr(), self._packed.data_ptr(), self._flat.data_ptr(), None,
                             ws_main.t["ain"].data_ptr() + lo * unit_bytes, ws.scratch.data_ptr())
            run = _Run(plan, ws, bufs, self._gflat_box, lo=lo, hi=hi, stream=self._chunk_streams[i], staged=True,
                       extra_ws=ws_main if i == 0 else None)
            run.blob = ws.t["blob"]
            run.xin_off = int(stats[15])
            run.pool_blob = None
            runs.append(run)
            valids.append(valid)
            total = [a + b for a, b in zip(total, stats)]
        return runs, torch.cat(valids), total

    @staticmethod
    def _read_attention_maps(plan, ws):
        """[(sample, module call index, token id, (14, 14) tensor)] for every 1-channel module output of the forward pass
        just launched (nmn_modules.py:82-87,160-168,200-208 and 1-channel And / Or results), read back from the map arena
        through ``pnmn_debug_plan_maps``.  The module call index counts the modules the reference would call for that
        sample in execution order (``scene`` and skipped tokens call none)."""
        lib = L.lib()
        n = int(lib.pnmn_debug_plan_maps(plan, None, 0))
        rec = torch.zeros(max(n, 1), 4, dtype=torch.int32)
        lib.pnmn_debug_plan_maps(plan, ctypes.c_void_p(rec.data_ptr()), n)
        arena = ws.t["maps"]
        grid = arena[: arena.numel() // 256 * 256].view(-1, 16, 16)
        units = rec[:n, 3].to(torch.int64)
        maps = grid[units.to(grid.device)][:, :14, :14].cpu() if n else torch.zeros(0, 14, 14)
        return [(int(rec[i, 0]), int(rec[i, 1]), int(rec[i, 2]), maps[i]) for i in range(n)]

    # ---- program compiler -------------------------------------------------------------------------------------------
    def _compile(self, programs_host: torch.Tensor, need_grad: bool, device, by_row: bool = False, forward_half: bool = False):
        lib = L.lib()
        B, Lp = programs_host.shape
        ptr = ctypes.cast(programs_host.data_ptr(), ctypes.POINTER(ctypes.c_int64))
        flags = (L.PLAN_INPUT_BY_ROW if by_row else 0) | (L.PLAN_FORWARD_HALF if forward_half else 0)
        if device is not None:  # helper thread: the pinned staging buffers and their events belong to this device
            with torch.cuda.device(device):
                plan = lib.pnmn_plan_create_ex(self._model_handle, ptr, B, Lp, 1 if need_grad else 0, flags)
        else:
            plan = lib.pnmn_plan_create_ex(self._model_handle, ptr, B, Lp, 1 if need_grad else 0, flags)
        if not plan:
            raise RuntimeError("pnmn_plan_create failed: " + lib.pnmn_last_error().decode())
        return plan

    def _compile_and_upload(self, programs_host: torch.Tensor, need_grad: bool, device, by_row: bool = False):
        """Helper-thread half of ``precompile``: compile, then copy the task tables to the device on the upload stream, so
        that the forward pass itself issues no host -> device copy (one queued inside forward would wait for the compute
        stream and then find the copy engine busy with the next batch's features)."""
        plan = self._compile(programs_host, need_grad, device, by_row=by_row)
        if device is None or os.environ.get("PNMN_NO_EARLY_UPLOAD"):  # (switch: comparison runs)
            return plan, None, None
        lib = L.lib()
        try:
            sizes = (ctypes.c_int64 * L.SZ_COUNT)()
            lib.pnmn_plan_sizes(plan, sizes)
            with torch.cuda.device(device), torch.cuda.stream(self._upload_stream):
                blob = _BLOBS.acquire(int(sizes[L.SZ_BLOB]), device, self._upload_stream)
                L.check(lib.pnmn_plan_upload(plan, ctypes.c_void_p(blob.data_ptr()),
                                             ctypes.c_void_p(self._upload_stream.cuda_stream)), "pnmn_plan_upload")
                event = torch.cuda.Event()
                event.record(self._upload_stream)
            return plan, blob, event
        except Exception:
            lib.pnmn_plan_destroy(plan)
            raise

    def prestage(self, features: torch.Tensor) -> None:
        """Optional: run the part of ``forward`` that does not depend on the programs -- packing the weights into MMA tiles
        and laying out the image features -- NOW, on the current stream.  In the joint-training step the programs are
        sampled by the program generator's forward pass (modules/elbo.py:230-239) while the features and the weights are
        known from the start: the ``forward`` that follows (same ``features`` tensor) then goes straight from the program
        compiler to the executor.  The reference has no counterpart; results are identical with or without it."""
        if not features.is_cuda:
            raise RuntimeError("NeuralModuleNetwork (B200) needs CUDA tensors; there is no CPU fallback")
        self._drop_prestaged()
        with torch.cuda.device(features.device):
            self._ensure_flat()
            lib = L.lib()
            dev = features.device
            feats = features.contiguous() if features.dtype == torch.float16 else features.contiguous().float()
            B = feats.shape[0]
            if B == 0:
                return
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty(lib.pnmn_model_packed_floats(self._model_handle), dtype=torch.float32, device=dev)
            if self._pack_table is None or self._pack_table.device != dev:
                host = torch.empty(max(int(lib.pnmn_model_pack_table_bytes(self._model_handle)), 1), dtype=torch.uint8)
                lib.pnmn_model_pack_table(self._model_handle, ctypes.c_void_p(host.data_ptr()))
                self._pack_table = host.to(dev)
            ws = _POOL.acquire(dev)
            need = int(lib.pnmn_model_ain_floats(self._model_handle, B))
            cur = ws.t.get("ain")
            if cur is None or cur.numel() < need:
                ws.t["ain"] = None
                ws.t["ain"] = torch.zeros(int(need * 1.25) + 1024, dtype=torch.float32, device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.pnmn_nmn_prestage(self._model_handle, ctypes.c_void_p(self._pack_table.data_ptr()),
                                          ctypes.c_void_p(self._flat.data_ptr()), ctypes.c_void_p(self._packed.data_ptr()),
                                          ctypes.c_void_p(feats.data_ptr()), 1 if feats.dtype == torch.float16 else 0,
                                          ctypes.c_void_p(ws.t["ain"].data_ptr()), B, stream), "pnmn_nmn_prestage")
            self._prestaged = (feats, ws, self._flat.data_ptr())

    def _drop_prestaged(self) -> None:
        if self._prestaged is not None:
            _POOL.release(self._prestaged[1])
            self._prestaged = None

    def precompile(self, programs: torch.Tensor, need_grad: Optional[bool] = None) -> None:
        """Optional look-ahead for input pipelines: start compiling ``programs`` (host tensor, (B, L) token ids) into an
        executor plan on a helper thread.  A later ``forward`` whose programs have the same contents picks the plan up
        instead of compiling inline (2-4 ms of host time per 256 programs).  At most three plans wait at a time (the oldest
        is dropped), so a pipeline can submit batches i+1 and i+2 before it runs batch i.  The reference has no counterpart -- its
        interpreter walks the programs inside forward (nmn.py:191-238) -- and results are identical with or without it."""
        self._ensure_flat()
        if programs.device.type != "cpu":
            raise ValueError("precompile needs the programs in host memory (that is where the compiler runs)")
        host = programs.detach().to(torch.int64).contiguous().clone()
        if need_grad is None:
            need_grad = self.training and any(p.requires_grad for p in self._exec_params)
        device = self._flat.device if self._flat is not None and self._flat.is_cuda else None
        if device is not None and (self._upload_stream is None or self._upload_stream.device != device):
            self._upload_stream = torch.cuda.Stream(device)
        while len(self._precompiled) >= 3:
            self._destroy_pending(self._precompiled.pop(0))
        self._precompiled.append((host, bool(need_grad),
                                  _compile_pool().submit(self._compile_and_upload, host, bool(need_grad), device)))

    def _take_precompiled(self, programs_host: torch.Tensor, need_grad: bool):
        for k, (host, ng, fut) in enumerate(self._precompiled):
            if ng == need_grad and host.shape == programs_host.shape and torch.equal(host, programs_host):
                del self._precompiled[k]
                return fut.result()
        return None

    @staticmethod
    def _destroy_pending(entry) -> None:
        try:
            plan, blob, _ = entry[2].result()
            L.lib().pnmn_plan_destroy(plan)
            if blob is not None:
                _BLOBS.release(blob)
        except Exception:
            pass

    def _drop_precompiled(self) -> None:
        while self._precompiled:
            self._destroy_pending(self._precompiled.pop())

    def _classifier_split(self, final: torch.Tensor) -> torch.Tensor:
        """nmn.py:75-83 on the repo's own kernels: every product (1x1 conv 128->1024 over the B*196 pixels, Linear
        50176->1024, Linear 1024->num_answers, and their gradients) is ``pnmn_gemm_split``; ReLU / MaxPool2d(2,2) / flatten
        between the first two is one pass over the conv output."""
        conv, fc1, fc2 = self.classifier[0], self.classifier[4], self.classifier[6]
        B, C, H, W = final.shape
        x = final.permute(0, 2, 3, 1).reshape(B * H * W, C)
        if H == 14 and W == 14 and conv.out_channels % 64 == 0:
            y = _ConvReluPool.apply(x, conv.weight.view(conv.out_channels, C), conv.bias, B)
        else:
            y = _SplitLinear.apply(x, conv.weight.view(conv.out_channels, C), conv.bias)   # channels-last, pre-activation
            y = F.relu(y).view(B, H, W, conv.out_channels).permute(0, 3, 1, 2)  # NCHW view of channels-last data
            y = F.max_pool2d(y, kernel_size=2, stride=2)
            y = y.contiguous(memory_format=torch.contiguous_format).reshape(B, -1)  # (C, 7, 7) order, nmn_modules.py:250
        z = F.relu(_SplitLinear.apply(y, fc1.weight, fc1.bias))
        logits = _SplitLinear.apply(z, fc2.weight, fc2.bias)
        for hook in self.classifier._forward_hooks.values():  # tests / tools observe the classifier through hooks
            hook(self.classifier, (final,), logits)
        return logits

    def allreduce_gradients(self, group=None) -> None:
        """Data-parallel gradient averaging over NCCL (one process per GPU): ONE all-reduce for the stem + module
        gradients, which the executor's backward leaves in a single flat buffer, plus one per classifier tensor.
        (The reference has no distributed path: it wraps the model in ``nn.DataParallel``,
        trainers/_trainer.py:98-100, which mis-executes the NMN)"""
        from .dist import allreduce_gradients

        if self._grad_overlap is not None:
            # every classifier gradient is already travelling (hooks); what is left is the executor's flat gradient buffer
            # when the parameters' .grad are its views (the normal case), else whatever the parameter walk finds
            params = self._exec_params or []
            flat_ok = (self._gflat is not None and self._gviews is not None and len(params) > 0
                       and params[0].grad is self._gviews[0] and params[-1].grad is self._gviews[-1])
            if flat_ok:
                self._grad_overlap.finish(buckets=[self._gflat])
            else:
                self._grad_overlap.finish([self])
        else:
            allreduce_gradients([self], group=group)

    def enable_gradient_overlap(self, group=None) -> None:
        """Start the all-reduce of each classifier gradient as soon as autograd has produced it (they are 206 of the
        257 MB a step reduces and come first in the backward pass), underneath the module executor's backward;
        ``allreduce_gradients`` then only waits for them and reduces the executor's flat gradient buffer."""
        from .dist import GradientOverlap

        if self._grad_overlap is not None:
            self._grad_overlap.remove()
        # classifier.4.weight (51 M elements) gets its own early collective; the five small tensors share one at the end
        self._grad_overlap = GradientOverlap(self.classifier.parameters(), group=group, min_numel=1 << 20)

    def get_metrics(self, reset: bool = True) -> Dict[str, float]:
        """``{"answer_accuracy", "average_invalid"}`` (nmn.py:277-296)."""
        return {
            "answer_accuracy": self._answer_accuracy.get_metric(reset=reset),
            "average_invalid": self._average_invalid_programs.get_metric(reset=reset),
        }


class _NullCtx:
    pass


class _AnswerLoss(torch.autograd.Function):
    """(predictions, loss) = answer head over the classifier's logits (``pnmn_answer_loss_forward`` / ``_backward``,
    csrc/loss.cu; nmn.py:245-269).  ``blob`` / ``xin_off``: the plan's task-table buffer and the byte offset of its per-row
    validity table -- ``tables`` lists (blob, offset, first row, end row) per plan; ``correct``: optional device int64 scalar
    that receives the number of correct predictions."""

    @staticmethod
    def forward(ctx, logits, answers, tables, unknown, correct):
        logits = logits.contiguous().float()
        B, A = logits.shape
        dev = logits.device
        if answers is not None:
            answers = answers.detach().to(dev, torch.int64).contiguous()
        predictions = torch.empty(B, dtype=torch.int64, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        invalid = torch.empty(B, dtype=torch.uint8, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for blob, xin_off, lo, hi in tables:       # one validity table per plan (a batch may be compiled in chunks of rows)
            if hi <= lo:
                continue
            L.check(L.lib().pnmn_answer_loss_forward(
                ctypes.c_void_p(logits.data_ptr() + 4 * A * lo),
                ctypes.c_void_p(answers.data_ptr() + 8 * lo) if answers is not None else None,
                ctypes.c_void_p(blob.data_ptr() + xin_off), hi - lo, A, int(unknown), ctypes.c_void_p(predictions.data_ptr() + 8 * lo),
                ctypes.c_void_p(loss.data_ptr() + 4 * lo), ctypes.c_void_p(invalid.data_ptr() + lo),
                ctypes.c_void_p(correct.data_ptr()) if correct is not None else None, stream), "pnmn_answer_loss_forward")
        ctx.save_for_backward(logits, predictions, invalid)
        ctx.answers = answers
        ctx.mark_non_differentiable(predictions)
        return predictions, loss

    @staticmethod
    def backward(ctx, _grad_predictions, grad_loss):
        logits, predictions, invalid = ctx.saved_tensors
        answers = ctx.answers
        B, A = logits.shape
        grad_loss = grad_loss.contiguous().float()
        dlogits = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
            L.check(L.lib().pnmn_answer_loss_backward(
                ctypes.c_void_p(logits.data_ptr()), ctypes.c_void_p(answers.data_ptr()) if answers is not None else None,
                ctypes.c_void_p(invalid.data_ptr()), ctypes.c_void_p(predictions.data_ptr()),
                ctypes.c_void_p(grad_loss.data_ptr()), B, A, ctypes.c_void_p(dlogits.data_ptr()), stream),
                "pnmn_answer_loss_backward")
        return dlogits, None, None, None, None


_COMPILE_POOL = None


def _compile_pool():
    """Two helper threads for NeuralModuleNetwork.precompile (ctypes releases the GIL around the C++ compiler): a pipeline
    that knows its batches two steps ahead keeps two plans in flight, so a compile has two step times to finish."""
    global _COMPILE_POOL
    if _COMPILE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COMPILE_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="pnmn-plan")
    return _COMPILE_POOL


def _gemm(A: torch.Tensor, a_rs: int, a_ks: int, Bm: torch.Tensor, b_rs: int, b_ks: int, M: int, N: int, K: int,
          bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C (M, N) = sum_k A(m, k) * B(n, k) (+ bias) with A(m, k) = A.flat[m*a_rs + k*a_ks], B likewise: ``pnmn_gemm_split``
    (csrc/gemm.cu) -- fp32 in / out, bf16 (hi, lo) split products on the tensor cores, operands read in their home layout."""
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    lib = L.lib()
    with torch.cuda.device(A.device):
        ws_floats = int(lib.pnmn_gemm_split_workspace(M, N, K))   # > 0: the contraction is split; partial tiles are added in order
        ws = torch.empty(ws_floats, dtype=torch.float32, device=A.device) if ws_floats else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)
        L.check(lib.pnmn_gemm_split(ctypes.c_void_p(A.data_ptr()), a_rs, a_ks, ctypes.c_void_p(Bm.data_ptr()), b_rs, b_ks,
                                    ctypes.c_void_p(out.data_ptr()), N, M, N, K,
                                    ctypes.c_void_p(bias.data_ptr()) if bias is not None else None, 0,
                                    ctypes.c_void_p(ws.data_ptr()) if ws is not None else None, ws_floats, stream), "pnmn_gemm_split")
    return out


class _SplitLinear(torch.autograd.Function):
    """y = x @ w.T + b (x: (M, K), w: (N, K)) and its three gradient products, all through ``pnmn_gemm_split``: the weight --
    205 MB for classifier[4] -- is read in place by every product (w as stored for y and dw, w with swapped strides for dx),
    no split or transposed copy of it is ever written."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = x.contiguous(), w.contiguous()
        ctx.save_for_backward(x, w)
        M, K = x.shape
        return _gemm(x, K, 1, w, K, 1, M, w.shape[0], K, bias=b.contiguous() if b is not None else None)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        M, K = x.shape
        N = w.shape[0]
        dx = dw = db = None
        with torch.cuda.device(g.device):
            if ctx.needs_input_grad[0]:
                dx = _gemm(g, N, 1, w, 1, K, M, K, N)          # dx[m][k] = sum_n g[m][n] * w[n][k]
            if ctx.needs_input_grad[1]:
                dw = _gemm(g, 1, N, x, 1, K, N, K, M)          # dw[n][k] = sum_m g[m][n] * x[m][k]
            if ctx.needs_input_grad[2]:
                db = g.sum(0)
        return dx, dw, db


class _ConvReluPool(torch.autograd.Function):
    """classifier[0:4] (nmn.py:75-79): 1x1 conv as a GEMM over the B*196 pixels + ReLU + MaxPool2d(2,2) + flatten, as ONE
    autograd node: ``pnmn_gemm_split`` for the products, ``pnmn_relu_pool_fwd_bias`` / ``pnmn_relu_pool_bwd`` for the pooling
    pass (the conv bias is added inside it: max(v) + b == max(v + b))."""

    @staticmethod
    def forward(ctx, x, w, b, B):
        x, w = x.contiguous(), w.contiguous()
        M, K = x.shape
        C = w.shape[0]
        y = _gemm(x, K, 1, w, K, 1, M, C, K)
        pooled = torch.empty((B, C * 49), dtype=torch.float32, device=y.device)
        code = torch.empty((B, C * 49), dtype=torch.uint8, device=y.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
        L.check(L.lib().pnmn_relu_pool_fwd_bias(ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(b.contiguous().data_ptr()),
                                                ctypes.c_void_p(pooled.data_ptr()), ctypes.c_void_p(code.data_ptr()), B, C,
                                                stream), "pnmn_relu_pool_fwd_bias")
        ctx.save_for_backward(x, w, code)
        ctx.B = B
        return pooled

    @staticmethod
    def backward(ctx, g):
        x, w, code = ctx.saved_tensors
        B, M, K, C = ctx.B, x.shape[0], x.shape[1], w.shape[0]
        g = g.contiguous()
        dx = dw = db = None
        with torch.cuda.device(g.device):
            gy = torch.empty((M, C), dtype=torch.float32, device=g.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
            L.check(L.lib().pnmn_relu_pool_bwd(ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(code.data_ptr()),
                                               ctypes.c_void_p(gy.data_ptr()), B, C, stream), "pnmn_relu_pool_bwd")
            if ctx.needs_input_grad[0]:
                dx = _gemm(gy, C, 1, w, 1, K, M, K, C)
            if ctx.needs_input_grad[1]:
                dw = _gemm(gy, 1, C, x, 1, K, C, K, M)
            if ctx.needs_input_grad[2]:
                # a pooled gradient reaches the bias when its window's maximum was positive (bit 2 of the code)
                db = (g.view(B, C, 49) * ((code.view(B, C, 49) & 4) != 0)).sum((0, 2))
        return dx, dw, db, None


class _ReluPoolFlatten(torch.autograd.Function):
    """ReLU -> MaxPool2d(2, 2) -> flatten of the channels-last 1x1-conv output [B*196, C] (nmn.py:77-79; flatten order
    (C, 7, 7), nmn_modules.py:250-251) as ONE pass over the 205 MB tensor; the backward routes each pooled gradient to the
    first maximum of its window (ATen's tie rule) if that maximum was positive (``pnmn_relu_pool_fwd`` / ``_bwd``)."""

    @staticmethod
    def forward(ctx, y, B):
        y = y.contiguous()
        C = y.shape[1]
        pooled = torch.empty((B, C * 49), dtype=torch.float32, device=y.device)
        code = torch.empty((B, C * 49), dtype=torch.uint8, device=y.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
        L.check(L.lib().pnmn_relu_pool_fwd(ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(pooled.data_ptr()),
                                           ctypes.c_void_p(code.data_ptr()), B, C, stream), "pnmn_relu_pool_fwd")
        ctx.save_for_backward(code)
        ctx.shape = (B, C)
        return pooled

    @staticmethod
    def backward(ctx, g):
        (code,) = ctx.saved_tensors
        B, C = ctx.shape
        g = g.contiguous()
        gy = torch.empty((B * 196, C), dtype=torch.float32, device=g.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        L.check(L.lib().pnmn_relu_pool_bwd(ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(code.data_ptr()),
                                           ctypes.c_void_p(gy.data_ptr()), B, C, stream), "pnmn_relu_pool_bwd")
        return gy, None


class _MatmulPrecision(torch.autograd.Function):
    """Identity that switches cuBLAS/cuDNN fp32 math between IEEE and TF32 tensor cores.  Two of them bracket
    the classifier (plain library GEMMs: 1x1 conv, Linear 50176->1024, Linear 1024->28, nmn.py:75-83); autograd
    replays them in reverse order, so the classifier's backward GEMMs run under the same setting."""

    @staticmethod
    def forward(ctx, x, on_forward, on_backward):
        ctx.on_backward = on_backward
        _MatmulPrecision.set(on_forward)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        _MatmulPrecision.set(ctx.on_backward)
        return g, None, None

    @staticmethod
    def set(tf32: bool):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
