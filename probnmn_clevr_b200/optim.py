"""Gradient clamp + Adam in one pass over flat parameter buffers.

The reference's step ends with ``parameter.grad.clamp_(min=-5, max=5)`` over every parameter of the three trained models
(trainers/joint_training_trainer.py:182-188: ~260 element-wise launches) followed by ``torch.optim.Adam.step()``
(trainers/_trainer.py:103-108,193).  ``FusedClampAdam`` is a ``torch.optim.Optimizer`` with Adam's hyper-parameters and
state layout (``state[p] = {"step", "exp_avg", "exp_avg_sq"}``, so ``ReduceLROnPlateau`` and
``CheckpointManager(optimizer=...)`` work unchanged and checkpoints interchange with ``torch.optim.Adam``) whose ``step``
runs ``pnmn_clamp_adam`` (``csrc/optim.cu``): one launch per contiguous parameter range.  The drop-in models keep their
parameters as views of one flat buffer and their CUDA backward writes one flat gradient buffer with the same layout, so a
whole model is ONE launch that reads parameter, gradient and both moments once and writes parameter and moments once
(28 bytes per element, the HBM floor of clamp + Adam).  Parameters that are not laid out that way fall back to one launch
per tensor -- still the CUDA kernel; there is no eager path.

Semantics: a parameter whose ``.grad`` is ``None`` is skipped, like ``torch.optim.Adam``.  Inside a flat range whose
gradient buffer is attached as a whole (the module executor's), tensors that received no gradient this step hold zeros
and are stepped with g = 0 -- what the reference does under its pinned torch 1.4, where ``optimizer.zero_grad()`` leaves
zero tensors rather than ``None``.
"""
import ctypes
from typing import Dict, Iterable, List, Optional

import torch

from . import _lib as L


class _Range:
    """A maximal run of parameters that are consecutive views of one storage (64-float aligned gaps allowed)."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        first, last = params[0], params[-1]
        self.offsets = [p.storage_offset() - first.storage_offset() for p in params]
        self.numel = last.storage_offset() + last.numel() - first.storage_offset()
        self.exp_avg: Optional[torch.Tensor] = None
        self.exp_avg_sq: Optional[torch.Tensor] = None
        self.step = 0
        self.synced = True   # the per-parameter "step" tensors of the optimizer state agree with self.step
        self.pflat: Optional[torch.Tensor] = None
        self.uniform = True
        self.calls = 0

    def flat(self, tensors: List[torch.Tensor], full: bool = True) -> Optional[torch.Tensor]:
        """The range as ONE tensor if ``tensors`` (parameters or their gradients) are laid out like the parameters.
        ``full = False`` checks every tensor for ``None`` but the layout of the first, middle and last only (the
        222-parameter range of the module network costs ~0.2 ms of host time per full check; it is re-checked in full
        every 64th step)."""
        t0 = tensors[0]
        if t0 is None:
            return None
        storage = t0.untyped_storage().data_ptr()
        base = t0.storage_offset()
        n = len(tensors)
        picks = range(n) if full or n <= 32 else (0, n // 2, n - 1)
        if not full and any(t is None for t in tensors):
            return None
        for i in picks:
            t, off = tensors[i], self.offsets[i]
            if (t is None or t.dtype != torch.float32 or not t.is_contiguous() or t.untyped_storage().data_ptr() != storage
                    or t.storage_offset() != base + off):
                return None
        return torch.empty(0, dtype=torch.float32, device=t0.device).set_(t0.untyped_storage(), base, (self.numel,))


class FusedClampAdam(torch.optim.Optimizer):
    r"""Adam (``amsgrad=False``) with an optional element-wise gradient clamp applied first.

    Parameters
    ----------
    params: iterable of parameters or parameter-group dicts (as ``torch.optim.Adam``)
    lr, betas, eps, weight_decay: as ``torch.optim.Adam``
    clamp: float or None -- gradients are clamped to ``[-clamp, clamp]`` before the update
        (``parameter.grad.clamp_(min=-5, max=5)``, joint_training_trainer.py:187-188)
    write_clamped_grad: also store the clamped values in ``.grad`` (off by default: the gradients are zeroed right after)
    modules: optional models whose ``zero_grad`` should be used by ``zero_grad`` (their fast paths clear a flat gradient
        buffer with one memset and keep the gradient views attached)
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clamp: Optional[float] = None,
                 write_clamped_grad: bool = False, modules: Optional[Iterable[torch.nn.Module]] = None):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.clamp = float(clamp) if clamp else 0.0
        self.write_clamped_grad = bool(write_clamped_grad)
        self._modules = list(modules) if modules is not None else []
        self._ranges: Optional[List[List[_Range]]] = None
        self._range_key = None
        self.launches_last_step = 0

    # ---- ranges --------------------------------------------------------------------------------------------------------
    def _layout_key(self):
        return tuple((p.data_ptr(), p.numel()) for g in self.param_groups for p in (g["params"][0], g["params"][-1]))

    def _build_ranges(self):
        groups = []
        for g in self.param_groups:
            ranges, run = [], []
            for p in g["params"]:
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedClampAdam handles fp32 CUDA parameters only (there is no CPU fallback)")
                if run:
                    q = run[-1]
                    same = p.untyped_storage().data_ptr() == q.untyped_storage().data_ptr() and p.is_contiguous()
                    gap = p.storage_offset() - (q.storage_offset() + q.numel())
                    if not (same and 0 <= gap < 64 and p.storage_offset() % 4 == 0):
                        ranges.append(_Range(run))
                        run = []
                run.append(p)
            if run:
                ranges.append(_Range(run))
            groups.append(ranges)
        # carry existing per-parameter state over into range-wide moment buffers
        for ranges in groups:
            for r in ranges:
                p0 = r.params[0]
                r.exp_avg = torch.zeros(r.numel, dtype=torch.float32, device=p0.device)
                r.exp_avg_sq = torch.zeros(r.numel, dtype=torch.float32, device=p0.device)
                steps = set()
                for p, off in zip(r.params, r.offsets):
                    st = self.state.get(p)
                    ea, es = r.exp_avg[off:off + p.numel()].view(p.shape), r.exp_avg_sq[off:off + p.numel()].view(p.shape)
                    if st:
                        ea.copy_(st["exp_avg"])
                        es.copy_(st["exp_avg_sq"])
                        steps.add(int(st["step"]))
                    else:
                        steps.add(0)
                    self.state[p] = {"step": torch.tensor(float(int(st["step"]) if st else 0)), "exp_avg": ea, "exp_avg_sq": es}
                r.step = max(steps)
                r.uniform = len(steps) == 1
                r.pflat = r.flat(r.params)
        self._ranges = groups
        self._range_key = self._layout_key()

    def _sync_steps(self, r: _Range) -> None:
        if not r.synced:
            for p in r.params:
                self.state[p]["step"].fill_(float(r.step))
            r.synced = True

    def state_dict(self):
        """``torch.optim.Adam``'s format.  While a range is stepped as a whole its step count is kept once; the
        per-parameter ``"step"`` entries are refreshed here (read ``optimizer.state`` through this method)."""
        for ranges in self._ranges or []:
            for r in ranges:
                self._sync_steps(r)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._ranges = None   # moments are re-packed into range-wide buffers on the next step

    # ---- step ----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, only=None):
        """``only``: an optional ``nn.Module`` -- step just the parameter ranges that belong to it (a caller that knows when
        each model's gradients are final can update that model right away, on the stream its backward pass ran on)."""
        mine = None if only is None else {id(p) for p in only.parameters()}
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._ranges is None or self._range_key != self._layout_key():
            self._build_ranges()
        lib = L.lib()
        launches = 0
        for g, ranges in zip(self.param_groups, self._ranges):
            lr, (b1, b2), eps, wd = g["lr"], g["betas"], g["eps"], g["weight_decay"]
            for r in ranges:
                p0 = r.params[0]
                if mine is not None and id(p0) not in mine:
                    continue
                grads = [p.grad for p in r.params]
                pflat = r.pflat if r.uniform else None
                r.calls += 1
                gflat = r.flat(grads, full=r.calls % 64 == 1) if pflat is not None else None
                with torch.cuda.device(p0.device):
                    stream = ctypes.c_void_p(torch.cuda.current_stream(p0.device).cuda_stream)
                    if gflat is not None and pflat.data_ptr() % 16 == 0 and gflat.data_ptr() % 16 == 0:
                        r.step += 1
                        L.check(lib.pnmn_clamp_adam(ctypes.c_void_p(pflat.data_ptr()), ctypes.c_void_p(gflat.data_ptr()),
                                                    ctypes.c_void_p(r.exp_avg.data_ptr()), ctypes.c_void_p(r.exp_avg_sq.data_ptr()),
                                                    r.numel, r.step, lr, b1, b2, eps, wd, self.clamp,
                                                    1 if self.write_clamped_grad else 0, stream), "pnmn_clamp_adam")
                        launches += 1
                        r.synced = False   # (the 222 per-parameter "step" tensors are brought up to date on demand)
                        continue
                    # per-tensor launches (gradients not in one flat buffer, some missing, or unequal step counts)
                    self._sync_steps(r)
                    r.uniform = False
                    for p, grad in zip(r.params, grads):
                        if grad is None:
                            continue
                        if grad.is_sparse:
                            raise RuntimeError("FusedClampAdam does not support sparse gradients")
                        st = self.state[p]
                        gc = grad if grad.is_contiguous() and grad.data_ptr() % 16 == 0 and p.data_ptr() % 16 == 0 else None
                        if gc is None or not p.is_contiguous():
                            raise RuntimeError("FusedClampAdam needs contiguous, 16-byte aligned parameters and gradients")
                        st["step"] += 1
                        L.check(lib.pnmn_clamp_adam(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(gc.data_ptr()),
                                                    ctypes.c_void_p(st["exp_avg"].data_ptr()),
                                                    ctypes.c_void_p(st["exp_avg_sq"].data_ptr()), p.numel(), int(st["step"]),
                                                    lr, b1, b2, eps, wd, self.clamp, 1 if self.write_clamped_grad else 0,
                                                    stream), "pnmn_clamp_adam")
                        launches += 1
                    steps = {int(self.state[p]["step"]) for p in r.params}
                    if len(steps) == 1:
                        r.uniform, r.step = True, steps.pop()
        self.launches_last_step = launches if only is None else self.launches_last_step + launches
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        if self._modules and set_to_none:
            for m in self._modules:
                m.zero_grad(set_to_none=True)
            return
        super().zero_grad(set_to_none=set_to_none)
