"""Data parallelism over questions: one process per GPU, every rank runs the joint step on its own shard
of the batch and the gradients are averaged with ONE collective per flat buffer.

The reference only knows single-process ``nn.DataParallel`` (trainers/_trainer.py:98-100), which mis-executes the NMN
(its module table is a plain dict that ``replicate`` does not copy).  The path itself has no exchange
step — samples are independent in forward — so the gradient average is the only collective (NCCL over NVLink on the
GPU box; the same code runs over ``gloo`` in the CPU tests).
"""
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of this rank's contiguous shard of a global batch (remainder rows go to the lowest ranks)."""
    base, rem = divmod(n_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gradient_buckets(models: Iterable[torch.nn.Module]) -> List[torch.Tensor]:
    """Tensors to all-reduce: a model whose backward left all its gradients inside one flat buffer contributes that
    buffer (the executor's / the seq2seq kernels' backward write one), everything else its ``.grad`` tensors."""
    buckets: List[torch.Tensor] = []
    for m in models:
        params = [p for p in m.parameters() if p.grad is not None]
        if not params:
            continue
        # gradients that are views into the same storage (a flat gradient buffer) form ONE bucket; buckets are emitted in
        # parameter order so that every rank issues the same sequence of collectives
        by_storage = {}
        for p in params:
            by_storage.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p.grad)
        for gs in by_storage.values():
            if len(gs) == 1:
                buckets.append(gs[0])
                continue
            flat = torch.empty(0, dtype=gs[0].dtype, device=gs[0].device).set_(gs[0].untyped_storage())
            lo = min(g.storage_offset() for g in gs)
            hi = max(g.storage_offset() + g.numel() for g in gs)
            buckets.append(flat[lo:hi])
    return buckets


class GradientOverlap:
    """Starts the all-reduce of a parameter's gradient the moment autograd has accumulated it, instead of after the whole
    backward pass.  In the NMN step the classifier's gradients come first (classifier.4.weight alone is 205 MB of the
    257 MB the step reduces) and the module executor's backward + weight-gradient kernels (~1.5 ms) follow: the large
    collective then travels over NVLink underneath them.  ``finish`` waits for the early collectives and reduces whatever
    is left (the flat buffers the CUDA backward writes, which autograd never sees).  Every rank runs the same autograd
    graph, so the hooks fire -- and the collectives are issued -- in the same order everywhere."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 min_numel: int = 1, weight: float = 1.0):
        """Gradients of at least ``min_numel`` elements start their own collective from the hook; smaller ones are only
        collected there and travel TOGETHER in one flat buffer at ``finish`` (a collective costs the host 0.1-0.3 ms to
        launch whatever its size, and the step is host-bound at eight processes per box)."""
        self.group = group
        self.weight = weight
        self.min_numel = min_numel
        self._pending: List[Tuple[torch.Tensor, "dist.Work"]] = []
        self._small: List[torch.Tensor] = []
        self._hooks = []
        self._avg: Optional[bool] = None
        for p in params:
            if p.requires_grad:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _op(self):
        """NCCL averages inside the collective (no per-tensor division kernel afterwards); gloo only sums."""
        if self._avg is None:
            self._avg = dist.get_backend(self.group) == "nccl"
        return dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        g = p.grad
        if g is None or not dist.is_initialized():
            return
        if g.numel() < self.min_numel:
            self._small.append(g)
            return
        if self.weight != 1.0:
            g.mul_(self.weight)
        self._pending.append((g, dist.all_reduce(g, op=self._op(), group=self.group, async_op=True)))

    def finish(self, models: Optional[Iterable[torch.nn.Module]] = None, buckets: Optional[List[torch.Tensor]] = None) -> int:
        """Average every gradient of ``models``: wait for the collectives the hooks started, all-reduce the rest.
        A caller that knows what is left (e.g. one flat gradient buffer) passes it as ``buckets`` and saves the walk over
        every parameter.  Returns the number of collectives this step used."""
        world = dist.get_world_size(self.group)
        if buckets is None:
            early = {g.data_ptr() for g, _ in self._pending} | {g.data_ptr() for g in self._small}
            buckets = [b for b in gradient_buckets(models) if b.data_ptr() not in early]
        else:
            buckets = list(buckets)
        op = self._op()
        # the small hooked gradients: one flat buffer, one collective, copied back afterwards
        small, packed = self._small, None
        if small:
            packed = torch.cat([g.reshape(-1) for g in small])
            buckets.append(packed)
        handles = []
        for b in buckets:
            if self.weight != 1.0:
                b.mul_(self.weight)
            handles.append(dist.all_reduce(b, op=op, group=self.group, async_op=True))
        for g, h in self._pending:
            h.wait()
            if not self._avg:
                g.div_(world)
        for b, h in zip(buckets, handles):
            h.wait()
            if not self._avg:
                b.div_(world)
        if packed is not None:
            off = 0
            for g in small:
                g.copy_(packed[off:off + g.numel()].view_as(g))
                off += g.numel()
        n = len(self._pending) + len(buckets)
        self._pending = []
        self._small = []
        return n

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(models: Iterable[torch.nn.Module], group: Optional[dist.ProcessGroup] = None,
                        weight: float = 1.0) -> int:
    """Average the gradients of ``models`` over the process group; returns the number of collectives issued.
    ``weight`` is this rank's share of the global objective (e.g. rows_on_rank / rows_total * world_size) so that ranks
    with unequal shard sizes reproduce the single-process global-batch mean exactly (SURVEY.md §8e caveat)."""
    world = dist.get_world_size(group)
    buckets = gradient_buckets(models)
    avg = dist.get_backend(group) == "nccl"   # NCCL averages inside the collective; gloo only sums
    handles = []
    for b in buckets:
        if weight != 1.0:
            b.mul_(weight)
        handles.append(dist.all_reduce(b, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    if not avg:
        for b in buckets:
            b.div_(world)
    return len(buckets)
