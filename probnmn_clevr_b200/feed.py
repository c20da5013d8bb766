"""Host -> device input pipeline for the hot path.

The reference moves every batch to the GPU synchronously inside the training loop (trainers/_trainer.py:283-287: one
blocking ``.to(device)`` per key) from per-item h5py reads (data/readers.py:100-103).  At the rates of the CUDA path the
205 MB of image features per 256-question batch (3.7 ms over PCIe 5) would serialise with compute, so batches are staged
from PINNED host memory on a side stream while the previous step runs, and handed to the compute stream with an event.

The device side is a ring of ``depth`` preallocated buffer sets (no allocator traffic in steady state: a fresh 205 MB
block per step made the caching allocator wait for, or cudaMalloc around, blocks still in use by the compute stream).
A slot is overwritten only after the compute stream has passed the point where its previous tenant was handed back,
which with ``depth`` = 3 lies more than a full step in the past.
"""
import os
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import torch


class DevicePrefetcher:
    """``submit(key, tensors)`` starts the asynchronous copies; ``get(key)`` makes the current stream wait for them.
    The tensors returned by ``get(i)`` belong to the caller until the NEXT ``get``: that call marks the slot as released
    at the current point of the compute stream, so work enqueued on them BEFORE the next ``get`` is safe, work enqueued
    after it races with a later ``submit`` that recycles the slot (``depth`` - 1 submits later)."""

    def __init__(self, device: torch.device, depth: int = 3):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.chunks = int(os.environ.get("PNMN_FEED_CHUNKS", "1"))  # large tensors are copied in this many slices
        self._bufs: List[Optional[List[torch.Tensor]]] = [None] * depth
        self._free: List[Optional[torch.cuda.Event]] = [None] * depth  # compute stream is done with the slot's tenant
        self._next = 0
        self._slots: Dict[Hashable, Tuple[int, torch.cuda.Event]] = {}
        self._out: Optional[int] = None  # slot handed out by the last get()

    def submit(self, key: Hashable, tensors: Sequence[torch.Tensor]) -> None:
        for t in tensors:
            if t.device.type == "cpu" and not t.is_pinned():
                raise ValueError("DevicePrefetcher needs pinned host tensors (pageable copies are synchronous)")
        slot = self._next
        self._next = (self._next + 1) % self.depth
        if any(s == slot for s, _ in self._slots.values()) or slot == self._out:
            raise RuntimeError(f"DevicePrefetcher: more than {self.depth - 1} batches in flight")
        bufs = self._bufs[slot]
        if bufs is None or len(bufs) != len(tensors) or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, tensors)):
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in tensors]
            self._bufs[slot] = bufs
            self._free[slot] = None
            # The new blocks come from the COMPUTE stream's allocator pool: the caching allocator hands out memory whose
            # previous owner may still have work pending on that stream (e.g. a gradient freed by zero_grad whose in-place
            # average has not run yet).  The copy stream must not write them before that work has drained.  Only on
            # (re)allocation, i.e. never in steady state.
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])
            for b, t in zip(bufs, tensors):
                n = t.shape[0] if t.dim() > 0 else 1
                if self.chunks > 1 and t.numel() * t.element_size() >= (32 << 20) and n >= self.chunks:
                    # one DMA per slice: anything else that needs the copy engine waits for a slice, not for the whole batch
                    step = (n + self.chunks - 1) // self.chunks
                    for lo in range(0, n, step):
                        b[lo:lo + step].copy_(t[lo:lo + step], non_blocking=True)
                else:
                    b.copy_(t, non_blocking=True)
            event = torch.cuda.Event()
            event.record(self.stream)
        self._slots[key] = (slot, event)

    def get(self, key: Hashable) -> List[torch.Tensor]:
        current = torch.cuda.current_stream(self.device)
        if self._out is not None:
            # everything issued so far on the compute stream (the previous batch's whole step) precedes this point
            ev = torch.cuda.Event()
            ev.record(current)
            self._free[self._out] = ev
        slot, event = self._slots.pop(key)
        current.wait_event(event)
        self._out = slot
        return list(self._bufs[slot])

    def pending(self) -> int:
        return len(self._slots)
