"""Host -> device input pipeline for the hot path.

The reference moves every batch to the GPU synchronously inside the training loop (trainers/_trainer.py:283-287: one
blocking ``.to(device)`` per key) from per-item h5py reads (data/readers.py:100-103).  At the rates of the CUDA path the
205 MB of image features per 256-question batch (4 ms over PCIe 5) would serialise with compute, so batches are staged
from PINNED host memory on a side stream while the previous step runs, and handed to the compute stream with an event.
"""
from typing import Dict, Hashable, List, Sequence, Tuple

import torch


class DevicePrefetcher:
    """``submit(key, tensors)`` starts the asynchronous copies; ``get(key)`` makes the current stream wait for them."""

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._slots: Dict[Hashable, Tuple[List[torch.Tensor], torch.cuda.Event]] = {}

    def submit(self, key: Hashable, tensors: Sequence[torch.Tensor]) -> None:
        for t in tensors:
            if t.device.type == "cpu" and not t.is_pinned():
                raise ValueError("DevicePrefetcher needs pinned host tensors (pageable copies are synchronous)")
        with torch.cuda.stream(self.stream):
            out = [t.to(self.device, non_blocking=True) for t in tensors]
            event = torch.cuda.Event()
            event.record(self.stream)
        self._slots[key] = (out, event)

    def get(self, key: Hashable) -> List[torch.Tensor]:
        out, event = self._slots.pop(key)
        current = torch.cuda.current_stream(self.device)
        current.wait_event(event)
        for t in out:
            t.record_stream(current)  # the caching allocator must not recycle the block while `current` uses it
        return out

    def pending(self) -> int:
        return len(self._slots)
