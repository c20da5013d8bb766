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
import ctypes
import os
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import torch

from . import _lib as L


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


class ImageFeatureCache:
    r"""
    Device-resident image features keyed by ``image_index`` -- the B200 counterpart of the reference's
    ``ClevrImageFeaturesReader(in_memory=True)`` (data/readers.py:63-108), which keeps the whole H5 dataset in host memory
    and is indexed per item by ``JointTrainingDataset.__getitem__`` (data/datasets.py:209-228).

    The features are stored ONCE in HBM as fp16 holding exactly the operand value the executor derives from an fp32 feature
    (``pnmn_round_features_f16``: tf32 rounding, then a saturating fp16 copy), 392 KB per (1024, 14, 14) image: the 70 000
    images of the CLEVR train split take 27.4 GB of the 180 GB.  A training step then sends only its image indices over
    PCIe (2 KB instead of 205 MB for 256 questions; many questions share an image) and ``gather`` builds the batch on the
    device.  ``NeuralModuleNetwork.forward`` takes the fp16 batch as it is and returns results identical to those for the
    fp32 features.

    Parameters
    ----------
    features: array-like (num_images, C, H, W) -- a NumPy array, a torch tensor or an ``h5py`` dataset (float32 / float64, as
        written by scripts/preprocess/extract_features.py:119-121); read in chunks of ``chunk_images``
    device: the GPU that holds the cache
    split: optional split name (the reference reader's ``.split`` property)
    """

    def __init__(self, features, device, split: Optional[str] = None, chunk_images: int = 256):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ImageFeatureCache keeps the features in GPU memory; there is no CPU fallback")
        self._split = split
        n = len(features)
        shape = tuple(features.shape[1:])
        if (shape[0] * shape[1] * shape[2]) % 4:
            raise ValueError("feature size per image must be a multiple of 4")
        self.features = torch.empty((n,) + shape, dtype=torch.float16, device=self.device)
        lib = L.lib()
        staging = [torch.empty((chunk_images,) + shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        events: List[Optional[torch.cuda.Event]] = [None, None]
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            for k, start in enumerate(range(0, n, chunk_images)):
                stop = min(start + chunk_images, n)
                host = staging[k % 2]
                if events[k % 2] is not None:
                    events[k % 2].synchronize()          # the previous copy out of this staging buffer has been issued and run
                host[: stop - start].copy_(torch.as_tensor(features[start:stop]))
                dev = host[: stop - start].to(self.device, non_blocking=True)
                L.check(lib.pnmn_round_features_f16(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(self.features[start:stop].data_ptr()),
                                                    dev.numel(), ctypes.c_void_p(stream.cuda_stream)), "pnmn_round_features_f16")
                ev = torch.cuda.Event()
                ev.record(stream)
                events[k % 2] = ev
            stream.synchronize()

    def __len__(self) -> int:
        return self.features.shape[0]

    def __getitem__(self, index):
        """One image's features as the reference reader returns them (host, float32) -- for code that still goes through a
        per-item Dataset; the hot path uses ``gather``."""
        return self.features[index].float().cpu().numpy()

    def gather(self, image_indices: torch.Tensor) -> torch.Tensor:
        """(B,) image indices (host or device) -> (B, C, H, W) fp16 batch on the device."""
        idx = image_indices.to(self.device, torch.int64, non_blocking=True)
        return self.features.index_select(0, idx)

    @property
    def split(self):
        return self._split
