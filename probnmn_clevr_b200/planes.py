"""Torch-side helpers for the executor's activation format ("planes", see csrc/executor.h).

These are used by tests and debugging only; the product path converts layouts inside the CUDA
library (csrc/layout.cu).
"""
import torch

FMT = {16: (16, 256), 18: (18, 324), 22: (22, 484)}  # row stride S -> (S, slots per plane P)
GUARD_FLOATS = 16384 // 4


def fmt_for_dilation(d: int):
    return FMT[16] if d <= 2 else (FMT[18] if d == 4 else FMT[22])


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round to nearest (ties away), keep 10 mantissa bits."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def to_planes(x: torch.Tensor, S: int) -> torch.Tensor:
    """(C,14,14) -> (C/4, P, 4) with zero padding; P = S*S."""
    C = x.shape[0]
    buf = torch.zeros(C // 4, S, S, 4, dtype=x.dtype, device=x.device)
    buf[:, :14, :14, :] = x.view(C // 4, 4, 14, 14).permute(0, 2, 3, 1)
    return buf.view(C // 4, S * S, 4)


def from_planes(p: torch.Tensor, S: int) -> torch.Tensor:
    """(C/4, P, 4) -> (C,14,14)."""
    kc = p.shape[0]
    return p.view(kc, S, S, 4)[:, :14, :14, :].permute(0, 3, 1, 2).reshape(kc * 4, 14, 14)


def map_to_slots(m: torch.Tensor) -> torch.Tensor:
    """(14,14) attention map -> 256-slot P16 map."""
    buf = torch.zeros(16, 16, dtype=m.dtype, device=m.device)
    buf[:14, :14] = m
    return buf.view(256)


def map_from_slots(s: torch.Tensor) -> torch.Tensor:
    return s.view(16, 16)[:14, :14]


def to_half_planes(x: torch.Tensor, S: int) -> torch.Tensor:
    """(C,14,14) -> fp16 (C/8, P, 8) shadow layout."""
    C = x.shape[0]
    buf = torch.zeros(C // 8, S, S, 8, dtype=torch.float16, device=x.device)
    buf[:, :14, :14, :] = x.view(C // 8, 8, 14, 14).permute(0, 2, 3, 1).to(torch.float16)
    return buf.view(C // 8, S * S, 8)
