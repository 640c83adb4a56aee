"""Synthetic test images on any torch device, bit-identical to oracle/synth.py.

torch has no uint64 arithmetic, so the u64 wrap-around math runs in int64 (two's-complement
multiply/add wrap identically) with logical right shifts spelled out.
"""
import torch

_GOLD = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB


def _s64(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z, n):
    return (z >> n) & ((1 << (64 - n)) - 1)


def _mix(z):
    z = (z ^ _lsr(z, 30)) * _s64(_M1)
    z = (z ^ _lsr(z, 27)) * _s64(_M2)
    return z ^ _lsr(z, 31)


def synth_batch(count, w, h, ncomp=3, kind="photo", seed=1, first=0, device="cpu", chunk=16):
    """uint8 [count, h, w, ncomp]; image i is oracle.synth_image(w, h, ncomp, n=first+i, kind, seed)."""
    out = torch.empty((count, h, w, ncomp), dtype=torch.uint8, device=device)
    y = torch.arange(h, dtype=torch.int64, device=device).view(1, h, 1)
    x = torch.arange(w, dtype=torch.int64, device=device).view(1, 1, w)
    for lo in range(0, count, chunk):
        hi = min(count, lo + chunk)
        n = torch.arange(first + lo, first + hi, dtype=torch.int64, device=device).view(-1, 1, 1)
        idx = (n * h + y) * w + x
        for c in range(ncomp):
            if c == 3:
                out[lo:hi, ..., c] = 255
                continue
            z = _mix(idx * 4 + c + _s64(seed * _GOLD))
            if kind == "photo":
                v = 3 * x + 2 * y + 40 * c + 7 * n + (z & 15)
            elif kind == "noise":
                v = z
            else:
                raise ValueError(kind)
            out[lo:hi, ..., c] = (v & 255).to(torch.uint8)
    return out
