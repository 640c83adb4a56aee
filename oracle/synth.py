"""Deterministic integer-only synthetic images.

All arithmetic is u64 with wrap-around so host, device and Python agree bit for bit:
    idx = (n*H + y)*W + x
    z   = splitmix64_finaliser(idx*4 + c + seed*0x9E3779B97F4A7C15)
    photo: px = (3x + 2y + 40c + 7n + (z & 15)) & 255
    noise: px = z & 255
Alpha (c == 3) is 255.  Grayscale uses c = 0 only.
"""
import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def synth_image(w, h, ncomp=3, n=0, kind="photo", seed=1):
    """Return uint8 [h, w, ncomp] (ncomp 1, 3 or 4)."""
    with np.errstate(over="ignore"):
        y, x = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), indexing="ij")
        idx = (np.uint64(n) * np.uint64(h) + y) * np.uint64(w) + x
        out = np.empty((h, w, ncomp), dtype=np.uint8)
        for c in range(ncomp):
            if c == 3:
                out[..., c] = 255
                continue
            z = _mix(idx * np.uint64(4) + np.uint64(c) + np.uint64(seed) * _GOLD)
            if kind == "photo":
                v = np.uint64(3) * x + np.uint64(2) * y + np.uint64(40 * c) + np.uint64(7 * n) + (z & np.uint64(15))
            elif kind == "noise":
                v = z
            else:
                raise ValueError(kind)
            out[..., c] = (v & np.uint64(255)).astype(np.uint8)
    return out


def synth_batch(count, w, h, ncomp=3, kind="photo", seed=1, first=0):
    return np.stack([synth_image(w, h, ncomp, n=first + i, kind=kind, seed=seed) for i in range(count)])
