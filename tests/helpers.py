"""Shared helpers of the test-suite."""
import hashlib

import numpy as np


def kat_image(entry, fixture_pixels):
    """Materialise the input image of a known-answer entry."""
    import oracle
    gen = entry["gen"]
    if "fixture" in gen:
        return fixture_pixels[gen["fixture"]]
    w, h, c, n, kind = gen["synth"]
    return oracle.synth_image(w, h, c, n, kind)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def psnr(a, b):
    m = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if m == 0 else 10 * np.log10(255.0 ** 2 / m)
