"""ctypes front-end for the CPU emulation of the kernel source (TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "libjpeg_emu.so")
_SRC = [os.path.join(_HERE, "emu_driver.cpp"), os.path.join(_HERE, "cuda_emu.h"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_stuff.cuh"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_kernel.cuh"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_transform.cuh"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_entropy.cuh"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_launch.h"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_device.h"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_tables.h"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_decode.cuh"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_decode.h"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_decode_host.cpp"),
        os.path.join(_ROOT, "imagecodecs_b200", "csrc", "jpeg_host.cpp")]
_lib = None


def build():
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRC):
        return
    csrc = os.path.join(_ROOT, "imagecodecs_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                    "-I" + _HERE, "-I" + csrc, "-o", _SO, _SRC[0], _SRC[-1], _SRC[-2]], check=True)


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.emu_encode.restype = C.c_int
        L.emu_encode.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                                                C.c_void_p, C.c_void_p]
        L.emu_encode_split.restype = C.c_int
        L.emu_encode_split.argtypes = L.emu_encode.argtypes
        L.emu_ticket_map.restype = C.c_int
        L.emu_ticket_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def emu_set_round_order(order):
    """Order of the emulated threads within a subsequence round: 0 alternating, 1 descending (predecessor's old state), 2 ascending,
    3 = eight racing host threads."""
    _load().emu_set_round_order(int(order))


def emu_decode(jpeg, sub_log2=-1, want_rounds=False):
    """Decode with the product's decoder source run on the CPU (tests only). Returns uint8 [h,w,3] / [h,w] or the nj error code.
    sub_log2: -1 = the library's own choice between restart intervals and subsequences, 0 = intervals only, n = subsequences
    of 2**n bytes wherever the stream allows them; want_rounds: also return how many rounds the subsequence decode took."""
    L = _load()
    L.emu_decode_sub.restype = C.c_int
    L.emu_decode_sub.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    buf = np.frombuffer(jpeg, dtype=np.uint8)
    w, h, nc, rounds = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    out = np.empty(1, np.uint8)
    rc = L.emu_decode_sub(buf.ctypes.data, buf.size, out.ctypes.data, 0, C.byref(w), C.byref(h), C.byref(nc), sub_log2, None)
    if rc != -1:
        return (rc, 0) if want_rounds else rc
    out = np.empty(w.value * h.value * nc.value, np.uint8)
    rc = L.emu_decode_sub(buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(nc), sub_log2, C.byref(rounds))
    if rc != 0:
        return (rc, rounds.value) if want_rounds else rc
    px = out.reshape(h.value, w.value, 3) if nc.value == 3 else out.reshape(h.value, w.value)
    return (px, rounds.value) if want_rounds else px


def emu_ticket_map(tiles, force_schedule=False):
    """The kernel's ticket -> (launch-wide tile, image) mapping for images with these tile counts."""
    tiles = np.ascontiguousarray(tiles, dtype=np.int32)
    total = int(tiles.sum())
    g = np.zeros(total, np.uint32); img = np.zeros(total, np.uint32)
    n = _load().emu_ticket_map(tiles.ctypes.data, len(tiles), 1 if force_schedule else 0, g.ctypes.data, img.ctypes.data)
    assert n == total
    return g, img


def emu_encode(batch, qmode=0, quality=3, sub=0, win_words=0, n_ctas=2, stages=False, cap=None, flags=0, bottom_up=False, split=True):
    """batch: uint8 [n,h,w,c].  Returns list of scan bytes (entropy-coded segment + EOI) [, coefs, bits].
    split: the split pipeline (transform -> entropy -> stuff, what the library runs) or the fused r01 kernel."""
    batch = np.ascontiguousarray(batch, dtype=np.uint8)
    if batch.ndim == 3:
        batch = batch[None]
    n, h, w, c = batch.shape
    mcu = 16 if (sub and c != 1) else 8
    bpm = 1 if c == 1 else (6 if sub else 3)
    nblk = ((w + mcu - 1) // mcu) * ((h + mcu - 1) // mcu) * bpm
    if cap is None:
        cap = nblk * 416 + 64
    out = np.zeros((n, cap), dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.uint32)
    coefs = np.zeros((n * nblk, 64), dtype=np.int16) if stages else None
    bits = np.zeros(n * nblk, dtype=np.uint32) if stages else None
    fn = _load().emu_encode_split if split else _load().emu_encode
    rc = fn(batch.ctypes.data, n, w, h, c, -w * c if bottom_up else 0, flags, sub, qmode, quality, win_words, n_ctas,
                            out.ctypes.data, cap, sizes.ctypes.data, status.ctypes.data,
                            coefs.ctypes.data if stages else None, bits.ctypes.data if stages else None)
    if rc != 0:
        raise RuntimeError("emulated kernel reported error %d" % rc)
    scans = [out[i, :min(int(sizes[i]), cap)].tobytes() for i in range(n)]
    if stages:
        return scans, sizes, status, coefs.reshape(n, nblk, 64), bits.reshape(n, nblk)
    return scans, sizes, status
