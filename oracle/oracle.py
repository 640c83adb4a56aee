"""ctypes front-end for the CPU checkers (oracle/jpeg_oracle.c, oracle/_ref)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORC_SO = os.path.join(_HERE, "libjpeg_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libtje_ref.so")

QMODE_TJE, QMODE_IJG = 0, 1
SUB_444, SUB_420 = 0, 1

_orc = None
_ref = None


def build(force=False):
    """Compile the restatement and, where /root/reference exists, the reference itself."""
    src = os.path.join(_HERE, "jpeg_oracle.c")
    stale = (not os.path.exists(_ORC_SO)) or os.path.getmtime(_ORC_SO) < os.path.getmtime(src)
    need_ref = os.path.exists("/root/reference/jpeg_enc.h") and not os.path.exists(_REF_SO)
    if force or stale or need_ref:
        subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _load_orc():
    global _orc
    if _orc is None:
        build()
        L = C.CDLL(_ORC_SO)
        L.orc_encode.restype = C.c_int
        L.orc_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ssize_t,
                                 C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_num_blocks.restype = C.c_size_t
        L.orc_encode_ex.restype = C.c_int
        L.orc_encode_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_headers_ex.restype = C.c_size_t
        L.orc_headers_ex.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_size_t]
        L.orc_num_blocks.argtypes = [C.c_int] * 4
        L.orc_headers.restype = C.c_size_t
        L.orc_headers.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_size_t]
        L.orc_build_qt.restype = C.c_int
        L.orc_build_qt.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_build_pqt.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_build_huff.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_read_bmp_mem.restype = C.c_int
        L.orc_read_bmp_mem.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _orc = L
    return _orc


def have_ref():
    return os.path.exists(_REF_SO) or os.path.exists("/root/reference/jpeg_enc.h")


def _load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            build()
        L = C.CDLL(_REF_SO)
        L.ref_tje_encode_mem.restype = C.c_int
        L.ref_tje_encode_mem.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ref_tje_encode_file.restype = C.c_int
        L.ref_tje_encode_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_nj_decode.restype = C.c_int
        L.ref_nj_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _ref = L
    return _ref


def _geom(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, nc = img.shape
    return img, w, h, nc


def num_blocks(w, h, ncomp, sub=SUB_444):
    return int(_load_orc().orc_num_blocks(w, h, ncomp, sub))


def worst_case_bytes(w, h, ncomp, sub=SUB_444):
    return 1024 + num_blocks(w, h, ncomp, sub) * 416 + 16


def restart_interval(ncomp, sub):
    """MCUs per restart interval of the GPU encoder's opt-in restart mode: one tile (24 blocks)."""
    return 24 if ncomp == 1 else (4 if sub == SUB_420 else 8)


def oracle_encode(img, qmode=QMODE_TJE, quality=3, sub=SUB_444, restart=0):
    """Encode [h,w,c] uint8 with the restatement; returns bytes or None if rejected.
    restart > 0: extended restart-interval mode (DRI + RSTn, see jpeg_oracle.c)."""
    img, w, h, nc = _geom(img)
    L = _load_orc()
    cap = 2048 + img.size * 2
    while True:
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = L.orc_encode_ex(img.ctypes.data, w, h, nc, 0, qmode, quality, sub, restart,
                             out.ctypes.data, cap, C.byref(n), None, None, None, 0, None)
        if rc != 1:
            return None
        if n.value <= cap:
            return out[:n.value].tobytes()
        cap = n.value


def oracle_stages(img, qmode=QMODE_TJE, quality=3, sub=SUB_444):
    """Return dict(jpeg, coefs[nb,64] int16 zigzag, block_bits[nb] u32, raw (unstuffed bytes), raw_bits)."""
    img, w, h, nc = _geom(img)
    L = _load_orc()
    nb = num_blocks(w, h, nc, sub)
    cap = worst_case_bytes(w, h, nc, sub)
    out = np.empty(cap, dtype=np.uint8)
    raw = np.zeros(cap, dtype=np.uint8)
    coefs = np.empty((nb, 64), dtype=np.int16)
    bits = np.empty(nb, dtype=np.uint32)
    n = C.c_size_t(0)
    rb = C.c_uint64(0)
    rc = L.orc_encode(img.ctypes.data, w, h, nc, 0, qmode, quality, sub,
                      out.ctypes.data, cap, C.byref(n), coefs.ctypes.data, bits.ctypes.data,
                      raw.ctypes.data, cap, C.byref(rb))
    if rc != 1:
        return None
    return dict(jpeg=out[:n.value].tobytes(), coefs=coefs, block_bits=bits,
                raw=raw[:(rb.value + 7) // 8].copy(), raw_bits=rb.value)


def oracle_headers(w, h, ncomp_out=3, sub=SUB_444, qmode=QMODE_TJE, quality=3, restart=0):
    out = np.empty(1024, dtype=np.uint8)
    n = _load_orc().orc_headers_ex(w, h, ncomp_out, sub, qmode, quality, restart, out.ctypes.data, out.size)
    return out[:n].tobytes()


def oracle_tables(qmode, quality):
    """Return (qt_luma u8[64], qt_chroma u8[64], pqt_luma f32[64], pqt_chroma f32[64], hlen u8[4,256], hcode u16[4,256])."""
    L = _load_orc()
    ql = np.empty(64, np.uint8); qc = np.empty(64, np.uint8)
    if L.orc_build_qt(qmode, quality, ql.ctypes.data, qc.ctypes.data) != 1:
        return None
    pl = np.empty(64, np.float32); pc = np.empty(64, np.float32)
    L.orc_build_pqt(ql.ctypes.data, pl.ctypes.data)
    L.orc_build_pqt(qc.ctypes.data, pc.ctypes.data)
    hl = np.empty((4, 256), np.uint8); hc = np.empty((4, 256), np.uint16)
    L.orc_build_huff(hl.ctypes.data, hc.ctypes.data)
    return ql, qc, pl, pc, hl, hc


def ref_encode(img, quality=3):
    """Encode with the compiled, unmodified reference (tje_encode_with_func). Returns (rc, bytes)."""
    img, w, h, nc = _geom(img)
    L = _load_ref()
    cap = 2048 + img.size * 2
    while True:
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = L.ref_tje_encode_mem(quality, w, h, nc, img.ctypes.data, out.ctypes.data, cap, C.byref(n))
        if n.value <= cap:
            return rc, out[:n.value].tobytes()
        cap = n.value


def ref_encode_file(path, img, quality=3):
    img, w, h, nc = _geom(img)
    return _load_ref().ref_tje_encode_file(os.fsencode(path), quality, w, h, nc, img.ctypes.data)


def ref_decode(jpeg):
    """Decode with the reference's NanoJPEG. Returns uint8 [h,w,3] or [h,w] or None."""
    L = _load_ref()
    buf = np.frombuffer(jpeg, dtype=np.uint8)
    cap = 1 << 20
    while True:
        out = np.empty(cap, dtype=np.uint8)
        w = C.c_int(0); h = C.c_int(0); col = C.c_int(0)
        rc = L.ref_nj_decode(buf.ctypes.data, buf.size, out.ctypes.data, cap, C.byref(w), C.byref(h), C.byref(col))
        if rc == -1:
            cap = w.value * h.value * 3 + 16
            continue
        if rc != 0:
            return None
        if col.value:
            return out[:w.value * h.value * 3].reshape(h.value, w.value, 3).copy()
        return out[:w.value * h.value].reshape(h.value, w.value).copy()


def read_bmp(data):
    """codecs.cpp:255-320 semantics: returns uint8 [|h|, w, 3] with bytes left in B,G,R order."""
    L = _load_orc()
    buf = np.frombuffer(data, dtype=np.uint8)
    w = C.c_int(0); h = C.c_int(0)
    px = np.empty(buf.size, dtype=np.uint8)
    rc = L.orc_read_bmp_mem(buf.ctypes.data, buf.size, px.ctypes.data, px.size, C.byref(w), C.byref(h))
    if rc != 1:
        return None
    ah = abs(h.value)
    return px[:ah * w.value * 3].reshape(ah, w.value, 3).copy()
