"""imagecodecs_b200 -- B200 (sm_100a) baseline JPEG encoder behind the JPEG write path of
jstrom2002/ImageCodecs (Image::writeJpg -> tje_encode_to_file, codecs.cpp:851-854).

The product is the C-ABI shared library ``libjpeg_gpu.so`` (include/jpeg_gpu.h).  This module
is only the ctypes plumbing the tests and bench.py use to call it; there is NO Python or CPU
implementation of the encoder here -- if the library or a CUDA device is missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JPEG_GPU_LIB") or os.path.join(_HERE, "libjpeg_gpu.so")   # override: kernel-variant experiments

QMODE_TJE, QMODE_IJG = 0, 1
SUB_444, SUB_420 = 0, 1
FLAG_SWAP_RB = 1
FLAG_RESTART = 2
OK, ERR_ARG, ERR_CAPACITY, ERR_CUDA = 0, 1, 2, 3

WRITE_FUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)   # jpeg_enc.h:152


class Image(C.Structure):          # jpeg_gpu_image
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int),
                ("stride", C.c_int), ("quality_mode", C.c_int), ("quality", C.c_int),
                ("subsampling", C.c_int), ("pixels_on_device", C.c_int), ("flags", C.c_int)]


class Output(C.Structure):         # jpeg_gpu_output
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("size", C.c_size_t), ("status", C.c_int)]


class Stream(C.Structure):         # jpeg_gpu_stream
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


class Decoded(C.Structure):        # jpeg_gpu_decoded
    _fields_ = [("pixels", C.c_void_p), ("capacity", C.c_size_t), ("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int), ("status", C.c_int)]


class BatchOpts(C.Structure):      # jpeg_gpu_batch_opts
    _fields_ = [("device", C.c_int), ("outputs_on_device", C.c_int), ("stream", C.c_void_p),
                ("debug_window_words", C.c_int)]


#: every symbol include/jpeg_gpu.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "jpeg_gpu_init": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "jpeg_gpu_shutdown": (None, []),
    "jpeg_gpu_device_count": (C.c_int, []),
    "jpeg_gpu_last_error": (C.c_char_p, []),
    "jpeg_gpu_max_encoded_size": (C.c_size_t, [C.c_int] * 4),
    "jpeg_gpu_emit_headers": (C.c_size_t, [C.c_int] * 6 + [C.c_void_p, C.c_size_t]),
    "jpeg_gpu_emit_headers_for": (C.c_size_t, [C.POINTER(Image), C.c_void_p, C.c_size_t]),
    "jpeg_gpu_encode_batch": (C.c_int, [C.POINTER(Image), C.c_int, C.POINTER(Output), C.POINTER(BatchOpts)]),
    "jpeg_gpu_plan_create": (C.c_void_p, [C.POINTER(Image), C.c_int, C.c_int, C.c_int]),
    "jpeg_gpu_plan_set_pixels": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "jpeg_gpu_plan_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "jpeg_gpu_plan_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jpeg_gpu_plan_launches": (C.c_int, [C.c_void_p]),
    "jpeg_gpu_plan_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "jpeg_gpu_plan_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "jpeg_gpu_plan_pass_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "jpeg_gpu_plan_is_fused": (C.c_int, [C.c_void_p]),
    "jpeg_gpu_plan_fetch": (C.c_int, [C.c_void_p, C.POINTER(Output), C.c_int, C.c_void_p]),
    "jpeg_gpu_plan_encoded_size": (C.c_size_t, [C.c_void_p, C.c_int]),
    "jpeg_gpu_plan_num_blocks": (C.c_size_t, [C.c_void_p]),
    "jpeg_gpu_plan_attach_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "jpeg_gpu_plan_destroy": (None, [C.c_void_p]),
    "jpeg_gpu_decode_info": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jpeg_gpu_decode": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jpeg_gpu_decode_batch": (C.c_int, [C.POINTER(Stream), C.c_int, C.POINTER(Decoded), C.c_int, C.POINTER(C.c_float)]),
    "jpeg_gpu_decode_timed": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_float)]),
    "jpeg_gpu_encode_to_file": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "jpeg_gpu_encode_to_file_at_quality": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "jpeg_gpu_encode_with_func": (C.c_int, [WRITE_FUNC, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class JpegGpuError(RuntimeError):
    pass


def lib():
    """Load libjpeg_gpu.so (raises if it has not been built -- there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JpegGpuError("%s is missing: run `python -m imagecodecs_b200.build` "
                               "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return (lib().jpeg_gpu_last_error() or b"").decode()


def init(device_ids=None):
    L = lib()
    if device_ids is None:
        n = L.jpeg_gpu_init(None, 0)
    else:
        arr = (C.c_int * len(device_ids))(*device_ids)
        n = L.jpeg_gpu_init(arr, len(device_ids))
    if n <= 0:
        raise JpegGpuError("jpeg_gpu_init failed: " + last_error())
    return n


def max_encoded_size(w, h, ncomp, sub=SUB_444):
    return int(lib().jpeg_gpu_max_encoded_size(w, h, ncomp, sub))


def emit_headers(w, h, ncomp, qmode=QMODE_TJE, quality=3, sub=SUB_444, flags=0):
    buf = np.empty(2048, np.uint8)
    if flags:
        im = Image(None, w, h, ncomp, 0, qmode, quality, sub, 0, flags)
        n = lib().jpeg_gpu_emit_headers_for(C.byref(im), buf.ctypes.data, buf.size)
    else:
        n = lib().jpeg_gpu_emit_headers(w, h, ncomp, qmode, quality, sub, buf.ctypes.data, buf.size)
    return buf[:n].tobytes()


def _describe(px, qmode, quality, sub, flags=0, bottom_up=False):
    """px: numpy uint8 [h,w,c] / [h,w] (host) or anything with data_ptr()/shape (torch CUDA tensor).
    bottom_up: the array holds the rows last-to-first (BMP order); described with a negative stride."""
    on_dev = 0
    if hasattr(px, "data_ptr"):
        shape = tuple(px.shape)
        ptr = px.data_ptr()
        on_dev = 1 if px.is_cuda else 0
        if not px.is_contiguous():
            raise ValueError("pixels must be contiguous")
    else:
        shape = px.shape
        ptr = px.ctypes.data
    if len(shape) == 2:
        shape = shape + (1,)
    h, w, c = shape
    if bottom_up:
        return Image(ptr + (h - 1) * w * c, w, h, c, -w * c, qmode, quality, sub, on_dev, flags)
    return Image(ptr, w, h, c, 0, qmode, quality, sub, on_dev, flags)


def encode_batch(images, qmode=QMODE_TJE, quality=3, sub=SUB_444, device=-1, capacity=None, win_words=0, flags=0,
                 bottom_up=False):
    """Encode a list of images (numpy host arrays or torch CUDA tensors, uint8 [h,w,c]).

    qmode / quality / sub / flags may be scalars or per-image sequences.  Returns (list[bytes|None], statuses).
    """
    L = lib()
    n = len(images)
    per = lambda v, i: v[i] if isinstance(v, (list, tuple, np.ndarray)) else v
    keep = []
    descs = (Image * n)()
    outs = (Output * n)()
    bufs = []
    for i, px in enumerate(images):
        if not hasattr(px, "data_ptr"):
            px = np.ascontiguousarray(px, dtype=np.uint8)
        keep.append(px)
        descs[i] = _describe(px, per(qmode, i), per(quality, i), per(sub, i), per(flags, i), bottom_up)
        cap = capacity if capacity is not None else 2048 + 6 * descs[i].width * descs[i].height + 4096
        b = np.empty(cap, np.uint8)
        bufs.append(b)
        outs[i] = Output(b.ctypes.data, cap, 0, 0)
    opts = BatchOpts(device, 0, None, win_words)
    ok = L.jpeg_gpu_encode_batch(descs, n, outs, C.byref(opts))
    res = [bufs[i][:outs[i].size].tobytes() if outs[i].status == OK else None for i in range(n)]
    st = [outs[i].status for i in range(n)]
    if ok != sum(1 for s in st if s == OK):
        raise JpegGpuError("inconsistent batch result: " + last_error())
    if any(s == ERR_CUDA for s in st):
        raise JpegGpuError("CUDA failure: " + last_error())
    return res, st


def decode(jpeg, timed=False):
    """Decode one JPEG file (bytes) on the GPU; returns uint8 [h,w,3] / [h,w] (and the kernels' device ms if timed)."""
    L = lib()
    buf = np.frombuffer(jpeg, dtype=np.uint8)
    w, h, nc = C.c_int(0), C.c_int(0), C.c_int(0)
    if not L.jpeg_gpu_decode_info(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(nc)):
        raise JpegGpuError("decode: " + last_error())
    out = np.empty(w.value * h.value * nc.value, np.uint8)
    ms = C.c_float(0)
    ok = L.jpeg_gpu_decode_timed(buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(nc), C.byref(ms))
    if not ok:
        raise JpegGpuError("decode: " + last_error())
    img = out.reshape(h.value, w.value, 3) if nc.value == 3 else out.reshape(h.value, w.value)
    return (img, ms.value) if timed else img


def decode_batch(files, timed=False, alloc=None):
    """Decode a list of JPEG files (bytes) in one call; returns list of uint8 arrays (None where a file failed).
    alloc(nbytes) -> 1-D uint8 array supplies the pixel buffers (e.g. views of pinned memory); default: np.empty."""
    L = lib()
    n = len(files)
    bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
    ins = (Stream * n)(*[Stream(b.ctypes.data, b.size) for b in bufs])
    outs = (Decoded * n)()
    pix = []
    w, h, nc = C.c_int(0), C.c_int(0), C.c_int(0)
    for i, b in enumerate(bufs):
        ok = L.jpeg_gpu_decode_info(b.ctypes.data, b.size, C.byref(w), C.byref(h), C.byref(nc))
        size = w.value * h.value * nc.value if ok else 1
        a = alloc(size) if alloc is not None else np.empty(size, np.uint8)
        pix.append(a)
        outs[i] = Decoded(a.ctypes.data, a.size if ok else 0, 0, 0, 0, 0)
    ms = C.c_float(0)
    L.jpeg_gpu_decode_batch(ins, n, outs, 0, C.byref(ms) if timed else None)    # timed: one piece, kernels back to back; else pipelined chunks
    res = []
    for i in range(n):
        o = outs[i]
        if o.status != OK:
            res.append(None)
        else:
            res.append(pix[i].reshape(o.height, o.width, 3) if o.ncomp == 3 else pix[i].reshape(o.height, o.width))
    return (res, ms.value) if timed else res


class Plan:
    """A prepared batch on one device (jpeg_gpu_plan_*)."""

    def __init__(self, descs, device=0, win_words=0):
        self._L = lib()
        init()
        self.n = len(descs)
        self._descs = (Image * self.n)(*descs)
        self._h = self._L.jpeg_gpu_plan_create(self._descs, self.n, device, win_words)
        if not self._h:
            raise JpegGpuError("plan_create failed: " + last_error())

    @classmethod
    def for_arrays(cls, images, qmode=QMODE_TJE, quality=3, sub=SUB_444, device=0, win_words=0, flags=0):
        per = lambda v, i: v[i] if isinstance(v, (list, tuple, np.ndarray)) else v
        descs = [_describe(px, per(qmode, i), per(quality, i), per(sub, i), per(flags, i)) for i, px in enumerate(images)]
        p = cls(descs, device, win_words)
        p._keep = list(images)
        return p

    @property
    def launches(self):
        return self._L.jpeg_gpu_plan_launches(self._h)

    @property
    def num_blocks(self):
        return self._L.jpeg_gpu_plan_num_blocks(self._h)

    def set_pixels(self, i, dev_ptr):
        if not self._L.jpeg_gpu_plan_set_pixels(self._h, i, dev_ptr):
            raise JpegGpuError("set_pixels failed")

    def upload(self, i, host_ptr, stream=None):
        if not self._L.jpeg_gpu_plan_upload(self._h, i, host_ptr, stream):
            raise JpegGpuError("upload failed: " + last_error())

    def enable_timing(self, on=True):
        if not self._L.jpeg_gpu_plan_enable_timing(self._h, 1 if on else 0):
            raise JpegGpuError("enable_timing failed: " + last_error())

    def kernel_times(self):
        """(encode_ms, stuff_ms) of the last completed run (needs enable_timing)."""
        a, b = C.c_float(0), C.c_float(0)
        if not self._L.jpeg_gpu_plan_kernel_times(self._h, C.byref(a), C.byref(b)):
            raise JpegGpuError("kernel_times unavailable")
        return a.value, b.value

    def pass_times(self):
        """(transform_ms, entropy_ms, stuff_ms) of the last run (timing enabled, run complete)."""
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        if not self._L.jpeg_gpu_plan_pass_times(self._h, C.byref(a), C.byref(b), C.byref(c)):
            raise JpegGpuError("pass_times unavailable")
        return a.value, b.value, c.value

    @property
    def fused(self):
        return bool(self._L.jpeg_gpu_plan_is_fused(self._h))

    def attach_debug(self, coefs_ptr, bits_ptr):
        self._L.jpeg_gpu_plan_attach_debug(self._h, coefs_ptr, bits_ptr)

    def run(self, stream=None):
        if not self._L.jpeg_gpu_plan_run(self._h, stream):
            raise JpegGpuError("plan_run failed: " + last_error())

    def encoded_size(self, i):
        return int(self._L.jpeg_gpu_plan_encoded_size(self._h, i))

    def fetch_into(self, outs, on_device=0, stream=None):
        return self._L.jpeg_gpu_plan_fetch(self._h, outs, on_device, stream)

    def fetch(self, stream=None):
        """Returns list[bytes|None] (host copies of the complete JPEG files)."""
        outs = (Output * self.n)()
        bufs = []
        for i in range(self.n):
            sz = self.encoded_size(i)
            b = np.empty(max(sz, 1), np.uint8)
            bufs.append(b)
            outs[i] = Output(b.ctypes.data, b.size, 0, 0)
        self.fetch_into(outs, 0, stream)
        return [bufs[i][:outs[i].size].tobytes() if outs[i].status == OK else None for i in range(self.n)]

    def close(self):
        if self._h:
            self._L.jpeg_gpu_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
