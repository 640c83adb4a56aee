"""CPU: the host half of the product library -- ABI surface, table construction, marker
emission.  No compute calls: this container has no GPU, and the library must say so loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import imagecodecs_b200 as jg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "jpeg_gpu.h")).read()
    declared = set(re.findall(r"JPEG_GPU_API\s+[\w\s\*]+?\b(jpeg_gpu_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(jg.SYMBOLS), declared ^ set(jg.SYMBOLS)
    L = C.CDLL(jg.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_library_is_self_contained():
    """static cudart, no torch / python / oracle dependency in the product library."""
    import subprocess
    out = subprocess.run(["ldd", jg.LIB_PATH], capture_output=True, text=True).stdout
    for bad in ("torch", "python", "oracle", "libcudart", "libtje"):
        assert bad not in out, out


@pytest.mark.parametrize("qm,q", [(0, 1), (0, 2), (0, 3), (1, 1), (1, 50), (1, 75), (1, 90), (1, 100)])
@pytest.mark.parametrize("nc,sub", [(3, 0), (4, 0), (3, 1), (1, 0)])
def test_headers_match_oracle(qm, q, nc, sub):
    for (w, h) in [(8, 8), (395, 348), (65535, 1)]:
        want = oracle.oracle_headers(w, h, 1 if nc == 1 else 3, sub, qm, q)
        got = jg.emit_headers(w, h, nc, qm, q, sub)
        assert got == want
        if nc != 1:
            assert len(got) == 655       # scan starts at offset 655


def test_headers_match_reference_golden_files(kat, golden_dir):
    for e in kat:
        if "file" not in e:
            continue
        ref = open(os.path.join(golden_dir, e["file"]), "rb").read()
        assert jg.emit_headers(e["w"], e["h"], e["ncomp"], jg.QMODE_TJE, e["tje_quality"], jg.SUB_444) == ref[:655]


def test_header_layout_facts():
    h = jg.emit_headers(395, 348, 3, jg.QMODE_TJE, 3, jg.SUB_444)
    assert h[:4] == b"\xff\xd8\xff\xe0" and h[6:11] == b"JFIF\0" and h[11:13] == b"\x01\x02"
    assert h[20:24] == b"\xff\xfe\x00\x1e" and h[24:52] == b"Created by Tiny JPEG Encoder"
    sof = h.index(b"\xff\xc0")
    assert h[sof + 5:sof + 9] == bytes([348 >> 8, 348 & 255, 395 >> 8, 395 & 255])   # height, then width
    assert h[-14:] == bytes([0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0])


def test_rejected_arguments():
    assert jg.emit_headers(8, 8, 3, jg.QMODE_TJE, 4, jg.SUB_444) == b""
    assert jg.emit_headers(8, 8, 3, jg.QMODE_IJG, 0, jg.SUB_444) == b""
    assert jg.emit_headers(8, 8, 2, jg.QMODE_TJE, 3, jg.SUB_444) == b""
    assert jg.emit_headers(70000, 8, 3, jg.QMODE_TJE, 3, jg.SUB_444) == b""
    assert jg.emit_headers(8, 8, 1, jg.QMODE_TJE, 3, jg.SUB_420) == b""
    assert jg.max_encoded_size(0, 8, 3) == 0
    assert jg.max_encoded_size(1920, 1080, 3) >= 1024 + 97200 * 416


def test_quantiser_tables_in_headers_follow_the_reference_quirks():
    """Tables are stored in natural order and written raw (jpeg_enc.h:1243, :508)."""
    ql, qc, pl, pc, hl, hc = oracle.oracle_tables(oracle.QMODE_TJE, 1)
    h = jg.emit_headers(8, 8, 3, jg.QMODE_TJE, 1, jg.SUB_444)
    d0 = h.index(b"\xff\xdb")
    assert h[d0 + 5:d0 + 69] == ql.tobytes() and ql[1] == 11 and ql[8] == 12
    d1 = h.index(b"\xff\xdb", d0 + 2)
    assert h[d1 + 5:d1 + 69] == qc.tobytes() and qc[7] == 72       # the "from paper" chroma table
    # tje quality 2 = base/10 floored, min 1
    ql2 = oracle.oracle_tables(oracle.QMODE_TJE, 2)[0]
    assert ql2[0] == 1 and ql2[63] == 9 and ql2.min() == 1


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    L = jg.lib()
    assert L.jpeg_gpu_init(None, 0) == 0
    assert b"CUDA" in L.jpeg_gpu_last_error()
    img = oracle.synth_image(16, 16, 3)
    with pytest.raises(jg.JpegGpuError):
        jg.encode_batch([img])
    # the tje twin returns 0 (error) like the reference's convention, and writes nothing
    calls = []
    cb = jg.WRITE_FUNC(lambda ctx, data, size: calls.append(size))
    assert L.jpeg_gpu_encode_with_func(cb, None, 3, 16, 16, 3, img.ctypes.data) == 0
    assert calls == []


def test_cpp_facade_bmp_round_trip_without_gpu(tmp_path):
    """The façade's BMP feeder (codecs.cpp:255-375 semantics, incl. the reference's own row padding of w % 4):
    read -> flip()/swapBR() -> write(".bmp") -> read back; no GPU involved."""
    import struct, subprocess
    import oracle
    exe = os.path.join(os.path.dirname(jg.LIB_PATH), "write_jpg_like_reference")
    assert os.path.exists(exe), "built by imagecodecs_b200.build"
    px = oracle.synth_image(37, 21, 3)
    h, w, _ = px.shape
    rows = b"".join(px[y].tobytes() + b"\0" * (w % 4) for y in range(h - 1, -1, -1))
    src = tmp_path / "in.bmp"
    src.write_bytes(b"BM" + struct.pack("<IIIIiiHHIIiiII", 54 + len(rows), 0, 54, 40, w, h, 1, 24, 0, len(rows), 0, 0, 0, 0) + rows)
    for ops, want in (("", px), ("f", px[::-1]), ("s", px[:, :, ::-1]), ("fs", px[::-1, :, ::-1])):
        out = tmp_path / ("out_%s.bmp" % ops)
        r = subprocess.run([exe, "--ops", ops, str(src), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = oracle.read_bmp(out.read_bytes())
        assert got is not None and np.array_equal(got, want), ops


def test_restart_headers_match_oracle():
    """JPEG_GPU_FLAG_RESTART adds the DRI segment (one tile of MCUs per interval) in front of SOS."""
    for (w, h, nc, qm, q, sub) in [(395, 348, 3, 0, 3, 0), (64, 64, 3, 1, 75, 1), (100, 50, 1, 1, 85, 0)]:
        got = jg.emit_headers(w, h, nc, qm, q, sub, flags=jg.FLAG_RESTART)
        want = oracle.oracle_headers(w, h, 1 if nc == 1 else 3, sub, qm, q, restart=oracle.restart_interval(nc, sub))
        assert got == want and len(got) == len(jg.emit_headers(w, h, nc, qm, q, sub)) + 6
