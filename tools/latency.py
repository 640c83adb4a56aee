"""Latency of the single-image drop-in call (jpeg_gpu_encode_with_func) vs the CPU reference."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imagecodecs_b200 as jg, oracle
jg.init([0]); L = jg.lib()
cat = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixture_pixels.npz"))["cat_bgr"]
for name, img in (("cat.bmp 395x348", cat), ("1080p", oracle.synth_image(1920, 1080, 3))):
    img = np.ascontiguousarray(img); h, w, c = img.shape
    n = [0]; cb = jg.WRITE_FUNC(lambda ctx, d, s: n.__setitem__(0, n[0] + s))
    for q in (3,):
        ts = []
        for it in range(12):
            n[0] = 0; t0 = time.perf_counter()
            ok = L.jpeg_gpu_encode_with_func(cb, None, q, w, h, c, img.ctypes.data)
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); rc, ref = oracle.ref_encode(img, q); tr = time.perf_counter() - t0
        print("%s q%d: gpu call median %.3f ms (first %.1f ms), %d bytes; CPU reference %.1f ms" % (name, q, 1e3 * sorted(ts)[6], 1e3 * ts[0], n[0], 1e3 * tr))
