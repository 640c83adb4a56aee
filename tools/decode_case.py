"""Decode timing: encode one synthetic image on the GPU (with / without restart intervals), decode it with
jpeg_gpu_decode, compare with nothing here (tests do that) and print kernel and call times.
usage: python tools/decode_case.py [--w 1920 --h 1080 --q 75 --sub 1]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imagecodecs_b200 as jg
from imagecodecs_b200.synth import synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1920); ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--q", type=int, default=75); ap.add_argument("--sub", type=int, default=1)
a = ap.parse_args()
jg.init([0])
img = synth_batch(1, a.w, a.h, 3, "photo")[0].numpy()
for flags, name in ((jg.FLAG_RESTART, "restart intervals (parallel entropy decode)"), (0, "no restart markers (subsequence decode)")):
    files, st = jg.encode_batch([img], 1, a.q, a.sub, device=0, flags=flags)
    jpeg = files[0]
    jg.decode(jpeg)                                       # warm-up
    ks, calls = [], []
    for _ in range(10):
        t0 = time.perf_counter(); px, ms = jg.decode(jpeg, timed=True); calls.append((time.perf_counter() - t0) * 1e3); ks.append(ms)
    mp = a.w * a.h / 1e6
    print("%dx%d q%d sub%d, %s: %d bytes; kernels %.3f ms (%.0f MP/s), whole call %.3f ms (%.0f MP/s)" % (
        a.w, a.h, a.q, a.sub, name, len(jpeg), np.median(ks), mp / np.median(ks) * 1e3, np.median(calls), mp / np.median(calls) * 1e3))

# a batch in one call: the GPU is filled by images x intervals (or x subsequences)
for n in (16, 256):
    imgs = synth_batch(n, a.w, a.h, 3, "photo").numpy()
    for flags, name in ((jg.FLAG_RESTART, "restart streams"), (0, "restart-free streams")):
        files, st = jg.encode_batch([imgs[i] for i in range(n)], 1, a.q, a.sub, device=0, flags=flags)
        jg.decode_batch(files[:2])
        t0 = time.perf_counter(); out, ms = jg.decode_batch(files, timed=True); call = (time.perf_counter() - t0) * 1e3
        mp = n * a.w * a.h / 1e6
        print("%d x %dx%d q%d sub%d %s in one call: kernels %.3f ms (%.0f MP/s), whole call incl. host parse + copies %.1f ms (%.0f MP/s)" % (
            n, a.w, a.h, a.q, a.sub, name, ms, mp / ms * 1e3, call, mp / call * 1e3))
