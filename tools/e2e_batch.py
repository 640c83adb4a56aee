"""Time jpeg_gpu_encode_batch (the one-call API) with pinned / pageable host buffers."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagecodecs_b200 as jg
from imagecodecs_b200.synth import synth_batch
jg.init([0])
N, W, H = 256, 1920, 1080
px = synth_batch(N, W, H, 3, device="cuda").cpu()
for pinned in (False, True):
    host = px.pin_memory() if pinned else px
    cap = 2 * 1024 * 1024
    out = torch.empty((N, cap), dtype=torch.uint8)
    if pinned: out = out.pin_memory()
    descs = (jg.Image * N)(*[jg.Image(host[i].data_ptr(), W, H, 3, 0, 1, 75, 1, 0) for i in range(N)])
    outs = (jg.Output * N)(*[jg.Output(out[i].data_ptr(), cap, 0, 0) for i in range(N)])
    opts = jg.BatchOpts(0, 0, None, 0)
    L = jg.lib()
    for it in range(4):
        t0 = time.perf_counter()
        ok = L.jpeg_gpu_encode_batch(descs, N, outs, C.byref(opts))
        dt = time.perf_counter() - t0
        print("pinned=%s iter %d: ok=%d %.1f ms  %.1f GP/s" % (pinned, it, ok, dt * 1e3, N * W * H / 1e9 / dt))
