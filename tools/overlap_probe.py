"""Does running sub-batches on several streams help (pass A of one sub-batch beside pass B of another)?
usage: python tools/overlap_probe.py [--n 256] [--parts 2] [--qmode 1 --q 75 --sub 1]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imagecodecs_b200 as jg
from imagecodecs_b200.synth import synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256); ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--qmode", type=int, default=1); ap.add_argument("--q", type=int, default=75); ap.add_argument("--sub", type=int, default=1)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
jg.init([0])
px = synth_batch(a.n, 1920, 1080, 3, "photo", device="cuda")
torch.cuda.synchronize()
for parts in sorted({1, a.parts, 2 * a.parts}):
    per = a.n // parts
    streams = [torch.cuda.Stream() for _ in range(parts)]
    plans = [jg.Plan.for_arrays([px[i] for i in range(k * per, (k + 1) * per)], a.qmode, a.q, a.sub, device=0) for k in range(parts)]
    def step():
        for pl, st in zip(plans, streams): pl.run(C.c_void_p(st.cuda_stream))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for st in streams: st.wait_event(e0)
    for _ in range(a.steps): step()
    for st in streams:
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print("parts=%d (%d images each, one stream per part): %.3f ms/step  %.1f GP/s" % (parts, per, ms, a.n * 1920 * 1080 / 1e6 / ms))
    for pl in plans: pl.close()
