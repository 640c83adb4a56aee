"""Run one encode configuration a few times (for ncu / quick timing).
usage: python tools/prof_case.py --w 1920 --h 1080 --n 64 --qmode 1 --q 75 --sub 1 [--kind photo] [--steps 5]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imagecodecs_b200 as jg
from imagecodecs_b200.synth import synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1920); ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--n", type=int, default=64); ap.add_argument("--nc", type=int, default=3)
ap.add_argument("--qmode", type=int, default=1); ap.add_argument("--q", type=int, default=75)
ap.add_argument("--sub", type=int, default=1); ap.add_argument("--kind", default="photo")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--restart", action="store_true", help="JPEG_GPU_FLAG_RESTART: one restart interval per tile")
ap.add_argument("--mixed", action="store_true", help="image i is 16*(i%%5) rows shorter: different tile counts in one launch")
a = ap.parse_args()
jg.init([0])
px = synth_batch(a.n, a.w, a.h, a.nc, a.kind, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
torch.cuda.synchronize()
imgs = [px[i][:a.h - 16 * (i % 5)].contiguous() if a.mixed else px[i] for i in range(a.n)]
plan = jg.Plan.for_arrays(imgs, a.qmode, a.q, a.sub, device=0, flags=jg.FLAG_RESTART if a.restart else 0)
for _ in range(3): plan.run(sp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(a.steps): plan.run(sp)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
mp = sum(int(im.shape[0]) * a.w for im in imgs) / 1e6
sz = sum(plan.encoded_size(i) for i in range(a.n))
plan.enable_timing(True); plan.run(sp); torch.cuda.synchronize(); ta, tb, tc = plan.pass_times(); plan.enable_timing(False)
print("%s %dx%dx%d n=%d qmode=%d q=%d sub=%d %s%s: %.3f ms/step (transform %.3f + entropy %.3f + stuff %.3f)  %.1f GP/s  out %.3f B/px  roofline %.4f" % (
    "fused" if plan.fused else "split", a.w, a.h, a.nc, a.n, a.qmode, a.q, a.sub, a.kind, (" mixed" if a.mixed else "") + (" restart" if a.restart else ""), ms, ta, tb, tc, mp / ms, sz / (mp * 1e6),
    (mp * 1e6 * a.nc + sz) / (ms * 1e-3) / 6550.1e9))
