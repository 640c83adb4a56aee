"""Every BASELINE.json configuration, device-timed, on 1..8 GPUs of one box.

    python tools/bench_configs.py [--configs 1,2,3,4,5] [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --configs 3,5

bench.py (the driver's contract) measures configs[1]; this tool reports the same quantities for
all five shapes: MP/s over all ranks (CUDA events on the launching stream, max over ranks, inputs
resident in HBM), the fraction of the measured HBM roofline (input bytes + compressed scan bytes
over the step time), the byte-pinned native twin of each shape, and a parity spot check against
the oracle / the compiled reference.  Configs 3 and 5 are sharded by image index over the ranks
(strong scaling: the batch is fixed), the others run on rank 0's GPU only.
Prints one JSON line per (config, variant) and a markdown table; rank 0 writes both to gpurun_out/.
"""
import argparse, ctypes as C, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import imagecodecs_b200 as jg
from imagecodecs_b200.synth import synth_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    want = [int(c) for c in args.configs.split(",")]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    jg.init([local])
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    peak, peak_src = bench.measured_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def measure(imgs, qm, q, sub, active):
        """Device-timed step of this rank's share; returns (ms max over ranks, total MP, total bytes)."""
        ms = 0.0; mp = 0.0; by = 0.0
        plan = None
        if active and imgs:
            plan = jg.Plan.for_arrays(imgs, qm, q, sub, device=0)
            for _ in range(2): plan.run(sptr)
        barrier()
        if plan is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps): plan.run(sptr)
            e1.record(stream); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            px = sum(int(im.shape[0]) * int(im.shape[1]) for im in imgs)
            nc = [1 if im.dim() == 2 or im.shape[2] == 1 else int(im.shape[2]) for im in imgs]
            per = lambda v, i: v[i] if isinstance(v, (list, tuple)) else v
            hdr = [len(jg.emit_headers(int(im.shape[1]), int(im.shape[0]), nc[i], per(qm, i), per(q, i), per(sub, i))) for i, im in enumerate(imgs[:3])]
            out = sum(plan.encoded_size(i) for i in range(len(imgs))) - len(imgs) * hdr[0]
            mp = px / 1e6
            by = sum(int(im.numel()) for im in imgs) + out
        barrier()
        ms = reduce(ms, dist.ReduceOp.MAX if world > 1 else None)
        mp = reduce(mp, dist.ReduceOp.SUM if world > 1 else None)
        by = reduce(by, dist.ReduceOp.SUM if world > 1 else None)
        return plan, ms, mp, by

    rows = []

    def report(cfg, name, variant, imgs, qm, q, sub, sharded, check=None):
        active = sharded or rank == 0
        plan, ms, mp, by = measure(imgs, qm, q, sub, active)
        parity = None
        if rank == 0 and plan is not None and check is not None:
            parity = check(plan)
        if plan is not None:
            plan.close()
        if rank == 0:
            n_used = world if sharded else 1
            line = {"config": cfg, "workload": name, "variant": variant, "n_gpus": n_used, "ms_per_step": round(ms, 4),
                    "value": round(mp / (ms * 1e-3), 1), "unit": "MP/s", "roofline_frac_of_measured_hbm": round(by / (ms * 1e-3) / 1e9 / (peak * n_used), 4),
                    "bytes_per_px": round(by / (mp * 1e6), 3), "parity_spot_check": parity, "steps": args.steps}
            print(json.dumps(line), flush=True)
            rows.append(line)

    import oracle

    def check_first(arr, qm, q, sub, ref=False):
        def f(plan):
            got = plan.fetch(sptr)[0]
            host = arr.cpu().numpy()
            if ref:
                rc, want_b = oracle.ref_encode(host, q)
                return rc == 1 and got == want_b
            return got == oracle.oracle_encode(host, qm, q, sub)
        return f

    share = lambda n: (rank * n // world, (rank + 1) * n // world)

    if 1 in want:
        # config 1: data/cat.bmp through the codecs.h path == tje quality 3 on the BGR-as-RGB pixels; the fixture travels in tests/golden
        fx = np.load(os.path.join(ROOT, "tests", "golden", "fixture_pixels.npz"))
        cat = torch.from_numpy(fx["cat_bgr"]).to(dev)
        for tq in (3, 2, 1):
            report(1, "cat.bmp 395x348 via readBmp (BGR-as-RGB)", "tje-%d 4:4:4 (native, byte-pinned)" % tq, [cat], 0, tq, 0, False,
                   check_first(cat, 0, tq, 0, ref=True))
    if 2 in want:
        px = synth_batch(256, 1920, 1080, 3, "photo", seed=1, first=0, device=dev) if rank == 0 else None
        imgs = [px[i] for i in range(256)] if rank == 0 else []
        report(2, "256 x 1920x1080 RGB", "IJG q75 4:2:0 (extended)", imgs, 1, 75, 1, False, check_first(px[0], 1, 75, 1) if rank == 0 else None)
        report(2, "256 x 1920x1080 RGB", "twin tje-2 4:4:4 (native)", imgs, 0, 2, 0, False, check_first(px[0], 0, 2, 0, ref=True) if rank == 0 else None)
        del px, imgs
    if 3 in want:
        lo, hi = share(128)
        px = synth_batch(hi - lo, 3840, 2160, 3, "photo", seed=1, first=lo, device=dev)
        imgs = [px[i] for i in range(hi - lo)]
        report(3, "128 x 3840x2160 RGB, sharded by image index", "IJG q90 4:4:4 (extended)", imgs, 1, 90, 0, True, check_first(px[0], 1, 90, 0))
        report(3, "128 x 3840x2160 RGB, sharded by image index", "twin tje-3 4:4:4 (native)", imgs, 0, 3, 0, True, check_first(px[0], 0, 3, 0, ref=True))
        del px, imgs
    if 4 in want:
        if rank == 0:
            g = synth_batch(1, 16384, 16384, 1, "photo", seed=1, first=0, device=dev)
            report(4, "1 x 16384x16384 gray", "IJG q85 (extended)", [g[0]], 1, 85, 0, False)
            del g
            c = synth_batch(1, 16384, 16384, 3, "photo", seed=1, first=0, device=dev)
            report(4, "1 x 16384x16384 RGB", "twin tje-2 4:4:4 (native)", [c[0]], 0, 2, 0, False)
            del c
        else:
            report(4, "", "", [], 1, 85, 0, False); report(4, "", "", [], 0, 2, 0, False)
    if 5 in want:
        lo, hi = share(16384)
        px = synth_batch(hi - lo, 512, 512, 3, "photo", seed=1, first=lo, device=dev)
        imgs = [px[i] for i in range(hi - lo)]
        qs = [(50, 75, 95)[(lo + i) % 3] for i in range(hi - lo)]
        tq = [(1, 2, 3)[(lo + i) % 3] for i in range(hi - lo)]
        report(5, "16384 x 512x512 RGB, sharded by image index", "IJG q in {50,75,95} by n%3, 4:2:0 (extended)", imgs, 1, qs, 1, True,
               check_first(px[0], 1, qs[0], 1))
        report(5, "16384 x 512x512 RGB, sharded by image index", "twin tje {1,2,3} by n%3, 4:4:4 (native)", imgs, 0, tq, 0, True,
               check_first(px[0], 0, tq[0], 0, ref=True))
        del px, imgs

    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = "configs_%dgpu" % world
        with open(os.path.join(ROOT, "gpurun_out", tag + ".jsonl"), "w") as fh:
            for r in rows: fh.write(json.dumps(r) + "\n")
        md = ["| config | workload | variant | GPUs | ms/step | MP/s | frac of HBM peak (%.0f GB/s x GPUs) | B/px moved | parity |" % peak,
              "|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            md.append("| %d | %s | %s | %d | %.3f | %.0f | %.4f | %.2f | %s |" % (r["config"], r["workload"], r["variant"], r["n_gpus"], r["ms_per_step"],
                                                                               r["value"], r["roofline_frac_of_measured_hbm"], r["bytes_per_px"], r["parity_spot_check"]))
        open(os.path.join(ROOT, "gpurun_out", tag + ".md"), "w").write("\n".join(md) + "\n")
        print("\n".join(md))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
