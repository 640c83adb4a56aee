"""First-contact GPU check: encode a spread of cases on cuda:0 and compare every stage with
the oracle.  Prints a compact report; meant for `gpurun -- python tools/gpu_check.py`."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import imagecodecs_b200 as jg
from oracle import oracle_stages, oracle_headers, synth_batch, QMODE_TJE, QMODE_IJG

def run_case(label, batch, qm, q, sub, win=0):
    n, h, w, c = batch.shape
    dev = torch.from_numpy(batch).cuda()
    imgs = [dev[i] for i in range(n)]
    plan = jg.Plan.for_arrays(imgs, qm, q, sub, device=0, win_words=win)
    nb = plan.num_blocks
    coefs = torch.zeros(nb * 64, dtype=torch.int16, device="cuda")
    bits = torch.zeros(nb, dtype=torch.int32, device="cuda")
    plan.attach_debug(coefs.data_ptr(), bits.data_ptr())
    s = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    plan.run(s)
    torch.cuda.synchronize()
    dt = time.time() - t0
    files = plan.fetch(s)
    coefs = coefs.cpu().numpy().reshape(n, -1, 64); bits = bits.cpu().numpy().astype(np.uint32).reshape(n, -1)
    ok = True
    for i in range(n):
        st = oracle_stages(batch[i], qm, q, sub)
        ce = np.array_equal(coefs[i], st["coefs"]); be = np.array_equal(bits[i], st["block_bits"])
        fe = files[i] == st["jpeg"]
        if not (ce and be and fe):
            ok = False
            print("  MISMATCH img", i, "coefs", ce, "bits", be, "file", fe, len(files[i] or b""), len(st["jpeg"]))
            if not ce:
                d = np.argwhere(coefs[i] != st["coefs"]); print("   coef diffs:", len(d), d[:5].tolist(),
                      coefs[i][tuple(d[0])], st["coefs"][tuple(d[0])])
            elif not be:
                d = np.nonzero(bits[i] != st["block_bits"])[0]; print("   bits diffs:", len(d), d[:8])
            elif files[i] is not None:
                a = np.frombuffer(files[i], np.uint8); b = np.frombuffer(st["jpeg"], np.uint8)
                m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
                print("   first byte diffs:", d[:8], "of", m)
    print("%-28s %-22s q=(%d,%d) sub=%d win=%d %s  %.1f ms" % (label, batch.shape, qm, q, sub, win, "ok" if ok else "FAIL", dt * 1e3))
    plan.close()
    return ok

def main():
    print("devices:", jg.init())
    B = synth_batch
    allok = True
    cases = [
        ("1 mcu", B(1, 8, 8, 3), 0, 3, 0, 0),
        ("edge 17x13", B(1, 17, 13, 3), 0, 3, 0, 0),
        ("rgba 17x13", B(1, 17, 13, 4), 0, 1, 0, 0),
        ("64x64 noise", B(1, 64, 64, 3, "noise"), 0, 3, 0, 0),
        ("multi-tile x2", B(2, 200, 120, 3), 0, 2, 0, 0),
        ("unaligned 395x348 x3", B(3, 395, 348, 3), 0, 3, 0, 0),
        ("512x512 q2 x5", B(5, 512, 512, 3), 0, 2, 0, 0),
        ("noise 256 tiny window", B(1, 256, 256, 3, "noise"), 0, 3, 0, 64),
        ("multi-group", B(2, 200, 120, 3), 0, 3, 0, 128),
        ("420 q75", B(2, 200, 120, 3), 1, 75, 1, 0),
        ("420 edge", B(1, 33, 47, 3), 1, 75, 1, 0),
        ("420 rgba", B(1, 33, 47, 4), 1, 90, 1, 0),
        ("gray", B(2, 200, 130, 1), 1, 85, 0, 0),
        ("1080p q1", B(1, 1920, 1080, 3), 0, 1, 0, 0),
        ("1080p q3", B(2, 1920, 1080, 3), 0, 3, 0, 0),
        ("1080p noise q3", B(1, 1920, 1080, 3, "noise"), 0, 3, 0, 0),
        ("1080p 420 q75 x4", B(4, 1920, 1080, 3), 1, 75, 1, 0),
        ("2048 gray q85", B(1, 2048, 2048, 1), 1, 85, 0, 0),
    ]
    for c in cases:
        try:
            allok &= run_case(*c)
        except Exception as e:
            allok = False
            print("EXC", c[0], repr(e))
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1

if __name__ == "__main__":
    sys.exit(main())
