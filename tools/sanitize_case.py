"""Small encode cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imagecodecs_b200 as jg
import oracle

jg.init([0])
ok = True
cases = [((3, 200, 120, 3, "photo"), 0, 2, 0, 0), ((2, 203, 117, 4, "photo"), 0, 3, 0, 0), ((2, 200, 120, 3, "photo"), 1, 75, 1, 0),
         ((1, 131, 67, 3, "photo"), 1, 90, 1, 0), ((2, 200, 130, 1, "photo"), 1, 85, 0, 0), ((1, 96, 96, 3, "noise"), 0, 3, 0, 0),
         ((1, 160, 96, 3, "noise"), 0, 3, 0, 64), ((1, 640, 480, 3, "photo"), 1, 75, 1, 0)]
for (n, w, h, c, kind), qm, q, sub, win in cases:
    b = oracle.synth_batch(n, w, h, c, kind)
    files, st = jg.encode_batch([b[i] for i in range(n)], qm, q, sub, device=0, win_words=win)
    good = all(files[i] == oracle.oracle_encode(b[i], qm, q, sub) for i in range(n))
    ok &= good
    print((n, w, h, c, kind), qm, q, sub, win, "ok" if good else "MISMATCH")
jg.lib().jpeg_gpu_shutdown()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
