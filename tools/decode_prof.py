"""N x 1080p streams decoded in one call (for the ncu launch list): python tools/decode_prof.py [N] [free]
(free: streams without restart markers -> subsequence decode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagecodecs_b200 as jg
from imagecodecs_b200.synth import synth_batch
jg.init([0])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
imgs = synth_batch(N, 1920, 1080, 3, "photo").numpy()
files, st = jg.encode_batch([imgs[i] for i in range(N)], 1, 75, 1, device=0, flags=0 if 'free' in sys.argv[2:] else jg.FLAG_RESTART)
out, ms = jg.decode_batch(files, timed=True)
print("kernels %.3f ms" % ms)
