"""Kernel-variant builds for A/B runs on the GPU box: python tools/build_variants.py name=-DFLAG[,-DFLAG2] ...
Each variant becomes variants/<name>.so (git-ignored, travels with gpurun); run with JPEG_GPU_LIB=variants/<name>.so."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagecodecs_b200 import build
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for arg in sys.argv[1:]:
    name, flags = arg.split("=", 1)
    out = os.path.join(ROOT, "variants", name + ".so")
    build.build_variant(out, [f for f in flags.split(",") if f], os.path.join("/tmp", "jgvar_" + name))
    print("built", out)
