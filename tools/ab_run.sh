#!/bin/bash
# usage: tools/ab_run.sh [variant.so ...] -- runs the standard timing cases for the in-tree lib (split pipeline, then the fused
# round-1 kernel) and each variant
cd "$(dirname "$0")/.."
run_cases() {
  python tools/prof_case.py --n 256 --qmode 1 --q 75 --sub 1 --steps 10
  python tools/prof_case.py --n 128 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --n 64 --qmode 0 --q 3 --sub 0 --steps 5
  python tools/prof_case.py --n 64 --qmode 0 --q 1 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 32 --qmode 1 --q 90 --sub 0 --steps 5
  python tools/prof_case.py --w 16384 --h 16384 --n 1 --nc 1 --qmode 1 --q 85 --sub 0 --steps 5
  python tools/prof_case.py --w 512 --h 512 --n 2048 --qmode 1 --q 75 --sub 1 --steps 5
  python tools/prof_case.py --n 256 --qmode 1 --q 75 --sub 1 --steps 5 --mixed
  python tools/prof_case.py --n 32 --qmode 0 --q 3 --sub 0 --steps 3 --kind noise
}
echo "== in-tree"; run_cases
if [ -z "$NO_FUSED" ]; then echo "== in-tree, fused"; JPEG_GPU_PIPELINE=fused run_cases; fi
for v in "$@"; do echo "== $v"; JPEG_GPU_LIB=$PWD/$v run_cases; done
