#!/bin/bash
# usage: tools/ab_small.sh [variant.so ...] -- two quick timing cases (1080p 4:2:0 q75 x256, tje-2 4:4:4 x128), each run twice
cd "$(dirname "$0")/.."
run_cases() {
  for r in 1 2; do
    python tools/prof_case.py --n 256 --qmode 1 --q 75 --sub 1 --steps 10
    python tools/prof_case.py --n 128 --qmode 0 --q 2 --sub 0 --steps 5
  done
}
echo "== in-tree"; run_cases
for v in "$@"; do echo "== $v"; JPEG_GPU_LIB=$PWD/$v run_cases; done
