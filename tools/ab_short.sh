#!/bin/bash
# usage: tools/ab_short.sh [variant.so ...] -- quick cases (4:2:0 q75, tje-2, tje-3, 4K q90 x 32, 16k gray) for the in-tree lib and each variant
cd "$(dirname "$0")/.."
run_cases() {
  python tools/prof_case.py --n 256 --qmode 1 --q 75 --sub 1 --steps 10
  python tools/prof_case.py --n 128 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --n 64 --qmode 0 --q 3 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 32 --qmode 1 --q 90 --sub 0 --steps 5
  python tools/prof_case.py --w 16384 --h 16384 --n 1 --nc 1 --qmode 1 --q 85 --sub 0 --steps 5
}
echo "== in-tree"; run_cases
for v in "$@"; do echo "== $v"; JPEG_GPU_LIB=$PWD/$v run_cases; done
