#!/bin/bash
# usage: tools/ab_few.sh [variant.so ...] -- launches of few images with tiles of 1025..1536 symbols (the few_images rule of code_tile)
cd "$(dirname "$0")/.."
run_cases() {
  python tools/prof_case.py --w 16384 --h 16384 --n 1 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 2 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 8 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 16 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 32 --qmode 0 --q 2 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 32 --qmode 1 --q 90 --sub 0 --steps 5
  python tools/prof_case.py --w 3840 --h 2160 --n 16 --qmode 1 --q 90 --sub 0 --steps 5
  python tools/prof_case.py --n 128 --qmode 0 --q 2 --sub 0 --steps 5
}
echo "== in-tree"; run_cases
for v in "$@"; do echo "== $v"; JPEG_GPU_LIB=$PWD/$v run_cases; done
