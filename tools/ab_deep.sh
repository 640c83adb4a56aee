#!/bin/bash
cd "$(dirname "$0")/.."
for v in variants/libjpeg_gpu_never.so variants/libjpeg_gpu_always.so; do
  echo "== $v"
  for n in 1 4 16 32 64 128; do JPEG_GPU_LIB=$PWD/$v python tools/prof_case.py --n $n --qmode 1 --q 75 --sub 1 --steps 10; done
  JPEG_GPU_LIB=$PWD/$v python tools/prof_case.py --w 3840 --h 2160 --n 16 --qmode 1 --q 90 --sub 0 --steps 5
  JPEG_GPU_LIB=$PWD/$v python tools/prof_case.py --w 3840 --h 2160 --n 128 --qmode 1 --q 90 --sub 0 --steps 3
done
