#!/bin/bash
# usage: tools/ncu_quick.sh <tag> [lib.so] -- one plain run, then one ncu --set full capture of the encode kernel (1080p 4:2:0 q75 x128)
cd "$(dirname "$0")/.."
[ -n "$2" ] && export JPEG_GPU_LIB=$PWD/$2
ARGS="${NCU_ARGS:---n 128 --qmode 1 --q 75 --sub 1 --steps 3}"
python tools/prof_case.py $ARGS || exit 1
ncu --set full --clock-control none --import-source on -k regex:"encode_tiles" -s 3 -c 1 -o gpurun_out/$1 -f python tools/prof_case.py $ARGS > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log
