#!/bin/bash
# usage: tools/ncu_quick.sh <tag> [lib.so] -- one plain run, then one ncu --set full capture of the pass-1 kernels
# (transform + entropy, or the fused encode kernel) of one step.  NCU_ARGS overrides the case, NCU_K the kernel regex.
cd "$(dirname "$0")/.."
[ -n "$2" ] && export JPEG_GPU_LIB=$PWD/$2
ARGS="${NCU_ARGS:---n 128 --qmode 1 --q 75 --sub 1 --steps 3}"
K="${NCU_K:-transform_kernel|entropy_kernel|encode_tiles}"
N="${NCU_C:-2}"
python tools/prof_case.py $ARGS || exit 1
# 3 warm-up runs + "steps" timed runs come first: skip 3 steps' worth of matching launches
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $((3 * N)) -c $N -o gpurun_out/$1 -f python tools/prof_case.py $ARGS > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log
