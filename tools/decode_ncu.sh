#!/bin/bash
# one plain run, then ncu --set full over the decode kernels of one 16-image batch call
cd "$(dirname "$0")/.."
python tools/decode_prof.py 16 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"sync_|decode_intervals|idct_kernel|upsample|color_kernel" -c 40 -o gpurun_out/r01_decode -f python tools/decode_prof.py 16 > gpurun_out/ncu_decode.log 2>&1
tail -1 gpurun_out/ncu_decode.log
