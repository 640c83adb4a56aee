#!/bin/bash
# usage (on the GPU box, via gpurun): tools/decode_free_run.sh
# decoder tests first, decode timings with and without restart markers, the whole GPU suite, smoke(), then the ncu
# launch list of one restart-free batch decode (after the same command ran plain).  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q -k "decoder or decode" > gpurun_out/dec_tests.log 2>&1; tail -2 gpurun_out/dec_tests.log
timeout 100 python tools/decode_case.py > gpurun_out/decode_case.log 2>&1; cat gpurun_out/decode_case.log
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -2 gpurun_out/gputests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 40 python tools/decode_prof.py 16 free > gpurun_out/decode_free_plain.log 2>&1 && \
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sync_|decode_intervals|idct_kernel|upsample|color_kernel" -c 300 --csv \
    --log-file gpurun_out/decode_free_launches.csv python tools/decode_prof.py 16 free > gpurun_out/ncu_decode_free.log 2>&1
cat gpurun_out/decode_free_plain.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_decode_legs.json 2> gpurun_out/bench_decode_legs.err; tail -c 1500 gpurun_out/bench_decode_legs.json
