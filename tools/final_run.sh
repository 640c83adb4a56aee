#!/bin/bash
# usage (on the GPU box, via gpurun): tools/final_run.sh <tag>
# GPU parity suite, plain bench, ncu launch list + --set full captures of one step (each only after the same command exited 0
# without ncu), then the full bench.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
TAG=${1:-r02}
K='transform_kernel|entropy_kernel|plan_chunks|count_ff|scan_groups|stuff_kernel'
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu --no-twin --no-configs"
$SHORT > gpurun_out/bench_short.json || exit 1
# launch list of the same command: 3 warm-up + 2 timed steps of 6 kernels, then the pass-time runs and the e2e chunks
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 180 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
# one whole step (6 kernels) of the headline configuration, after the warm-up steps
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 18 -c 6 \
    -o gpurun_out/${TAG}_final -f $SHORT > gpurun_out/ncu_full.log 2>&1
# the byte-pinned native twin (tje quality 2, 4:4:4): the six kernels of one step
TW="python tools/prof_case.py --n 64 --qmode 0 --q 2 --sub 0 --steps 3"
$TW > gpurun_out/twin_plain.log && ncu --set full --clock-control none --import-source on -k regex:"$K" -s 18 -c 6 \
    -o gpurun_out/${TAG}_twin444 -f $TW > gpurun_out/ncu_twin.log 2>&1
python bench.py > gpurun_out/bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels"]["transform"]["ms"], d["roofline"]["kernels"]["entropy"]["ms"],
      d["roofline"]["second_pass_ms"], d["e2e"]["value"], d["native_twin"]["value"], d["native_twin"]["roofline_native"]["frac"], d["cpu_baseline"]["value"])
PY
