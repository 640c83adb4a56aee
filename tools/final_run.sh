#!/bin/bash
# usage (on the GPU box, via gpurun): tools/final_run.sh <tag>
# GPU parity suite, plain bench, ncu launch list + one --set full capture of one step (each only after the
# same command exited 0 without ncu), then the full bench.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu --no-twin > gpurun_out/bench_short.json || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_tiles|plan_chunks|stuff_kernel" -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-twin > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_tiles|plan_chunks|stuff_kernel" -s 9 -c 3 \
    -o gpurun_out/${TAG}_final -f python bench.py --steps 2 --warmup 3 --no-cpu --no-twin > gpurun_out/ncu_full.log 2>&1
python bench.py > gpurun_out/bench_full.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["second_pass_ms"],
      d["e2e"]["value"], d["native_twin"]["value"], d["native_twin"]["roofline_frac"], d["cpu_baseline"]["value"])
PY
