#!/bin/bash
# usage: tools/ncu_times.sh <tag> <prof_case args...> -- per-launch durations (ncu, serialised, cold) of one step's kernels
cd "$(dirname "$0")/.."
tag=$1; shift
python tools/prof_case.py "$@" --steps 2 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stuff|count_ff|scan_groups|plan_chunks|transform_k|entropy_k" -s ${NCU_SKIP:-18} -c 12 --csv --log-file gpurun_out/times_$tag.csv python tools/prof_case.py "$@" --steps 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/times_$tag.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:50], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
