#!/bin/bash
# usage: VAR=DP_PRODUCER_STATS VALUES="1 0 1 0" bash tools/gpu_job_ab_env.sh  -- same-box A/B of one environment switch on the default bench
O=gpurun_out; mkdir -p $O
for v in $VALUES; do
env $VAR=$v timeout 300 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/ab_env.log 2>$O/ab_env.err
python - <<PY
import json
try:
    d=json.loads(open("$O/ab_env.log").read().strip().splitlines()[-1])
    print("$VAR=$v", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s", d.get("ms_per_step_repeats"))
except Exception as e:
    print("bench failed", e); print(open("$O/ab_env.err").read()[-800:])
PY
done
