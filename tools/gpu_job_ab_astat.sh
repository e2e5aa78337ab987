#!/bin/bash
# full GPU suite + benchmark step with the A-stationary GEMM off / single / clustered
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider > $O/ab8_tests.log 2>&1; tail -n 5 $O/ab8_tests.log | cut -c1-300
for v in 0 1 2; do DP_GEMM_ASTAT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ab8_bench_astat$v.log 2>&1; python - <<PY
import json
try:
    d=json.loads(open("$O/ab8_bench_astat$v.log").read().strip().splitlines()[-1])
    print("astat=$v", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), d["roofline"]["per_kernel_ms_per_step"])
except Exception as e:
    print("astat=$v failed", e); print(open("$O/ab8_bench_astat$v.log").read()[-1500:])
PY
done
