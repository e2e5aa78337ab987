#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 60 -p no:cacheprovider -k "fused_lora or row_owning or test_lora" > $O/rowln_tests.log 2>&1; tail -n 15 $O/rowln_tests.log | cut -c1-300
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_parity_bench_shape_gpu.py -m gpu -q -x --timeout 120 -p no:cacheprovider > $O/rowln_tests2.log 2>&1; tail -n 5 $O/rowln_tests2.log | cut -c1-300
for v in 0 1; do DP_LORA_FUSED=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/rowln_bench$v.log 2>/dev/null; python - <<PY
import json
try:
    d=json.loads(open("$O/rowln_bench$v.log").read().strip().splitlines()[-1])
    print("lora_fused=$v", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), d["roofline"]["per_kernel_ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
done
