#!/bin/bash
O=gpurun_out; mkdir -p $O
for v in 1024 512; do DP_HEAD_TILE_KMIN=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/mix_bench_$v.log 2>$O/mix_bench_$v.err; python - <<PY
import json
try:
    d=json.loads(open("$O/mix_bench_$v.log").read().strip().splitlines()[-1])
    print("kmin=$v", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), d["roofline"]["per_kernel_ms_per_step"]["gemm_kmajor_tcgen05"])
except Exception as e:
    print("bench $v failed", e); print(open("$O/mix_bench_$v.err").read()[-600:])
PY
done
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_bench_shape_gpu.py tests/test_submodules_gpu.py -m gpu -x -q --timeout 120 -p no:cacheprovider > $O/mix_tests.log 2>&1; tail -n 3 $O/mix_tests.log | cut -c1-300
