#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "batchnorm or small_lora_b4_224_train or conv" > $O/bn_tests.log 2>&1; tail -n 3 $O/bn_tests.log | cut -c1-250
DP_BENCH_DUMP=$O/bn_ev.csv timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/bn_bench.log 2>$O/bn_bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bn_bench.log").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s; bn_stats", d["roofline"]["per_kernel_ms_per_step"].get("bn_stats"))
except Exception as e:
    print("bench failed", e); print(open("$O/bn_bench.err").read()[-800:])
PY
grep "bn_stats" $O/bn_ev.csv | head -3
