#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s --timeout 120 -p no:cacheprovider -k "test_lora_fwd_bwd or fused_lora" > $O/lora_tests.log 2>&1; grep -E "lora bwd|passed|failed|Error|assert" $O/lora_tests.log | tail -n 12 | cut -c1-250
if [ "${BENCH:-1}" = "1" ]; then
for v in 1 0; do
DP_LORA_BWD_MMA=$v DP_BENCH_DUMP=$O/lora_ev_$v.csv timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/lora_bench_$v.log 2>$O/lora_bench_$v.err
python - <<PY
import json
try:
    d=json.loads(open("$O/lora_bench_$v.log").read().strip().splitlines()[-1])
    print("mma=$v", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s; lora_bwd", d["roofline"]["per_kernel_ms_per_step"].get("lora_bwd"))
except Exception as e:
    print("bench failed", e); print(open("$O/lora_bench_$v.err").read()[-800:])
PY
grep "lora_bwd" $O/lora_ev_$v.csv | head -2
done
fi
