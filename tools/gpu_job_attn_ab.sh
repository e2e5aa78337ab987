#!/bin/bash
# A/B of the T=257 attention kernel with 4 vs 8 softmax warps (DP_ATTN_HALVES=1/2): kernel test, micro-benchmark, bench
O=gpurun_out
mkdir -p $O
for v in 1 2; do
  DP_ATTN_HALVES=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "test_attention" -p no:cacheprovider 2>&1 | tail -n 2
  DP_ATTN_HALVES=$v timeout 300 python tools/attn_tune.py 2>&1 | tail -n 2
  DP_ATTN_HALVES=$v timeout 600 python bench.py > $O/ab5_bench_attn$v.log 2>&1; echo exit=$? >> $O/ab5_bench_attn$v.log
  python - <<PY
import json
l=[x for x in open("$O/ab5_bench_attn$v.log") if x.startswith("{")]
j=json.loads(l[-1]); print("halves=$v", "ms_per_step", j["ms_per_step"], "value", j["value"], "attn ms/step", j["roofline"]["per_kernel_ms_per_step"].get("attention_fwd"))
PY
done
