#!/bin/bash
# A/B of the CUDA-core prediction.3 kernels (DP_PRED_SIMT=0/1) + full GPU suite
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "pred1x1" -p no:cacheprovider 2>&1 | tail -n 3
for v in 0 1; do
  DP_PRED_SIMT=$v DP_BENCH_DUMP=$O/ab7_events_pred$v.csv timeout 600 python bench.py > $O/ab7_bench_pred$v.log 2>&1
  python - <<PY
import json
l=[x for x in open("$O/ab7_bench_pred$v.log") if x.startswith("{")]
j=json.loads(l[-1]); print("DP_PRED_SIMT=$v ms_per_step %.3f value %.0f e2e %.0f frac %.3f" % (j["ms_per_step"], j["value"], j["e2e"]["value"], j["roofline"]["frac"]))
PY
  grep -E "pred|hm_grad|colsum" $O/ab7_events_pred$v.csv | cut -d, -f2,4 | tr '\n' ' '; echo
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/ab7_tests.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" $O/ab7_tests.log | tail -5
