#!/bin/bash
# A/B of the two-stream half-batch backbone (DP_SPLIT_BATCH=0/1): bench, all configs, GPU tests
O=gpurun_out
mkdir -p $O
for v in 0 1; do
  DP_SPLIT_BATCH=$v timeout 600 python bench.py > $O/ab3_bench_split$v.log 2>&1; echo exit=$? >> $O/ab3_bench_split$v.log
  python - <<PY
import json
l=[x for x in open("$O/ab3_bench_split$v.log") if x.startswith("{")]
j=json.loads(l[-1]); print("split=$v", "ms_per_step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"])
PY
  DP_SPLIT_BATCH=$v timeout 600 python tools/run_configs.py cfg0 cfg2 cfg3 "infer b64 448" f4 > $O/ab3_cfg_split$v.log 2>&1
  cat $O/ab3_cfg_split$v.log | cut -c1-150
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/ab3_tests.log 2>&1; echo exit=$? >> $O/ab3_tests.log
grep -E "passed|failed|^FAILED|^ERROR" $O/ab3_tests.log | tail -8
