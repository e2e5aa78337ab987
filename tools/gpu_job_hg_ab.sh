#!/bin/bash
# A/B of the hourglass down/up branch on side stream 2 (DP_HG_STREAMS=0/1): bench, configs, GPU tests
O=gpurun_out
mkdir -p $O
for v in 0 1; do
  DP_HG_STREAMS=$v timeout 600 python bench.py > $O/ab4_bench_hg$v.log 2>&1; echo exit=$? >> $O/ab4_bench_hg$v.log
  python - <<PY
import json
l=[x for x in open("$O/ab4_bench_hg$v.log") if x.startswith("{")]
j=json.loads(l[-1]); print("hg=$v", "ms_per_step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"])
PY
  DP_HG_STREAMS=$v timeout 600 python tools/run_configs.py cfg0 cfg2 cfg3 > $O/ab4_cfg_hg$v.log 2>&1
  cat $O/ab4_cfg_hg$v.log | cut -c1-150
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/ab4_tests.log 2>&1; echo exit=$? >> $O/ab4_tests.log
grep -E "passed|failed|^FAILED|^ERROR" $O/ab4_tests.log | tail -8
