#!/bin/bash
# A/B of the backward stream overlap (weight-gradient GEMMs on the second stream): bench + f4 config with DP_BWD_OVERLAP=0/1
O=gpurun_out
mkdir -p $O
for v in 0 1; do
  DP_BWD_OVERLAP=$v timeout 600 python bench.py > $O/ab2_bench_overlap$v.log 2>&1; echo exit=$? >> $O/ab2_bench_overlap$v.log
  DP_BWD_OVERLAP=$v timeout 300 python tools/run_configs.py f4 > $O/ab2_f4_overlap$v.log 2>&1
  python - <<PY
import json
l=[x for x in open("$O/ab2_bench_overlap$v.log") if x.startswith("{")]
j=json.loads(l[-1]); print("overlap=$v", "ms_per_step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["value"])
PY
  tail -n 1 $O/ab2_f4_overlap$v.log
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/ab2_tests.log 2>&1; echo exit=$? >> $O/ab2_tests.log
grep -E "passed|failed|^FAILED|^ERROR" $O/ab2_tests.log | tail -8
