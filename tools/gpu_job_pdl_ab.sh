#!/bin/bash
# A/B of programmatic dependent launch and the uniform shared-memory carve-out (csrc/launch.cuh) on the benchmark step.
TAG=${1:-ab}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > $O/${TAG}_tests.log 2>&1; echo exit=$? >> $O/${TAG}_tests.log
grep -E "passed|failed|^FAILED|^ERROR" $O/${TAG}_tests.log | tail -8
for cfg in "1 1" "0 0" "1 0" "0 1"; do
  set -- $cfg
  DP_PDL=$1 DP_CARVEOUT=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > $O/${TAG}_bench_pdl$1_carve$2.log 2>&1
  echo "pdl=$1 carveout=$2 rc=$? $(tail -n 1 $O/${TAG}_bench_pdl$1_carve$2.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["value"], d["e2e"]["value"])' 2>&1 | tail -1)"
done
