#!/bin/bash
# One gpurun call without ncu: GPU tests, smoke, bench with the per-launch CUDA-event dump.
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/${TAG}_tests.log 2>&1; echo exit=$? >> $O/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo exit=$? >> $O/${TAG}_smoke.log
DP_BENCH_DUMP=$O/${TAG}_launches_events.csv timeout 600 python bench.py ${BENCH_ARGS} > $O/${TAG}_bench.log 2>&1; echo exit=$? >> $O/${TAG}_bench.log
grep -E "passed|failed|^FAILED|^ERROR" $O/${TAG}_tests.log | tail -15
tail -n 3 $O/${TAG}_smoke.log
tail -n 2 $O/${TAG}_bench.log | cut -c1-2500
