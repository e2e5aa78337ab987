#!/bin/bash
# One gpurun call: tests, smoke, bench, then ncu launch list + one --set full capture of the top kernel.
# usage: tools/gpu_job_profile.sh <tag> [kernel-regex]
TAG=${1:-r1}
KRE=${2:-gemm_kmajor}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > $O/${TAG}_tests.log 2>&1; echo exit=$? >> $O/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo exit=$? >> $O/${TAG}_smoke.log
DP_BENCH_DUMP=$O/${TAG}_launches_events.csv timeout 600 python bench.py > $O/${TAG}_bench.log 2>&1; echo exit=$? >> $O/${TAG}_bench.log
timeout 300 python tools/prof_step.py --steps 3 > $O/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 260 --csv --log-file $O/${TAG}_ncu_launches.csv python tools/prof_step.py --steps 3 > $O/${TAG}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 60 -c 4 -o $O/${TAG}_prof python tools/prof_step.py --steps 2 > $O/${TAG}_ncu2.log 2>&1
tail -2 $O/${TAG}_tests.log $O/${TAG}_smoke.log $O/${TAG}_bench.log $O/${TAG}_ncu1.log $O/${TAG}_ncu2.log | cut -c1-600
