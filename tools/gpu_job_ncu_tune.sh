#!/bin/bash
# usage: tools/gpu_job_ncu_tune.sh <tag> <shapes> <bn>
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 300 python tools/gemm_tune.py $2 $3 > $O/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_fwd -s 2 -c ${NCAP:-2} -o $O/${TAG}_prof python tools/gemm_tune.py $2 $3 > $O/${TAG}_ncu.log 2>&1
cat $O/${TAG}_plain.log; tail -n 3 $O/${TAG}_ncu.log
