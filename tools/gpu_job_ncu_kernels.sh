#!/bin/bash
# usage: tools/gpu_job_ncu_kernels.sh <tag> "<regex1> <regex2> ..."   -- one ncu --set full capture (2 launches) per regex
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 300 python tools/prof_step.py --steps 2 > $O/${TAG}_plain.log 2>&1 || { tail -n 5 $O/${TAG}_plain.log; exit 1; }
for k in $2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-30} -c ${NCAP:-2} -o $O/${TAG}_$k python tools/prof_step.py --steps 2 > $O/${TAG}_$k.log 2>&1
  tail -n 1 $O/${TAG}_$k.log
done
