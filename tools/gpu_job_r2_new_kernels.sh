#!/bin/bash
# ncu --set full of the kernels added at the end of round 2 (one steady-state launch each)
O=gpurun_out; mkdir -p $O; T=r2y
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-comparator"
timeout 300 $BENCH > $O/${T}_plain.log 2>&1 || { tail -n 5 $O/${T}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lora_bwd_mma|bn_stats_kernel|patch_im2col_staged' -s 9 -c 4 -o /tmp/${T}_new $BENCH > $O/${T}_ncu.log 2>&1
ncu -i /tmp/${T}_new.ncu-rep --page raw --csv > $O/${T}_full_new_kernels_raw.csv 2>/dev/null
ls -la $O/${T}_*; tail -n 2 $O/${T}_ncu.log | cut -c1-200
