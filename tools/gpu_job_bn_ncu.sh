#!/bin/bash
# ncu --set full (with source) of one step's BatchNorm kernels: where the ~12 us fixed cost per launch goes
O=gpurun_out; mkdir -p $O; T=r2x
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-comparator"
timeout 300 $BENCH > $O/${T}_plain.log 2>&1 || { tail -n 5 $O/${T}_plain.log; exit 1; }
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'bn_bwd_reduce|bn_bwd_apply|bn_stats_kernel|bn_finalize_apply' -s 180 -c 44 -o $O/${T}_bn $BENCH > $O/${T}_ncu.log 2>&1
ls -la $O/${T}_bn.ncu-rep; tail -n 2 $O/${T}_ncu.log | cut -c1-200
