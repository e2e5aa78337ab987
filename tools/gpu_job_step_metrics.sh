#!/bin/bash
# DRAM traffic / tensor-pipe activity of every launch of the program (metric subset, few passes per kernel) -> CSV
TAG=${1:-r1g}; O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 python tools/prof_step.py --steps 2 > $O/${TAG}_plain.log 2>&1 || { tail -n 5 $O/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_step_metrics.csv python tools/prof_step.py --steps 2 > $O/${TAG}_ncu1.log 2>&1
ls -la $O; tail -n 2 $O/${TAG}_ncu1.log
