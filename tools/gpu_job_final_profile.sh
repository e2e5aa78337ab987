#!/bin/bash
# Round-end evidence (all reports are converted to CSV ON THE BOX: gpurun_out/ may carry at most 64 MiB back):
#   1. launch list of the whole program (gpu__time_duration)
#   2. DRAM traffic / tensor-pipe activity of EVERY launch of one step (metric subset, 3 passes per kernel)
#   3. --set full of a few representative kernels (GEMM variants, wgrad, attention, LayerNorm, BatchNorm backward)
# usage: tools/gpu_job_final_profile.sh <tag>
TAG=${1:-r1f}; O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 python tools/prof_step.py --steps 2 > $O/${TAG}_plain.log 2>&1 || { tail -n 5 $O/${TAG}_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_ncu_launches.csv python tools/prof_step.py --steps 2 > $O/${TAG}_ncu0.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:'dp::' -s ${STEP_SKIP:-540} -c ${STEP_COUNT:-560} --csv --log-file $O/${TAG}_step_metrics.csv python tools/prof_step.py --steps 2 > $O/${TAG}_ncu1.log 2>&1
SEL='gemm_fwd|gemm_wgrad|attention_tc|layernorm_fwd|bn_bwd_reduce|bn_bwd_apply_kernel'
# (a) first backbone layer of a replayed step: LN, qkv, attention, proj, LN, fc1, fc2, LN   (158 selected launches per step)
timeout 600 ncu --set full --clock-control none -k regex:"$SEL" -s 316 -c 8 -o /tmp/${TAG}_fulla python tools/prof_step.py --steps 2 > $O/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_fulla.ncu-rep --page raw --csv > $O/${TAG}_full_backbone_raw.csv 2>/dev/null
# (b) end of the head forward + start of the backward: convs, weight gradients, input gradients, BatchNorm backward
timeout 900 ncu --set full --clock-control none -k regex:"$SEL" -s 411 -c 26 -o /tmp/${TAG}_fullb python tools/prof_step.py --steps 2 > $O/${TAG}_ncu3.log 2>&1
ncu -i /tmp/${TAG}_fullb.ncu-rep --page raw --csv > $O/${TAG}_full_heads_raw.csv 2>/dev/null
ls -la $O | head -20
tail -n 2 $O/${TAG}_ncu0.log $O/${TAG}_ncu1.log $O/${TAG}_ncu2.log $O/${TAG}_ncu3.log | cut -c1-200
