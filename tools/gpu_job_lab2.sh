#!/bin/bash
O=gpurun_out; mkdir -p $O
for s in "qkv 128" "qkv 192" "fc1 192" "fc2 128" "proj 128"; do DP_GEMM_TRACE=1 timeout 120 python tools/gemm_trace.py $s 2>&1 | grep trace | head -40; done > $O/lab2_trace.log 2>&1
cat $O/lab2_trace.log
