#!/bin/bash
O=gpurun_out; mkdir -p $O
for s in "qkv 128 3" "fc1 128 3" "qkv 128 2"; do DP_GEMM_TRACE=1 timeout 120 python tools/gemm_trace.py $s 2>&1 | grep trace | tail -20; done > $O/lab2_trace.log 2>&1
cat $O/lab2_trace.log
