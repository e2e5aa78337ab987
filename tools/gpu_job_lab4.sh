#!/bin/bash
O=gpurun_out; mkdir -p $O
DP_GEMM_TRACE=2 timeout 300 python tools/gemm_tune.py qkv,fc1,fc2 128 3,2 > $O/lab4_tune.log 2>&1; cat $O/lab4_tune.log
