#!/bin/bash
# round-2 diagnostic: what bounds the K = 384 main loop (operand delivery vs MMA issue), and the cuBLAS comparator
O=gpurun_out; mkdir -p $O
timeout 300 python tools/cublas_ref.py > $O/lab1_cublas.log 2>&1
for d in 0 8 24 40 56; do DP_GEMM_DEBUG=$d timeout 300 python tools/gemm_tune.py qkv,fc1,fc2,fc2_B 128,192,256 2; done > $O/lab1_tune.log 2>&1
cat $O/lab1_cublas.log $O/lab1_tune.log
