#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "a_stationary" > $O/lab3_tests.log 2>&1; tail -n 15 $O/lab3_tests.log | cut -c1-400
DP_GEMM_TRACE=2 timeout 300 python tools/gemm_tune.py qkv,fc1,fc2dg,proj 128 4,3,2 > $O/lab3_tune.log 2>&1; cat $O/lab3_tune.log
