#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "gemm" > $O/tune_tests.log 2>&1; tail -n 3 $O/tune_tests.log
for d in ${DEBUGS:-0 8}; do DP_GEMM_DEBUG=$d timeout 300 python tools/gemm_tune.py ${SHAPES:-qkv,proj,fc1,fc2} ${BNS:-128,192,256}; done > $O/tune2.log 2>&1
cat $O/tune2.log
