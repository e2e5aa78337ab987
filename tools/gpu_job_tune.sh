#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "${TESTK:-gemm}" > $O/tune_tests.log 2>&1; tail -n 6 $O/tune_tests.log | cut -c1-300
for d in ${DEBUGS:-0 8}; do DP_GEMM_DEBUG=$d timeout 300 python tools/gemm_tune.py ${SHAPES:-qkv,proj,fc1,fc2} ${BNS:-128,192,256} ${PAIRS:-2,1}; done > $O/tune2.log 2>&1
cat $O/tune2.log
