#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 60 -p no:cacheprovider -k "cta_pair" > $O/mix_tests.log 2>&1; tail -n 6 $O/mix_tests.log | cut -c1-300
timeout 200 python tools/gemm_tune.py qkv,fc1 128,192,256 1,2 2>&1 | tee $O/mix_tune.log
