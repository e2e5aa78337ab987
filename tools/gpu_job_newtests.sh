#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest ${TESTS:-tests/test_parity_bench_shape_gpu.py} -m gpu -x -q -s --timeout 600 -p no:cacheprovider > $O/new_tests.log 2>&1; grep -v "^  backbone\|^  pose_heads" $O/new_tests.log | tail -n 30 | cut -c1-300
