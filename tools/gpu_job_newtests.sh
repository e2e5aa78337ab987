#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_submodules_gpu.py tests/test_dropin_callers.py -m gpu -x -q --timeout 300 -p no:cacheprovider > $O/new_tests.log 2>&1; tail -n 40 $O/new_tests.log | cut -c1-400
