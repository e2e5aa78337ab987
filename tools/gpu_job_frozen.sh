#!/bin/bash
# frozen-statistics (heads in eval mode) training step: kernel test, golden cases, additivity
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -s --timeout 300 -p no:cacheprovider \
  -k "frozen_statistics or trainfz or additive or batchnorm" > $O/frozen_tests.log 2>&1
grep -E "relL2|all trainable|hm max-rel|additivity|passed|failed|Error|assert" $O/frozen_tests.log | awk '{ if ($0 ~ /relL2/) { if ($3+0 > 0.06) print } else print }' | tail -n 60 | cut -c1-220
