#!/bin/bash
O=gpurun_out; mkdir -p $O
DP_ATTN_FLASH=1 timeout 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 60 -p no:cacheprovider -k "test_attention and not bwd" > $O/attn_tests.log 2>&1; tail -n 12 $O/attn_tests.log | cut -c1-300
(B=64 T=1025 timeout 60 python tools/attn_tune.py; B=64 T=1025 DP_ATTN_FLASH_V=1 timeout 60 python tools/attn_tune.py; B=64 T=257 DP_ATTN_FLASH=1 timeout 60 python tools/attn_tune.py; B=64 T=257 timeout 60 python tools/attn_tune.py) 2>&1 | tee $O/attn_tune.log
