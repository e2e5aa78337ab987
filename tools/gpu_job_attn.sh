#!/bin/bash
O=gpurun_out; mkdir -p $O
for st in 1 2 4; do DP_ATTN_BWD_STAGES=$st timeout 60 python tools/attn_bwd_tune.py; done 2>&1 | tee $O/attn_tune.log
