#!/bin/bash
# A/B of the CTA-pair GEMM tiles for ViT-B / ViT-L (DP_PAIR_WIDE, DP_PAIR_QKV) on configs 2 and 3
O=gpurun_out; mkdir -p $O
for v in "0 0" "1 0" "1 1"; do
  set -- $v
  echo "DP_PAIR_WIDE=$1 DP_PAIR_QKV=$2"
  DP_PAIR_WIDE=$1 DP_PAIR_QKV=$2 timeout 600 python tools/run_configs.py cfg2 cfg3 2>&1 | tail -n 2 | cut -c1-140
done
timeout 600 python -m pytest tests/test_model_gpu.py -q --timeout 300 -p no:cacheprovider -k "base or large" 2>&1 | tail -n 2
