#!/usr/bin/env python
"""DP_GEMM_TRACE=1 python tools/gemm_trace.py qkv 128 : prints the per-tile timeline of CTA 0 (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemm_tune import SHAPES, BF, dev
from dino_pose_b200.backend import CudaBackend
name, bn = sys.argv[1], int(sys.argv[2])
pair = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m, n, k, epi = SHAPES[name]
be = CudaBackend(); prog = be.begin()
A = torch.randn(m, k, device=dev).to(BF); W = (torch.randn(n, k, device=dev) * 0.05).to(BF); bias = torch.randn(n, device=dev)
if epi == "res_f32":
    out = torch.empty(m, n, device=dev); res = torch.randn(m, n, device=dev); ls = torch.ones(n, device=dev)
    be.gemm(A, W, out, M=m, N=n, K=k, bias=bias, ls=ls, residual=res, out_dtype="f32", block_n=bn, cta_pair=pair)
else:
    out = torch.empty(m, n, device=dev, dtype=BF)
    be.gemm(A, W, out, M=m, N=n, K=k, bias=bias, act="gelu" if epi == "gelu_bf16" else "none", block_n=bn, cta_pair=pair)
for _ in range(3):
    prog.run()
torch.cuda.synchronize()
