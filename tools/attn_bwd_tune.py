#!/usr/bin/env python
"""Micro-benchmark of dp_attention_bwd (default B=64, T=257, 6 heads), graph-captured.  DP_ATTN_BWD_STAGES: bit mask of the three launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dino_pose_b200.backend import CudaBackend
BF = torch.bfloat16
dev = torch.device("cuda:0")
B, T, H = int(os.environ.get("B", 64)), int(os.environ.get("T", 257)), int(os.environ.get("H", 6))
D = H * 64
qkv = (torch.randn(B * T, 3 * D, device=dev) * 0.5).to(BF)
ctx = (torch.randn(B * T, D, device=dev) * 0.3).to(BF)
dctx = (torch.randn(B * T, D, device=dev) * 0.3).to(BF)
dqkv = torch.zeros(B * T, 3 * D, device=dev, dtype=BF)
stats = torch.zeros(2 * B * H * T, device=dev)
be = CudaBackend(); prog = be.begin()
be.attention_bwd(qkv, ctx, dctx, dqkv, stats, B=B, T=T, heads=H, scale=0.125)
prog.run(); torch.cuda.synchronize()
reps = 10
g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            prog.run()
torch.cuda.current_stream().wait_stream(side)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"stages={os.environ.get('DP_ATTN_BWD_STAGES','7')} B={B} T={T} H={H}: {us:7.1f} us  {10.0*B*H*T*T*64/us/1e6:6.1f} TFLOP/s (5 GEMMs)")
