#!/usr/bin/env python
"""Micro-benchmark of dp_attention_fwd (B=64, T=257, 6 heads = ViT-S/14 at 224x224), graph-captured, with a check."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dino_pose_b200.backend import CudaBackend
BF = torch.bfloat16
dev = torch.device("cuda:0")
B, T, H = int(os.environ.get("B", 64)), int(os.environ.get("T", 257)), int(os.environ.get("H", 6))
D = H * 64
qkv = (torch.randn(B * T, 3 * D, device=dev) * 0.5).to(BF)
ctx = torch.zeros(B * T, D, device=dev, dtype=BF)
be = CudaBackend(); prog = be.begin()
be.attention_fwd(qkv, ctx, B=B, T=T, heads=H, scale=0.125)
prog.run(); torch.cuda.synchronize()
q, k, v = [t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
ref = torch.softmax(q @ k.transpose(2, 3) * 0.125, -1) @ v
ref = ref.transpose(1, 2).reshape(B * T, D)
err = ((ctx.float() - ref).abs().max() / ref.abs().max()).item()
reps = 20
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
print(f"flash={os.environ.get('DP_ATTN_FLASH','0')} B={B} T={T} H={H}: {us:7.1f} us  {4.0*B*H*T*T*64/us/1e6:6.1f} TFLOP/s  max-rel err {err:.2e}")
