#!/usr/bin/env python
"""cuBLAS (torch.matmul, bf16) on the backbone GEMM shapes, same timing method as tools/gemm_tune.py: the library
comparator for the hand-written tcgen05 kernel (tuning aid, not a reported number)."""
import sys

import torch

BF = torch.bfloat16
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16448
SHAPES = {"qkv": (M, 1152, 384), "proj": (M, 384, 384), "fc1": (M, 1536, 384), "fc2": (M, 384, 1536),
          "qkv_B": (M, 2304, 768), "fc1_B": (M, 3072, 768), "fc2_B": (M, 768, 3072), "sq8k": (8192, 8192, 8192)}


def bench(m, n, k, reps=20, nbuf=6):
    As = [torch.randn(m, k, device=dev).to(BF) for _ in range(nbuf)]
    Ws = [(torch.randn(n, k, device=dev) * 0.05).to(BF) for _ in range(nbuf)]
    outs = [torch.empty(m, n, device=dev, dtype=BF) for _ in range(nbuf)]
    for i in range(nbuf):
        torch.matmul(As[i], Ws[i].t(), out=outs[i])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for r in range(reps):
                torch.matmul(As[r % nbuf], Ws[r % nbuf].t(), out=outs[r % nbuf])
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, 2.0 * m * n * k / us / 1e6


if __name__ == "__main__":
    for nm, (m, n, k) in SHAPES.items():
        us, tf = bench(m, n, k, reps=20 if nm != "sq8k" else 5, nbuf=6 if nm != "sq8k" else 2)
        print(f"cublas {nm:6s} M={m} N={n} K={k}: {us:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
