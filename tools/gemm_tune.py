#!/usr/bin/env python
"""Micro-benchmark of dp_gemm_bf16 on the backbone / head shapes of BASELINE configs[1] (CUDA events, L2 flushed
between repetitions by cycling through enough distinct buffers).  Tuning aid, not a reported number."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dino_pose_b200.backend import CudaBackend  # noqa: E402

BF = torch.bfloat16
dev = torch.device("cuda:0")
M = 16448
SHAPES = {
    # name: (M, N, K, epilogue)
    "qkv": (M, 1152, 384, "bias_bf16"),
    "proj": (M, 384, 384, "res_f32"),
    "fc1": (M, 1536, 384, "gelu_bf16"),
    "fc2": (M, 384, 1536, "res_f32"),
    "fc2dg": (M, 1536, 384, "auxin_bf16"),     # last block's fc2 input gradient: x gelu'(pre-activation)
    "proj_ln": (M, 384, 384, "res_f32_ln"),     # row-owning kernel, LayerNorm fused into the epilogue
    "fc2_ln": (M, 384, 1536, "res_f32_ln"),
    "fc1_B": (M, 3072, 768, "gelu_bf16"),
    "fc2_B": (M, 768, 3072, "res_f32"),
}


def bench(name, bn, reps=20, nbuf=6, pair=2):
    m, n, k, epi = SHAPES[name]
    be = CudaBackend()
    progs = []
    for i in range(nbuf):
        A = torch.randn(m, k, device=dev).to(BF)
        W = (torch.randn(n, k, device=dev) * 0.05).to(BF)
        bias = torch.randn(n, device=dev)
        prog = be.begin()
        if epi == "bias_bf16":
            out = torch.empty(m, n, device=dev, dtype=BF)
            be.gemm(A, W, out, M=m, N=n, K=k, bias=bias, block_n=bn, cta_pair=pair)
        elif epi == "auxin_bf16":
            out = torch.empty(m, n, device=dev, dtype=BF)
            aux = torch.randn(m, n, device=dev).to(BF)
            be.gemm(A, W, out, M=m, N=n, K=k, aux_in=aux, ld_aux=n, block_n=bn, cta_pair=pair)
        elif epi == "gelu_bf16":
            out = torch.empty(m, n, device=dev, dtype=BF)
            be.gemm(A, W, out, M=m, N=n, K=k, bias=bias, act="gelu", block_n=bn, cta_pair=pair)
        elif epi == "res_f32_ln":
            out = torch.empty(m, n, device=dev)
            res = torch.randn(m, n, device=dev)
            ls = torch.ones(n, device=dev)
            xn = torch.empty(m, n, device=dev, dtype=BF)
            be.gemm(A, W, out, M=m, N=n, K=k, bias=bias, ls=ls, residual=res, out_dtype="f32",
                    ln=dict(gamma=ls, beta=bias, out=xn, eps=1e-6))
        else:
            out = torch.empty(m, n, device=dev)
            res = torch.randn(m, n, device=dev)
            ls = torch.ones(n, device=dev)
            be.gemm(A, W, out, M=m, N=n, K=k, bias=bias, ls=ls, residual=res, out_dtype="f32", block_n=bn, cta_pair=pair)
        progs.append(prog)
    for p in progs:
        p.run()
    torch.cuda.synchronize()
    # one CUDA graph of `reps` launches: no CPU launch cost between kernels
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for r in range(reps):
                progs[r % nbuf].run()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    tf = 2.0 * m * n * k / us / 1e6
    if os.environ.get("DP_GEMM_TRACE") == "2":
        # SM clock during the replay: cycles / ns of CTA 0 of the last launch; cycles from its first MMA to its end
        import ctypes
        buf = (ctypes.c_longlong * 4096)()
        be.lib.dp_debug_read_trace(buf, 4096)
        cyc, ns = buf[4002] - buf[4000], buf[4003] - buf[4001]
        print(f"      probe: CTA0 {cyc} cycles in {ns} ns = {cyc / max(ns, 1):.3f} GHz", flush=True)
    return us, tf


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
    bns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 192, 256]
    dbg = os.environ.get("DP_GEMM_DEBUG", "0")
    pairs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 1]   # 2 = single CTA, 1 = CTA pair
    for nm in names:
        for bn in bns:
            for pr in pairs:
                try:
                    us, tf = bench(nm, bn, pair=pr)
                except Exception as ex:   # no variant compiled for this combination
                    print(f"debug={dbg} {nm:6s} bn={bn:3d} {'pair' if pr == 1 else 'single'}: {str(ex)[:80]}", flush=True)
                    continue
                print(f"debug={dbg} {nm:6s} bn={bn:3d} {'pair  ' if pr == 1 else 'single'} {us:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
