#!/usr/bin/env python
"""Micro-benchmark of the memory-bound kernels at the shapes of BASELINE configs[1] (batch 64, ViT-S): each op is
recorded on `nbuf` distinct buffer sets (L2-cold operands), the launches are captured in one CUDA graph and timed
with CUDA events.  Prints us per launch and the achieved GB/s of the op's algorithmic bytes.  Tuning aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dino_pose_b200.backend import CudaBackend  # noqa: E402

BF, F32 = torch.bfloat16, torch.float32
dev = torch.device("cuda:0")
B = 64


def r(*shape, dtype=F32):
    return torch.randn(*shape, device=dev).to(dtype)


def ops():
    M, D, R = B * 257, 384, 8
    P16, P47, P48 = B * 256, B * 47 * 47, B * 48 * 48

    def dw_fwd(be):
        x, w, b, out = r(B, 16, 16, 512, dtype=BF), r(512, 1, 3, 3), r(512), torch.empty(P16, 512, device=dev)
        be.dwconv3x3(x, w, b, None, out, NB=B, H=16, W=16, C=512)
        return P16 * 512 * (2 + 4)

    def dw_bwd(be):
        x, w, add, out = r(B, 16, 16, 512, dtype=BF), r(512, 1, 3, 3), r(B, 16, 16, 512, dtype=BF), torch.empty(P16, 512, device=dev, dtype=BF)
        be.dwconv3x3(x, w, None, add, out, NB=B, H=16, W=16, C=512, flip=True)
        return P16 * 512 * (2 + 2 + 2)

    def dw_wgrad(be):
        x, d, dw = r(B, 16, 16, 512, dtype=BF), r(B, 16, 16, 512, dtype=BF), torch.zeros(512, 1, 3, 3, device=dev)
        be.dwconv3x3_wgrad(x, d, dw, NB=B, H=16, W=16, C=512)
        return P16 * 512 * 4

    def lora_fwd(be):
        y, xin, xout, u = r(M, D), r(M, D), torch.empty(M, D, device=dev), torch.empty(M, R, device=dev)
        A, Bm, lam = r(D, R), r(R, D), r(D)
        seed = torch.zeros(1, device=dev, dtype=torch.int64)
        be.lora_fwd(y, A, Bm, lam, xin, xout, u, rows=M, D=D, R=R, scaling=2.0, p_drop=0.1, seed=seed)
        return M * D * 12

    def lora_bwd(be):
        g, y, u, gu = r(M, D), r(M, D), r(M, R), torch.empty(M, R, device=dev)
        Bm, lam, dA, dB = r(R, D), r(D), torch.zeros(D, R, device=dev), torch.zeros(R, D, device=dev)
        seed = torch.zeros(1, device=dev, dtype=torch.int64)
        be.lora_bwd(g, y, u, Bm, lam, dA, dB, gu, rows=M, D=D, R=R, scaling=2.0, p_drop=0.1, seed=seed)
        return M * D * 12

    def col2im_ups0(be):
        col, bias, big = r(P16, 16 * 128, dtype=BF), r(128), torch.empty(B, 47, 47, 128, device=dev)
        be.col2im(col, bias, big, NB=B, SH=16, SW=16, C=128, BH=47, BW=47, KH=4, KW=4, stride=3, pad=1)
        return P16 * 2048 * 2 + P47 * 128 * 4

    def mean_tok(be):
        f, o = r(B, 256, D, dtype=BF), torch.empty(B, D, device=dev)
        be.mean_tokens(f, o, B=B, N=256, D=D)
        return B * 256 * D * 2

    def mean_tok_bwd(be):
        f, o = r(B, 256, D, dtype=BF), r(B, D)
        be.mean_tokens_bwd(f, o, B=B, N=256, D=D)
        return B * 256 * D * 4

    def hm_grad(be):
        g, o = r(B, 24, 48, 48), torch.empty(P48, 32, device=dev, dtype=BF)
        be.hm_grad_to_nhwc(g, o, NB=B, K=24, Kp=32, OH=48, OW=48, up=1)
        return B * 24 * 2304 * 4 + P48 * 32 * 2

    def bn_apply_ups1(be):
        raw, out = r(P48, 128), torch.empty(P48, 128, device=dev, dtype=BF)
        sums = torch.zeros(2 * 128 * 9, device=dev, dtype=torch.float64)
        sums[:128] = 1.0
        sums[128:256] = float(P48)
        g, b_, rm, rv = r(128), r(128), torch.zeros(128, device=dev), torch.ones(128, device=dev)
        sc, sh, mu, inv = [torch.empty(128, device=dev) for _ in range(4)]
        be.bn_apply(raw, g, b_, None, None, out, P=P48, C=128, relu=True)
        return P48 * 128 * 6

    def bn_bwd_ups1(be):
        raw, d, draw = r(P48, 128), r(P48, 128, dtype=BF), torch.empty(P48, 128, device=dev, dtype=BF)
        sums = torch.zeros(2 * 128 * 9, device=dev, dtype=torch.float64)
        g, sc, sh, mu, inv = r(128), r(128), r(128), r(128), r(128).abs() + 0.5
        dg, db = torch.empty(128, device=dev), torch.empty(128, device=dev)
        be.bn_bwd_reduce(d, raw, None, sc, sh, mu, inv, sums, P=P48, C=128, relu=True, mode=0)
        be.bn_bwd_apply(d, raw, None, g, sc, sh, mu, inv, sums, draw, None, dg, db, P=P48, C=128, relu=True, mode=0)
        return P48 * 128 * (6 + 8)

    return {f.__name__: f for f in (dw_fwd, dw_bwd, dw_wgrad, lora_fwd, lora_bwd, col2im_ups0, mean_tok, mean_tok_bwd, hm_grad,
                                    bn_apply_ups1, bn_bwd_ups1)}


def bench(fn, nbuf=4, reps=12):
    be = CudaBackend()
    progs = []
    for _ in range(nbuf):
        prog = be.begin()
        nbytes = fn(be)
        progs.append(prog)
    for p in progs:
        p.run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                progs[i % nbuf].run()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, nbytes / us / 1e3


if __name__ == "__main__":
    table = ops()
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(table)
    for n in names:
        us, gbs = bench(table[n])
        print(f"{n:16s} {us:8.1f} us  {gbs:8.0f} GB/s (algorithmic bytes)", flush=True)
