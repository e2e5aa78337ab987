#!/usr/bin/env python
"""Micro-benchmark of dp_wgrad_bf16 on the head layers' shapes (batch 64, 224x224), graph-captured."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dino_pose_b200.backend import CudaBackend
BF = torch.bfloat16
dev = torch.device("cuda:0")
B = 64
# name: (kind, NB, H, W, Cm(dout ch), Cn(in ch), k, pad)   conv:  A = dout [NB,H,W,Cm], B = in [NB,H,W,Cn]
SHAPES = {
    "ups1": ("conv", B, 48, 48, 128, 128, 4, 1),   # convT_s1 as conv: A = x4 [47x47] ... approximated by 48x48
    "pred0": ("conv", B, 48, 48, 64, 128, 3, 1),
    "fr0": ("conv", B, 16, 16, 512, 384, 3, 1),
    "fr4": ("conv", B, 16, 16, 256, 512, 3, 1),
    "skip": ("plain", B * 256, 512, 512),
    "up2": ("plain", B * 64, 256, 2048),
}


def bench(name, splits, bn, reps=10):
    sh = SHAPES[name]
    be = CudaBackend()
    prog = be.begin()
    ws = torch.empty(16 << 20, device=dev) if os.environ.get('WS', '1') == '1' else None
    if sh[0] == "conv":
        _, nb, h, w, cm, cn, k, pad = sh
        A = torch.randn(nb, h, w, cm, device=dev).to(BF)
        Bt = torch.randn(nb, h, w, cn, device=dev).to(BF)
        out = torch.zeros(cm, cn, k, k, device=dev)
        be.wgrad(A, Bt, out, Mc=cm, Nc=cn, so_m=cn * k * k, so_n=k * k, so_t=1, conv=dict(KH=k, KW=k, pad=pad), splits=splits,
                 block_n=bn, workspace=ws)
        flops = 2.0 * nb * h * w * cm * cn * k * k
    else:
        _, P, cm, cn = sh
        A = torch.randn(P, cm, device=dev).to(BF)
        Bt = torch.randn(P, cn, device=dev).to(BF)
        out = torch.zeros(cm, cn, device=dev)
        be.wgrad(A, Bt, out, Mc=cm, Nc=cn, so_m=cn, so_n=1, P=P, splits=splits, block_n=bn, workspace=ws)
        flops = 2.0 * P * cm * cn
    prog.run(); torch.cuda.synchronize()
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
    return us, flops / us / 1e6


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
    dbg = os.environ.get("DP_WGRAD_DEBUG", "0")
    for nm in names:
        for bn in (128,):
            for splits in (0, 4, 9, 18):
                try:
                    us, tf = bench(nm, splits, bn)
                    print(f"debug={dbg} {nm:6s} bn={bn:3d} splits={splits:3d} {us:8.1f} us {tf:7.1f} TFLOP/s", flush=True)
                except Exception as ex:
                    print(f"debug={dbg} {nm:6s} bn={bn} splits={splits}: {str(ex)[:90]}", flush=True)
