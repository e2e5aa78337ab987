"""Per-kernel parity on a B200: every C-ABI entry point against the plain torch fp32 op it replaces
(or the numpy oracle for decode).  Called through the ctypes binding, i.e. through the C ABI."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def dev():
    return torch.device("cuda:0")


def backend():
    from dino_pose_b200.backend import CudaBackend
    return CudaBackend()


def run(fn):
    b = backend()
    prog = b.begin()
    fn(b)
    prog.run()
    torch.cuda.synchronize()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rnd(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev()).to(dtype)


# ---------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (257, 384, 384, 128), (1000, 1152, 384, 128),
                                      (300, 1536, 384, 256), (515, 384, 1536, 128), (200, 64, 128, 64),
                                      (130, 24, 64, 32), (512, 384, 640, 128), (16448, 384, 384, 128),
                                      (1000, 1152, 384, 192), (16448, 1536, 384, 192), (16448, 1152, 384, 256),
                                      (777, 768, 768, 192), (300, 104, 192, 128), (4096, 512, 512, 256)])
def test_gemm_plain(M, N, K, bn):
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias = rnd(N, seed=2)
    out = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out, M=M, N=N, K=K, bias=bias, block_n=bn))
    ref = A.float() @ W.float().t() + bias
    assert rel(out.float(), ref) < 1e-2


@pytest.mark.parametrize("M,N,K,bn", [(256, 192, 64, 192), (257, 384, 384, 192), (1000, 1152, 384, 192),
                                      (16448, 1536, 384, 192), (16448, 1152, 384, 256), (515, 512, 1536, 256),
                                      (640, 256, 128, 128), (129, 384, 384, 128), (16448, 384, 384, 192)])
def test_gemm_cta_pair(M, N, K, bn):
    """cta_group::2 kernel: 256-row tiles over two CTAs, incl. an odd number of 128-row blocks (phantom half)."""
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias = rnd(N, seed=2)
    out = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out, M=M, N=N, K=K, bias=bias, block_n=bn, cta_pair=1))
    ref = A.float() @ W.float().t() + bias
    assert rel(out.float(), ref) < 1e-2
    # fp32 out + LayerScale + in-place residual, and GELU (+ saved derivative) through the pair kernel
    ls = rnd(N, seed=4)
    x = rnd(M, N, seed=5)
    x0 = x.clone()
    run(lambda b: b.gemm(A, W, x, M=M, N=N, K=K, bias=bias, ls=ls, residual=x, out_dtype="f32", block_n=bn, cta_pair=1))
    assert rel(x, x0 + ref * ls) < 2e-3
    if bn != 128:
        out2 = torch.zeros(M, N, device=dev(), dtype=BF)
        aux = torch.zeros(M, N, device=dev(), dtype=BF)
        run(lambda b: b.gemm(A, W, out2, M=M, N=N, K=K, bias=bias, act="gelu", aux_out=aux, ld_aux=N, block_n=bn, cta_pair=1))
        xr = ref.clone().requires_grad_(True)
        F.gelu(xr).sum().backward()
        assert rel(aux.float(), xr.grad) < 1e-2       # side output of a GELU epilogue = gelu'(v), the backward's multiplier
        assert rel(out2.float(), F.gelu(ref)) < 1e-2


@pytest.mark.parametrize("kind", [3, 4])
@pytest.mark.parametrize("M,N,K", [(16448, 1152, 384), (16448, 1536, 384), (300, 256, 64), (2049, 384, 128),
                                   (1000, 1152, 512), (4100, 640, 384)])
def test_gemm_a_stationary(M, N, K, kind):
    """A-stationary kernel (gemm_astat.cuh): the 128 x K row block of A is copied into tensor memory (tcgen05.cp) once
    per row block and the MMAs run in the TS form; contiguous tile ranges per CTA, ragged last row block / column tile.
    kind 4: clusters of two CTAs, weight k-blocks fetched half / half and multicast (odd row-block counts: phantom half)."""
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias = rnd(N, seed=2)
    ref = A.float() @ W.float().t() + bias
    out = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out, M=M, N=N, K=K, bias=bias, cta_pair=kind))              # TMA-store epilogue
    assert rel(out.float(), ref) < 1e-2
    out2 = torch.zeros(M, N, device=dev(), dtype=BF)
    aux = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out2, M=M, N=N, K=K, bias=bias, act="gelu", aux_out=aux, ld_aux=N, cta_pair=kind))
    xr = ref.clone().requires_grad_(True)
    F.gelu(xr).sum().backward()
    assert rel(aux.float(), xr.grad) < 1e-2
    assert rel(out2.float(), F.gelu(ref)) < 1e-2
    out3 = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out3, M=M, N=N, K=K, bias=bias, act="gelu", cta_pair=kind))  # GELU + TMA store
    assert rel(out3.float(), F.gelu(ref)) < 1e-2
    mult = rnd(M, N, seed=9, dtype=BF)
    out4 = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out4, M=M, N=N, K=K, aux_in=mult, ld_aux=N, cta_pair=kind))
    assert rel(out4.float(), (A.float() @ W.float().t()) * mult.float()) < 1e-2
    ls = rnd(N, seed=4)
    x = rnd(M, N, seed=5)
    x0 = x.clone()
    run(lambda b: b.gemm(A, W, x, M=M, N=N, K=K, bias=bias, ls=ls, residual=x, out_dtype="f32", cta_pair=kind))
    assert rel(x, x0 + ref * ls) < 2e-3


@pytest.mark.parametrize("M,N,K", [(16448, 384, 384), (16448, 384, 1536), (257, 384, 384), (1000, 128, 128), (515, 256, 512),
                                   (129, 384, 64)])
def test_gemm_row_owning_with_fused_layernorm(M, N, K):
    """gemm_rowln.cu: one CTA owns 128 complete rows; v = residual + ls * (A W^T + bias) is written in place into the fp32
    residual stream and LayerNorm(v) leaves the same epilogue as bf16 (HF modeling_dinov2.py:250 + 373-379, 327 + 382-384)."""
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias, ls = rnd(N, seed=2), rnd(N, seed=4)
    gamma, beta = rnd(N, seed=6).abs() + 0.5, rnd(N, seed=7)
    x = rnd(M, N, seed=5) * 2.0 + 0.3
    x0 = x.clone()
    xn = torch.full((M, N), 7.0, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, x, M=M, N=N, K=K, bias=bias, ls=ls, residual=x, out_dtype="f32",
                         ln=dict(gamma=gamma, beta=beta, out=xn, eps=1e-6)))
    ref = x0 + (A.float() @ W.float().t() + bias) * ls
    assert rel(x, ref) < 2e-3
    ref_ln = F.layer_norm(x, (N,), gamma, beta, 1e-6)          # LayerNorm of what the kernel itself wrote
    assert rel(xn.float(), ref_ln) < 6e-3                       # bf16 rounding of the output
    # separate output tensor (the last block in training keeps x_mid / x_last apart from the stream)
    out2 = torch.zeros(M, N, device=dev())
    xn2 = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out2, M=M, N=N, K=K, bias=bias, ls=ls, residual=x0, out_dtype="f32",
                         ln=dict(gamma=gamma, beta=beta, out=xn2, eps=1e-6)))
    assert rel(out2, ref) < 2e-3 and rel(xn2.float(), F.layer_norm(out2, (N,), gamma, beta, 1e-6)) < 6e-3


@pytest.mark.parametrize("M,N,K,p_drop", [(16448, 384, 384, 0.0), (1000, 384, 384, 0.1), (257, 128, 128, 0.0), (515, 128, 512, 0.1)])
def test_gemm_with_fused_lora_adapter(M, N, K, p_drop):
    """gemm_rowln.cu MODE 1: out = residual + ls * (y + s * dropout(y A B)), y = ctx W^T + bias, rank 8 (reference
    model/lora.py:26-28,53-59 + HF:373-376).  Checked against the torch expression and, with dropout, against the separate
    dp_lora_fwd kernel on the SAME seed (the masks must be identical: the backward regenerates them from the seed)."""
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias, ls = rnd(N, seed=2), rnd(N, seed=4)
    la, lb = rnd(N, 8, scale=0.2, seed=6), rnd(8, N, scale=0.2, seed=7)
    x_in = rnd(M, N, seed=5)
    seed = torch.tensor([12345], device=dev(), dtype=torch.int64)
    out = torch.zeros(M, N, device=dev())
    y = torch.zeros(M, N, device=dev())
    u = torch.zeros(M, 8, device=dev())
    run(lambda b: b.gemm(A, W, out, M=M, N=N, K=K, bias=bias, ls=ls, residual=x_in, out_dtype="f32",
                         lora=dict(A=la, B=lb, scaling=2.0, p_drop=p_drop, seed=seed, y_out=y, u_out=u)))
    y_ref = A.float() @ W.float().t() + bias
    assert rel(y, y_ref) < 2e-3 and rel(u, y_ref @ la) < 2e-3
    # the separate adapter kernel on the y the fused kernel saved: same arithmetic, same mask
    sep = torch.zeros(M, N, device=dev())
    u2 = torch.zeros(M, 8, device=dev())
    run(lambda b: b.lora_fwd(y, la, lb, ls, x_in, sep, u2, rows=M, D=N, R=8, scaling=2.0, p_drop=p_drop, seed=seed))
    assert rel(out, sep) < 1e-5, rel(out, sep)
    assert rel(u, u2) < 1e-5
    if p_drop == 0.0:
        assert rel(out, x_in + (y_ref + (y_ref @ la @ lb) * 2.0) * ls) < 2e-3


def test_gemm_epilogues():
    M, N, K = 700, 384, 256
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias, scale, ls = rnd(N, seed=2), rnd(N, seed=3).abs() + 0.5, rnd(N, seed=4)
    res = rnd(M, N, seed=5)
    base = (A.float() @ W.float().t()) * scale + bias
    # fp32 out + layerscale + residual
    out = torch.zeros(M, N, device=dev())
    run(lambda b: b.gemm(A, W, out, M=M, N=N, K=K, bias=bias, scale=scale, ls=ls, residual=res, out_dtype="f32"))
    assert rel(out, res + base * ls) < 2e-3
    # gelu + aux_out
    out2 = torch.zeros(M, N, device=dev(), dtype=BF)
    aux = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out2, M=M, N=N, K=K, bias=bias, scale=scale, act="gelu", aux_out=aux, ld_aux=N))
    xb = base.clone().requires_grad_(True)
    F.gelu(xb).sum().backward()
    assert rel(aux.float(), xb.grad) < 1e-2          # with a GELU the side output is gelu'(v), the backward's multiplier
    assert rel(out2.float(), F.gelu(base)) < 1e-2
    aux0 = torch.zeros(M, N, device=dev(), dtype=BF)
    out20 = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out20, M=M, N=N, K=K, bias=bias, scale=scale, aux_out=aux0, ld_aux=N))
    assert rel(aux0.float(), base) < 1e-2            # without an activation it is v itself
    # relu, bf16 residual
    resb = res.to(BF)
    out3 = torch.zeros(M, N, device=dev(), dtype=BF)
    run(lambda b: b.gemm(A, W, out3, M=M, N=N, K=K, bias=bias, act="relu", residual=resb))
    assert rel(out3.float(), F.relu(A.float() @ W.float().t() + bias) + resb.float()) < 1e-2
    # fused GELU backward multiplier
    out4 = torch.zeros(M, N, device=dev(), dtype=BF)
    mult = rnd(M, N, seed=9, dtype=BF)
    run(lambda b: b.gemm(A, W, out4, M=M, N=N, K=K, aux_in=mult, ld_aux=N))
    assert rel(out4.float(), (A.float() @ W.float().t()) * mult.float()) < 1e-2


@pytest.mark.parametrize("M,N,bn", [(5000, 128, 128), (40000, 512, 256), (3000, 384, 192), (129, 64, 64)])
def test_gemm_fused_bn_statistics(M, N, bn):
    """Per-column sum / sum-of-squares of the pre-activation value (train-mode BatchNorm statistics fused into the
    producing GEMM's epilogue), fp64 accumulators, same contract as dp_bn_stats."""
    K = 128
    A = rnd(M, K, dtype=BF)
    W = rnd(N, K, scale=0.05, seed=1, dtype=BF)
    bias = rnd(N, seed=2)
    out = torch.zeros(M, N, device=dev())
    sums = torch.zeros(2 * N, device=dev(), dtype=torch.float64)
    run(lambda b: b.gemm(A, W, out, M=M, N=N, K=K, bias=bias, out_dtype="f32", block_n=bn, stats=sums, stats_c=N))
    ref = (A.float() @ W.float().t() + bias).double()
    assert rel(out, ref) < 2e-3
    assert rel(sums[:N], ref.sum(0)) < 1e-4
    assert rel(sums[N:], (ref * ref).sum(0)) < 1e-4


def test_gemm_fused_bn_statistics_shuffle2x2():
    NB, H, Wd, Cin, Cout = 3, 8, 8, 128, 256
    x = rnd(NB * H * Wd, Cin, dtype=BF)
    Wm = rnd(4 * Cout, Cin, scale=0.05, seed=1, dtype=BF)
    bias4 = rnd(Cout, seed=2).repeat(4).contiguous()
    out = torch.zeros(NB * 2 * H * 2 * Wd, Cout, device=dev())
    sums = torch.zeros(2 * Cout, device=dev(), dtype=torch.float64)
    run(lambda b: b.gemm(x, Wm, out, M=NB * H * Wd, N=4 * Cout, K=Cin, bias=bias4, out_dtype="f32", row_map="shuffle2x2",
                         map_a=Cout, OH=H, OW=Wd, NB=NB, stats=sums, stats_c=Cout))
    assert rel(sums[:Cout], out.double().sum(0)) < 1e-4
    assert rel(sums[Cout:], (out.double() ** 2).sum(0)) < 1e-4


def test_gemm_patch_rowmap_and_nchw():
    B, Np, T, D, K = 3, 16, 17, 128, 640
    A = rnd(B * Np, K, dtype=BF)
    W = rnd(D, K, scale=0.05, seed=1, dtype=BF)
    pos = rnd(T, D, seed=2)
    x = torch.zeros(B * T, D, device=dev())
    run(lambda b: b.gemm(A, W, x, M=B * Np, N=D, K=K, out_dtype="f32", residual=pos, row_map="patch_tokens",
                         map_a=Np, map_b=T))
    ref = (A.float() @ W.float().t()).view(B, Np, D) + pos[1:]
    assert rel(x.view(B, T, D)[:, 1:], ref) < 2e-3
    assert x.view(B, T, D)[:, 0].abs().max().item() == 0
    # NCHW fp32 output with 24 valid columns
    NB, OH, OW, Cin, Kc = 2, 8, 16, 64, 24
    A2 = rnd(NB * OH * OW, Cin, dtype=BF)
    W2 = rnd(Kc, Cin, scale=0.1, seed=3, dtype=BF)
    bias = rnd(Kc, seed=4)
    hm = torch.zeros(NB, Kc, OH, OW, device=dev())
    run(lambda b: b.gemm(A2, W2, hm, M=NB * OH * OW, N=Kc, K=Cin, bias=bias, out_dtype="f32", row_map="nchw",
                         n_valid=Kc, map_a=Kc, OH=OH, OW=OW, NB=NB, block_n=32))
    ref2 = (A2.float() @ W2.float().t() + bias).view(NB, OH, OW, Kc).permute(0, 3, 1, 2)
    assert rel(hm, ref2) < 2e-3


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,pad,OH,OW", [(2, 16, 16, 128, 128, 3, 1, 16, 16),
                                                         (3, 48, 48, 64, 64, 3, 1, 48, 48),
                                                         (2, 47, 47, 128, 128, 4, 2, 48, 48),
                                                         (2, 48, 48, 128, 128, 4, 1, 47, 47),
                                                         (9, 4, 4, 128, 128, 3, 1, 4, 4),
                                                         (3, 8, 8, 64, 256, 3, 1, 8, 8),
                                                         (1, 32, 32, 384, 512, 3, 1, 32, 32)])
def test_gemm_implicit_conv(NB, H, W, Cin, Cout, k, pad, OH, OW):
    x = rnd(NB, H, W, Cin, dtype=BF)
    w = rnd(Cout, Cin, k, k, scale=0.05, seed=1)
    Wm = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous().to(BF)
    bias = rnd(Cout, seed=2)
    out = torch.zeros(NB * OH * OW, Cout, device=dev(), dtype=BF)
    run(lambda b: b.gemm(x, Wm, out, M=NB * OH * OW, N=Cout, K=k * k * Cin, bias=bias,
                         conv=dict(KH=k, KW=k, pad=pad, OH=OH, OW=OW)))
    xn = x.float().permute(0, 3, 1, 2)
    xp = F.pad(xn, (pad, OW + k - 1 - pad - W, pad, OH + k - 1 - pad - H))
    ref = F.conv2d(xp, Wm.float().view(Cout, k, k, Cin).permute(0, 3, 1, 2), bias)
    assert ref.shape[-2:] == (OH, OW)
    assert rel(out.float().view(NB, OH, OW, Cout).permute(0, 3, 1, 2), ref) < 1e-2


@pytest.mark.parametrize("NB,H,Cin,Cout,bn", [(8, 16, 128, 512, 256), (8, 16, 512, 256, 256), (3, 48, 128, 128, 128), (5, 16, 128, 384, 192)])
def test_gemm_implicit_conv_cta_pair(NB, H, Cin, Cout, bn):
    """CTA-pair (cta_group::2, 256-pixel tiles) implicit 3x3 convolution: fp32 output + fused BatchNorm statistics (the heads'
    forward), and bf16 output + bf16 residual (their input gradients); odd tile counts leave a phantom half."""
    k, pad = 3, 1
    x = rnd(NB, H, H, Cin, dtype=BF)
    w = rnd(Cout, Cin, k, k, scale=0.05, seed=1)
    Wm = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous().to(BF)
    bias = rnd(Cout, seed=2)
    P = NB * H * H
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), Wm.float().view(Cout, k, k, Cin).permute(0, 3, 1, 2), bias, padding=pad)
    ref = ref.permute(0, 2, 3, 1).reshape(P, Cout)
    out = torch.zeros(P, Cout, device=dev())
    sums = torch.zeros(2 * Cout, device=dev(), dtype=torch.float64)
    run(lambda b: b.gemm(x, Wm, out, M=P, N=Cout, K=k * k * Cin, bias=bias, out_dtype="f32", stats=sums, stats_c=Cout,
                         conv=dict(KH=k, KW=k, pad=pad, OH=H, OW=H), block_n=bn, cta_pair=1))
    assert rel(out, ref) < 2e-3
    assert rel(sums[:Cout], out.double().sum(0)) < 1e-4 and rel(sums[Cout:], (out.double() ** 2).sum(0)) < 1e-4
    res = rnd(P, Cout, seed=7, dtype=BF)
    out2 = torch.zeros(P, Cout, device=dev(), dtype=BF)
    run(lambda b: b.gemm(x, Wm, out2, M=P, N=Cout, K=k * k * Cin, residual=res, conv=dict(KH=k, KW=k, pad=pad, OH=H, OW=H),
                         block_n=bn, cta_pair=1))
    assert rel(out2.float(), ref - bias + res.float()) < 1e-2


def test_gemm_implicit_conv_token_view():
    """A = patch tokens inside the [B,T,D] final-LayerNorm output (CLS skipped by pointer offset)."""
    B, g, D, Cout = 2, 16, 128, 128
    T = g * g + 1
    tok = rnd(B, T, D, dtype=BF)
    x = tok[:, 1:, :].unflatten(1, (g, g))  # [B,g,g,D] strided view
    w = rnd(Cout, D, 3, 3, scale=0.05, seed=1)
    Wm = w.permute(0, 2, 3, 1).reshape(Cout, 9 * D).contiguous().to(BF)
    out = torch.zeros(B * g * g, Cout, device=dev(), dtype=BF)
    run(lambda b: b.gemm(x, Wm, out, M=B * g * g, N=Cout, K=9 * D, conv=dict(KH=3, KW=3, pad=1, OH=g, OW=g)))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), Wm.float().view(Cout, 3, 3, D).permute(0, 3, 1, 2), None, 1, 1)
    assert rel(out.float().view(B, g, g, Cout).permute(0, 3, 1, 2), ref) < 1e-2


def test_gemm_shuffle2x2():
    NB, H, W, Cin, Cout = 2, 4, 4, 128, 256
    x = rnd(NB * H * W, Cin, dtype=BF)
    wt = rnd(Cin, Cout, 2, 2, scale=0.05, seed=1)
    Wm = wt.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous().to(BF)
    out = torch.zeros(NB * 2 * H * 2 * W, Cout, device=dev(), dtype=BF)
    bias4 = rnd(Cout, seed=2).repeat(4).contiguous()
    run(lambda b: b.gemm(x, Wm, out, M=NB * H * W, N=4 * Cout, K=Cin, bias=bias4, row_map="shuffle2x2", map_a=Cout,
                         OH=H, OW=W, NB=NB))
    xn = x.float().view(NB, H, W, Cin).permute(0, 3, 1, 2)
    ref = F.conv_transpose2d(xn, Wm.float().view(2, 2, Cout, Cin).permute(3, 2, 0, 1), bias4[:Cout], 2)
    assert rel(out.float().view(NB, 2 * H, 2 * W, Cout).permute(0, 3, 1, 2), ref) < 1e-2


# ---------------------------------------------------------------------------------- weight gradients
@pytest.mark.parametrize("P,Mc,Nc,bn", [(256, 128, 128, 128), (1000, 256, 192, 64), (4096, 24, 64, 64),
                                        (777, 512, 384, 128)])
def test_wgrad_plain(P, Mc, Nc, bn):
    A = rnd(P, Mc, dtype=BF)
    Bm = rnd(P, Nc, seed=1, dtype=BF)
    lda = A.stride(0)
    out = torch.zeros(Mc, Nc, device=dev())
    run(lambda b: b.wgrad(A, Bm, out, Mc=Mc, Nc=Nc, so_m=Nc, so_n=1, P=P, block_n=bn))
    ref = A.float().t() @ Bm.float()
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("case", ["plain", "conv3", "convT4", "decomposed"])
def test_wgrad_workspace_path_is_exact_and_deterministic(case):
    """Atomic-free split-K: partial tiles in a workspace + reduce kernel; `out` needs no zeroing, two runs are bit-identical."""
    ws = torch.empty(8 << 20, device=dev())
    outs = []
    for rep in range(2):
        if case == "plain":
            P, Mc, Nc = 16384, 512, 192
            A, Bm = rnd(P, Mc, dtype=BF), rnd(P, Nc, seed=1, dtype=BF)
            out = torch.full((Mc, Nc), 7.0, device=dev())
            run(lambda b: b.wgrad(A, Bm, out, Mc=Mc, Nc=Nc, so_m=Nc, so_n=1, P=P, workspace=ws))
            ref = A.float().t() @ Bm.float()
        elif case == "conv3":
            NB, H, W, Cin, Cout, k, pad = 6, 16, 16, 128, 256, 3, 1
            x, dy = rnd(NB, H, W, Cin, dtype=BF), rnd(NB, H, W, Cout, seed=1, dtype=BF)
            out = torch.full((Cout, Cin, k, k), 7.0, device=dev())
            run(lambda b: b.wgrad(dy, x, out, Mc=Cout, Nc=Cin, so_m=Cin * k * k, so_n=k * k, so_t=1,
                                  conv=dict(KH=k, KW=k, pad=pad), workspace=ws))
            w = torch.zeros(Cout, Cin, k, k, device=dev(), requires_grad=True)
            F.conv2d(x.float().permute(0, 3, 1, 2), w, None, 1, pad).backward(dy.float().permute(0, 3, 1, 2))
            ref = w.grad
        elif case == "convT4":
            NB, Cin, Cout = 3, 128, 128
            x, dy = rnd(NB, 47, 47, Cin, dtype=BF), rnd(NB, 48, 48, Cout, seed=1, dtype=BF)
            out = torch.full((Cin, Cout, 4, 4), 7.0, device=dev())
            run(lambda b: b.wgrad(x, dy, out, Mc=Cin, Nc=Cout, so_m=Cout * 16, so_n=16, so_t=1, conv=dict(KH=4, KW=4, pad=1),
                                  workspace=ws))
            w = torch.zeros(Cin, Cout, 4, 4, device=dev(), requires_grad=True)
            F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w, None, 1, 1).backward(dy.float().permute(0, 3, 1, 2))
            ref = w.grad
        else:
            P, Cout, Cin, taps = 2048, 128, 64, 9
            A, col = rnd(P, Cout, dtype=BF), rnd(P, taps * Cin, seed=1, dtype=BF)
            out = torch.full((Cout, Cin, taps), 7.0, device=dev())
            run(lambda b: b.wgrad(A, col, out, Mc=Cout, Nc=taps * Cin, so_m=Cin * taps, so_n=taps, so_no=1, n_inner=Cin, P=P,
                                  workspace=ws))
            ref = (A.float().t() @ col.float()).view(Cout, taps, Cin).permute(0, 2, 1)
        assert rel(out, ref) < 2e-3
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1])


def test_wgrad_plain_decomposed_offsets():
    """n = tap*Cin + ci scattered into a [Cout, Cin, 9] gradient (strided-conv layers)."""
    P, Cout, Cin, taps = 512, 128, 64, 9
    A = rnd(P, Cout, dtype=BF)
    col = rnd(P, taps * Cin, seed=1, dtype=BF)
    out = torch.zeros(Cout, Cin, taps, device=dev())
    run(lambda b: b.wgrad(A, col, out, Mc=Cout, Nc=taps * Cin, so_m=Cin * taps, so_n=taps, so_no=1, n_inner=Cin, P=P))
    ref = (A.float().t() @ col.float()).view(Cout, taps, Cin).permute(0, 2, 1)
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,pad", [(2, 16, 16, 128, 128, 3, 1), (3, 48, 48, 128, 64, 3, 1),
                                                   (5, 4, 4, 128, 128, 3, 1), (2, 8, 8, 64, 256, 3, 1)])
def test_wgrad_conv(NB, H, W, Cin, Cout, k, pad):
    x = rnd(NB, H, W, Cin, dtype=BF)
    dy = rnd(NB, H, W, Cout, seed=1, dtype=BF)
    out = torch.zeros(Cout, Cin, k, k, device=dev())
    run(lambda b: b.wgrad(dy, x, out, Mc=Cout, Nc=Cin, so_m=Cin * k * k, so_n=k * k, so_t=1,
                          conv=dict(KH=k, KW=k, pad=pad)))
    xn = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w = torch.zeros(Cout, Cin, k, k, device=dev(), requires_grad=True)
    F.conv2d(xn, w, None, 1, pad).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(out, w.grad) < 2e-3


def test_wgrad_convT_s1_as_conv():
    """ConvTranspose2d(k4,s1,p1) 47->48: dW[ci,co,ky,kx] = sum in[iy,ix,ci] * dOut[iy-1+ky, ix-1+kx, co]."""
    NB, Cin, Cout = 2, 128, 128
    x = rnd(NB, 47, 47, Cin, dtype=BF)
    dy = rnd(NB, 48, 48, Cout, seed=1, dtype=BF)
    out = torch.zeros(Cin, Cout, 4, 4, device=dev())
    run(lambda b: b.wgrad(x, dy, out, Mc=Cin, Nc=Cout, so_m=Cout * 16, so_n=16, so_t=1, conv=dict(KH=4, KW=4, pad=1)))
    w = torch.zeros(Cin, Cout, 4, 4, device=dev(), requires_grad=True)
    F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w, None, 1, 1).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(out, w.grad) < 2e-3


# ---------------------------------------------------------------------------------- LayerNorm / LoRA / attention
@pytest.mark.parametrize("D", [128, 384, 768, 1024])
def test_layernorm_fwd_bwd(D):
    B, T = 3, 17
    x = rnd(B * T, D, scale=2.0) + 0.3
    g, bt = rnd(D, seed=1) * 0.1 + 1, rnd(D, seed=2) * 0.1
    y = torch.zeros(B * T, D, device=dev(), dtype=BF)
    y32 = torch.zeros(B * T, D, device=dev())
    run(lambda b: b.layernorm_fwd(x, g, bt, y, y32, rows=B * T, D=D))
    ref = F.layer_norm(x, (D,), g, bt, 1e-6)
    assert rel(y32, ref) < 1e-5
    assert rel(y.float(), ref) < 1e-2
    # drop_cls
    yd = torch.zeros(B * (T - 1), D, device=dev(), dtype=BF)
    run(lambda b: b.layernorm_fwd(x, g, bt, yd, None, rows=B * T, D=D, T=T, drop_cls=True))
    assert rel(yd.float(), ref.view(B, T, D)[:, 1:].reshape(-1, D)) < 1e-2
    # backward
    dy = rnd(B * T, D, seed=3)
    add = rnd(B * T, D, seed=4)
    ls = rnd(D, seed=5)
    dx = torch.zeros(B * T, D, device=dev())
    dxs = torch.zeros(B * T, D, device=dev(), dtype=BF)
    run(lambda b: b.layernorm_bwd(dy, x, g, add, dx, rows=B * T, D=D, ls=ls, dx_scaled=dxs))
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), g, bt, 1e-6).backward(dy)
    assert rel(dx, xr.grad + add) < 1e-4
    assert rel(dxs.float(), (xr.grad + add) * ls) < 1e-2
    # backward with dropped CLS rows and bf16 dy
    dyd = rnd(B * (T - 1), D, seed=6, dtype=BF)
    dx2 = torch.zeros(B * T, D, device=dev())
    run(lambda b: b.layernorm_bwd(dyd, x, g, None, dx2, rows=B * T, D=D, T=T, drop_cls=True))
    xr = x.clone().requires_grad_(True)
    full = torch.zeros(B, T, D, device=dev())
    full[:, 1:] = dyd.float().view(B, T - 1, D)
    F.layer_norm(xr, (D,), g, bt, 1e-6).backward(full.view(B * T, D))
    assert rel(dx2, xr.grad) < 1e-4


@pytest.mark.parametrize("rows,D", [(1000, 384), (77, 128), (1000, 768)])
def test_lora_fwd_bwd(rows, D):
    """D = 384 / 128: the backward is the one-pass warp-MMA kernel (bf16 operands: 5e-3); D = 768: the two fp32 kernels."""
    R, s = 8, 2.0
    tol_bwd = 5e-3 if D <= 384 else 1e-4
    y, xin = rnd(rows, D), rnd(rows, D, seed=1)
    A, Bm, lam = rnd(D, R, scale=0.2, seed=2), rnd(R, D, scale=0.2, seed=3), rnd(D, seed=4).abs() + 0.5
    xout = torch.zeros(rows, D, device=dev())
    u = torch.zeros(rows, R, device=dev())
    run(lambda b: b.lora_fwd(y, A, Bm, lam, xin, xout, u, rows=rows, D=D, R=R, scaling=s, p_drop=0.0, seed=None))
    Ar, Br = A.clone().requires_grad_(True), Bm.clone().requires_grad_(True)
    ref = xin + (y + (y @ Ar @ Br) * s) * lam
    assert rel(xout, ref) < 1e-5
    assert rel(u, y @ A) < 1e-5
    g = rnd(rows, D, seed=5)
    ref.backward(g)
    dA, dB = torch.zeros_like(A), torch.zeros_like(Bm)
    gu = torch.zeros(rows, R, device=dev())
    run(lambda b: b.lora_bwd(g, y, u, Bm, lam, dA, dB, gu, rows=rows, D=D, R=R, scaling=s, p_drop=0.0, seed=None))
    print("lora bwd", D, rel(dA, Ar.grad), rel(dB, Br.grad))
    assert rel(dA, Ar.grad) < tol_bwd
    assert rel(dB, Br.grad) < tol_bwd
    # the gradients ACCUMULATE into dA / dB (the engine zeroes the flat gradient buffer once per step)
    run(lambda b: b.lora_bwd(g, y, u, Bm, lam, dA, dB, gu, rows=rows, D=D, R=R, scaling=s, p_drop=0.0, seed=None))
    assert rel(dA, 2 * Ar.grad) < tol_bwd and rel(dB, 2 * Br.grad) < tol_bwd
    # dropout: statistically ~10% dropped, kept values scaled by 1/(1-p)
    seed = torch.tensor([1234], device=dev(), dtype=torch.int64)
    xo2 = torch.zeros(rows, D, device=dev())
    run(lambda b: b.lora_fwd(y, A, Bm, lam, xin, xo2, None, rows=rows, D=D, R=R, scaling=s, p_drop=0.1, seed=seed))
    v = ((xo2 - xin) / lam - y) / s
    full = y @ A @ Bm
    dropped = (v.abs() < 1e-6) & (full.abs() > 1e-3)
    frac = dropped.float().mean().item()
    assert 0.08 < frac < 0.12
    kept = ~dropped & (full.abs() > 1e-2)
    assert rel(v[kept], full[kept] / 0.9) < 1e-2
    # the backward regenerates the SAME mask from (seed, element index): gradients of the masked forward
    mask = torch.where(full.abs() > 1e-3, (~dropped).float(), torch.ones_like(full)) / 0.9   # unknown where full ~ 0: keep
    Ar2, Br2 = A.clone().requires_grad_(True), Bm.clone().requires_grad_(True)
    (xin + (y + (y @ Ar2 @ Br2) * mask * s) * lam).backward(g)
    dA2, dB2 = torch.zeros_like(A), torch.zeros_like(Bm)
    run(lambda b: b.lora_bwd(g, y, u, Bm, lam, dA2, dB2, gu, rows=rows, D=D, R=R, scaling=s, p_drop=0.1, seed=seed))
    # elements with |full| < 1e-3 have an unknown mask bit (2 % of them at most matter): loose bound, a wrong index
    # convention gives O(1)
    assert rel(dA2, Ar2.grad) < 3e-2 and rel(dB2, Br2.grad) < 3e-2


@pytest.mark.parametrize("B,T,heads", [(2, 257, 6), (1, 1025, 6), (3, 65, 2), (2, 257, 12), (2, 1025, 2), (3, 300, 2),
                                       (2, 401, 3), (1, 528, 1), (1, 2304, 1), (2, 273, 1)])
def test_attention(B, T, heads):
    """T <= 272: resident-K/V tcgen05 kernels; longer: the flash-style tcgen05 kernel (ragged last key / query tiles)."""
    D = heads * 64
    qkv = rnd(B * T, 3 * D, dtype=BF)
    ctx = torch.zeros(B * T, D, device=dev(), dtype=BF)
    run(lambda b: b.attention_fwd(qkv, ctx, B=B, T=T, heads=heads, scale=0.125))
    q, k, v = [t.view(B, T, heads, 64).transpose(1, 2) for t in qkv.float().view(B, T, 3 * D).split(D, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    assert rel(ctx.float(), ref) < 1.5e-2


@pytest.mark.parametrize("B,T,heads", [(2, 257, 6), (1, 1025, 2), (3, 65, 2), (2, 129, 12), (1, 64, 1)])
def test_attention_bwd(B, T, heads):
    """dq | dk | dv against autograd of softmax(q k^T * scale) v in fp32 on the same bf16 inputs (HF:203-234)."""
    D = heads * 64
    qkv = rnd(B * T, 3 * D, dtype=BF)
    dctx = rnd(B * T, D, scale=0.5, seed=5, dtype=BF)
    q, k, v = [t.reshape(B, T, heads, 64).transpose(1, 2).detach().clone().requires_grad_(True)
               for t in qkv.float().view(B, T, 3 * D).split(D, dim=-1)]
    o = (torch.softmax(q @ k.transpose(2, 3) * 0.125, -1) @ v).transpose(1, 2).reshape(B * T, D)
    o.backward(dctx.float())
    ref = torch.cat([t.grad.transpose(1, 2).reshape(B * T, D) for t in (q, k, v)], -1)
    ctx = o.detach().to(BF)
    dqkv = torch.full((B * T, 3 * D), 7.0, device=dev(), dtype=BF)
    stats = torch.zeros(2 * B * heads * T, device=dev())
    run(lambda b: b.attention_bwd(qkv, ctx, dctx, dqkv, stats, B=B, T=T, heads=heads, scale=0.125))
    for j, name in enumerate("qkv"):
        got, want = dqkv[:, j * D:(j + 1) * D].float(), ref[:, j * D:(j + 1) * D]
        assert rel(got, want) < 2e-2, (name, rel(got, want))
        l2 = ((got - want).norm() / want.norm()).item()
        assert l2 < 1e-2, (name, l2)
    # row statistics written for the dK / dV kernel: log2-domain log-sum-exp of the scaled scores
    lse = torch.logsumexp(q.detach() @ k.detach().transpose(2, 3) * 0.125, -1) * 1.4426950408889634
    assert rel(stats[:B * heads * T].view(B, heads, T), lse) < 1e-4
    # deterministic: a second run gives the same bits
    again = torch.zeros_like(dqkv)
    run(lambda b: b.attention_bwd(qkv, ctx, dctx, again, stats, B=B, T=T, heads=heads, scale=0.125))
    assert torch.equal(again, dqkv)


@pytest.mark.parametrize("D,rows,dt", [(128, 700, BF), (384, 16448, BF), (768, 515, torch.float32), (1024, 1030, BF)])
def test_layernorm_param_grads_and_colsum_prod(D, rows, dt):
    x = rnd(rows, D, scale=2.0) + 0.3
    dy = rnd(rows, D, seed=4, dtype=dt)
    dg = torch.zeros(D, device=dev())
    db = torch.zeros(D, device=dev())
    run(lambda b: b.layernorm_bwd_params(dy, x, dg, db, rows=rows, D=D, eps=1e-6))
    xr = x.double()
    xhat = (xr - xr.mean(-1, keepdim=True)) * torch.rsqrt(xr.var(-1, unbiased=False, keepdim=True) + 1e-6)
    assert rel(dg, (dy.double() * xhat).sum(0)) < 2e-4
    assert rel(db, dy.double().sum(0)) < 2e-4
    a = rnd(rows, D, seed=6, dtype=BF)
    out = torch.zeros(D, device=dev())
    run(lambda b: b.colsum_prod(x, a, out, P=rows, C=D))
    assert rel(out, (x.double() * a.double()).sum(0)) < 2e-4


@pytest.mark.parametrize("NB,HW,K", [(2, 2304, 24), (3, 700, 17), (1, 256, 32), (5, 2304, 24)])
def test_pred1x1_fwd_bwd(NB, HW, K):
    """prediction.3 (Conv2d(64, K, 1) + bias, pose_heads.py:335-340) and its backward on CUDA cores vs torch fp32."""
    P, C = NB * HW, 64
    a = rnd(P, C, dtype=BF)
    w = rnd(K, C, 1, 1, scale=0.2, seed=1)
    bias = rnd(K, seed=2)
    out = torch.full((NB, K, HW), 7.0, device=dev())
    run(lambda b: b.pred1x1_fwd(a, w, bias, out, P=P, HW=HW, C=C, K=K))
    # fp64 reference (torch's fp32 conv on the GPU may run in TF32)
    x = a.double().view(NB, HW, C).permute(0, 2, 1).reshape(NB, C, HW, 1).requires_grad_(True)
    wr = w.double().requires_grad_(True)
    br = bias.double().requires_grad_(True)
    ref = F.conv2d(x, wr, br)
    assert rel(out, ref.view(NB, K, HW)) < 1e-5
    g = rnd(NB, K, HW, seed=3)
    ref.backward(g.double().view(NB, K, HW, 1))
    d = torch.zeros(P, C, device=dev(), dtype=BF)
    dW = torch.zeros(K, C, 1, 1, device=dev())
    db = torch.zeros(K, device=dev())
    run(lambda b: b.pred1x1_bwd(g, a, w, d, dW, db, P=P, HW=HW, C=C, K=K))
    assert rel(d.float(), x.grad.view(NB, C, HW).permute(0, 2, 1).reshape(P, C)) < 1e-2     # bf16 output
    assert rel(dW, wr.grad) < 1e-4
    assert rel(db, br.grad) < 1e-4


def test_patch_im2col_and_cls():
    B, H, W, Kp, D = 2, 28, 42, 640, 128
    for (B, H, W) in ((2, 28, 42), (3, 224, 224), (1, 448, 448)):
        n = (H // 14) * (W // 14)
        px = rnd(B, 3, H, W)
        out = torch.full((B * n, Kp), 7.0, device=dev(), dtype=BF)
        run(lambda b: b.patch_im2col(px, out, B=B, H=H, W=W, Kp=Kp))
        ref = F.unfold(px, 14, stride=14).transpose(1, 2).reshape(B * n, 588)
        assert torch.equal(out[:, :588], ref.to(BF))
        assert out[:, 588:].abs().max().item() == 0
    B = 2
    T = 7
    x = torch.zeros(B * T, D, device=dev())
    row = rnd(D, seed=3)
    run(lambda b: b.fill_cls(x, row, B=B, T=T, D=D))
    assert torch.equal(x.view(B, T, D)[:, 0], row.expand(B, D))


# ---------------------------------------------------------------------------------- decode
def test_decode_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode.npz"))
    for case in ("random", "edges", "nonsquare", "rect_map"):
        hm = torch.from_numpy(g[case + ".heatmaps"]).to(dev())
        Bn, K, H, W = hm.shape
        tw, th = [int(v) for v in g[case + ".target"]]
        idx = torch.zeros(Bn * K, 2, device=dev(), dtype=torch.int32)
        xy = torch.zeros(Bn * K, 2, device=dev(), dtype=torch.float64)
        conf = torch.zeros(Bn * K, device=dev())
        run(lambda b: b.decode(hm, idx, xy, conf, maps=Bn * K, H=H, W=W, target_w=tw, target_h=th))
        assert np.array_equal(idx.cpu().numpy().reshape(Bn, K, 2).astype(np.int64), g[case + ".idx"]), case
        got = xy.cpu().numpy().reshape(Bn, K, 2)
        ref = g[case + ".xy"]
        same = (got.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (case, got[~same], ref[~same])


def test_decode_large_batch_vs_oracle():
    from oracle import decode_oracle
    rng = np.random.default_rng(3)
    hm_np = (rng.standard_normal((64, 24, 48, 48)) * 0.07 + 0.06).astype(np.float32)
    hm = torch.from_numpy(hm_np).to(dev())
    idx = torch.zeros(64 * 24, 2, device=dev(), dtype=torch.int32)
    xy = torch.zeros(64 * 24, 2, device=dev(), dtype=torch.float64)
    run(lambda b: b.decode(hm, idx, xy, None, maps=64 * 24, H=48, W=48, target_w=224, target_h=224))
    ridx, rxy = decode_oracle.decode_batch(hm_np)
    assert np.array_equal(idx.cpu().numpy().reshape(64, 24, 2), ridx)
    assert np.array_equal(xy.cpu().numpy().reshape(64, 24, 2).view(np.uint64), rxy.view(np.uint64))


# ---------------------------------------------------------------------------------- heads kernels
def test_im2col_col2im():
    NB, IH, IW, C, k, s, p = 2, 16, 16, 64, 3, 2, 1
    OH = OW = 8
    x = rnd(NB, IH, IW, C, dtype=BF)
    col = torch.zeros(NB * OH * OW, k * k * C, device=dev(), dtype=BF)
    run(lambda b: b.im2col(x, col, NB=NB, IH=IH, IW=IW, C=C, OH=OH, OW=OW, KH=k, KW=k, stride=s, pad=p))
    ref = F.unfold(x.float().permute(0, 3, 1, 2), k, padding=p, stride=s)  # [NB, C*k*k, L]
    ref = ref.view(NB, C, k * k, OH * OW).permute(0, 3, 2, 1).reshape(NB * OH * OW, k * k * C)
    assert torch.equal(col.float(), ref)
    # col2im == conv_transpose forward (k4 s3 p1, 16->47)
    Cin, Cout = 64, 64
    xin = rnd(NB * 256, Cin, dtype=BF)
    wt = rnd(Cin, Cout, 4, 4, scale=0.05, seed=1)
    colT = (xin.float() @ wt.permute(2, 3, 1, 0).reshape(16 * Cout, Cin).t()).to(BF).contiguous()
    bias = rnd(Cout, seed=2)
    big = torch.zeros(NB, 47, 47, Cout, device=dev(), dtype=BF)
    run(lambda b: b.col2im(colT, bias, big, NB=NB, SH=16, SW=16, C=Cout, BH=47, BW=47, KH=4, KW=4, stride=3, pad=1))
    refT = F.conv_transpose2d(xin.float().view(NB, 16, 16, Cin).permute(0, 3, 1, 2), wt, bias, 3, 1)
    assert rel(big.float().permute(0, 3, 1, 2), refT) < 1.5e-2


def test_dwconv_and_wgrad():
    NB, H, W, C = 2, 16, 16, 128
    x = rnd(NB, H, W, C, dtype=BF)
    w, bias = rnd(C, 1, 3, 3, scale=0.3, seed=1), rnd(C, seed=2)
    out = torch.zeros(NB, H, W, C, device=dev(), dtype=BF)
    run(lambda b: b.dwconv3x3(x, w, bias, None, out, NB=NB, H=H, W=W, C=C))
    wr = w.clone().requires_grad_(True)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(xr, wr, bias, 1, 1, 1, C)
    assert rel(out.float().permute(0, 3, 1, 2), ref) < 1e-2
    dy = rnd(NB, H, W, C, seed=3, dtype=BF)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.zeros(NB, H, W, C, device=dev(), dtype=BF)
    add = rnd(NB, H, W, C, seed=4, dtype=BF)
    run(lambda b: b.dwconv3x3(dy, w, None, add, dx, NB=NB, H=H, W=W, C=C, flip=True))
    assert rel(dx.float().permute(0, 3, 1, 2), xr.grad + add.float().permute(0, 3, 1, 2)) < 1e-2
    dw = torch.zeros(C, 1, 3, 3, device=dev())
    run(lambda b: b.dwconv3x3_wgrad(x, dy, dw, NB=NB, H=H, W=W, C=C))
    assert rel(dw, wr.grad) < 1e-3


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode,rawdt,C", [(0, BF, 128), (1, BF, 128), (0, torch.float32, 512), (1, torch.float32, 64)])
def test_batchnorm_train_fwd_bwd(mode, rawdt, C, fused):
    P = 3 * 16 * 16
    raw = (rnd(P, C) * 1.5 + 0.2).to(rawdt)
    gamma, beta = rnd(C, seed=1) * 0.1 + 1, rnd(C, seed=2) * 0.1
    rm, rv = rnd(C, seed=3) * 0.1, rnd(C, seed=4).abs() + 0.5
    add1 = rnd(P, C, seed=5, dtype=BF)
    add2 = rnd(P, C, seed=6, dtype=BF) if mode == 0 else None
    sums = torch.zeros(2 * C * 9, device=dev(), dtype=torch.float64)     # (DP_BN_BWD_REPLICAS + 1) x 2C (backward)
    scale, shift, mean, invstd = [torch.zeros(C, device=dev()) for _ in range(4)]
    rm2, rv2 = rm.clone(), rv.clone()
    out = torch.zeros(P, C, device=dev(), dtype=BF)

    def fwd(b):
        b.bn_stats(raw, sums, P=P, C=C)
        if fused:   # one launch: per-block finalize, last block re-zeroes the sums
            b.bn_finalize_apply(raw, sums, gamma, beta, rm2, rv2, scale, shift, mean, invstd, add1, add2, out, P=P, C=C,
                                relu=True, mode=mode)
        else:
            b.bn_finalize(sums, gamma, beta, rm2, rv2, scale, shift, mean, invstd, C=C, count=P)
            b.bn_apply(raw, scale, shift, add1, add2, out, P=P, C=C, relu=True, mode=mode)
    run(fwd)
    rr = raw.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a1 = add1.float().requires_grad_(True)
    rm3, rv3 = rm.clone(), rv.clone()
    y = F.batch_norm(rr.view(1, P, C).permute(0, 2, 1), rm3, rv3, gr, br, True, 0.1, 1e-5).permute(0, 2, 1).reshape(P, C)
    ref = F.relu(y + a1) if mode == 1 else F.relu(y) + a1 + add2.float()
    assert rel(out.float(), ref) < 1.5e-2
    assert rel(rm2, rm3) < 1e-4 and rel(rv2, rv3) < 1e-4
    assert sums.abs().max().item() == 0
    dout = rnd(P, C, seed=7, dtype=BF)
    ref.backward(dout.float())
    draw = torch.zeros(P, C, device=dev(), dtype=BF)
    dres = torch.zeros(P, C, device=dev(), dtype=BF) if mode == 1 else None
    dg, db = torch.zeros(C, device=dev()), torch.zeros(C, device=dev())

    def bwd(b):
        b.bn_bwd_reduce(dout, raw, add1 if mode == 1 else None, scale, shift, mean, invstd, sums, P=P, C=C, relu=True,
                        mode=mode)
        b.bn_bwd_apply(dout, raw, add1 if mode == 1 else None, gamma, scale, shift, mean, invstd, sums, draw, dres, dg,
                       db, P=P, C=C, relu=True, mode=mode)
    for _ in range(2):   # the second pass checks that the accumulators were re-zeroed by the first
        run(bwd)
        assert sums.abs().max().item() == 0
        assert rel(draw.float(), rr.grad) < 2e-2
        assert rel(dg, gr.grad) < 1e-2 and rel(db, br.grad) < 1e-2
        if mode == 1:
            assert rel(dres.float(), a1.grad) < 1e-2


@pytest.mark.parametrize("mode,rawdt,C", [(0, BF, 128), (1, torch.float32, 64), (0, torch.float32, 1024)])
def test_batchnorm_frozen_statistics_fwd_bwd(mode, rawdt, C):
    """nn.BatchNorm2d in eval mode inside a training step (model.pose_heads.eval()): dp_bn_fold_eval derives scale / shift
    AND the saved mean / invstd from the running statistics, dp_bn_bwd_apply(eval_mode=2) returns dy * scale and
    dgamma = sum dy * xhat.  C = 1024 takes the two-launch (coefficient kernel) form, the others the fused one."""
    P = 3 * 16 * 16
    raw = (rnd(P, C) * 1.5 + 0.2).to(rawdt)
    gamma, beta = rnd(C, seed=1) * 0.1 + 1, rnd(C, seed=2) * 0.1
    rm, rv = rnd(C, seed=3) * 0.1, rnd(C, seed=4).abs() + 0.5
    add1 = rnd(P, C, seed=5, dtype=BF)
    add2 = rnd(P, C, seed=6, dtype=BF) if mode == 0 else None
    sums = torch.zeros(2 * C * 9, device=dev(), dtype=torch.float64)
    scale, shift, mean, invstd = [torch.zeros(C, device=dev()) for _ in range(4)]
    rm2, rv2 = rm.clone(), rv.clone()
    out = torch.zeros(P, C, device=dev(), dtype=BF)

    def fwd(b):
        b.bn_fold_eval(gamma, beta, rm2, rv2, None, scale, shift, C=C, mean=mean, invstd=invstd)
        b.bn_apply(raw, scale, shift, add1, add2, out, P=P, C=C, relu=True, mode=mode)
    run(fwd)
    rr = raw.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a1 = add1.float().requires_grad_(True)
    y = F.batch_norm(rr.view(1, P, C).permute(0, 2, 1), rm, rv, gr, br, False, 0.1, 1e-5).permute(0, 2, 1).reshape(P, C)
    ref = F.relu(y + a1) if mode == 1 else F.relu(y) + a1 + add2.float()
    assert rel(out.float(), ref) < 1.5e-2
    assert torch.equal(rm2, rm) and torch.equal(rv2, rv)            # running statistics untouched
    assert torch.equal(mean, rm) and rel(invstd, torch.rsqrt(rv + 1e-5)) < 1e-6
    dout = rnd(P, C, seed=7, dtype=BF)
    ref.backward(dout.float())
    draw = torch.zeros(P, C, device=dev(), dtype=BF)
    dres = torch.zeros(P, C, device=dev(), dtype=BF) if mode == 1 else None
    dg, db = torch.zeros(C, device=dev()), torch.zeros(C, device=dev())

    def bwd(b):
        b.bn_bwd_reduce(dout, raw, add1 if mode == 1 else None, scale, shift, mean, invstd, sums, P=P, C=C, relu=True,
                        mode=mode)
        b.bn_bwd_apply(dout, raw, add1 if mode == 1 else None, gamma, scale, shift, mean, invstd, sums, draw, dres, dg,
                       db, P=P, C=C, relu=True, mode=mode, eval_mode=2)
    for _ in range(2):
        run(bwd)
        assert sums[:2 * C * 8].abs().max().item() == 0      # the accumulators (the ninth block is coefficient scratch)
        assert rel(draw.float(), rr.grad) < 1e-2
        assert rel(dg, gr.grad) < 1e-2 and rel(db, br.grad) < 1e-2
        if mode == 1:
            assert rel(dres.float(), a1.grad) < 1e-2


def test_small_ops():
    # avgpool2 == bilinear(align_corners=False) at scale 1/2
    x = rnd(2 * 24, 96, 96)
    out = torch.zeros(2 * 24, 48, 48, device=dev())
    run(lambda b: b.avgpool2(x, out, planes=48, OH=48, OW=48))
    ref = F.interpolate(x.view(2, 24, 96, 96), size=(48, 48), mode="bilinear", align_corners=False).view(48, 48, 48)
    assert rel(out, ref) < 1e-6
    # heat-map gradient layout conversion (+ adjoint of the 2x reduction)
    g = rnd(2, 24, 48, 48, seed=1)
    o1 = torch.zeros(2 * 48 * 48, 32, device=dev(), dtype=BF)
    run(lambda b: b.hm_grad_to_nhwc(g, o1, NB=2, K=24, Kp=32, OH=48, OW=48, up=1))
    assert torch.equal(o1[:, :24], g.permute(0, 2, 3, 1).reshape(-1, 24).to(BF)) and o1[:, 24:].abs().max() == 0
    o2 = torch.zeros(2 * 96 * 96, 32, device=dev(), dtype=BF)
    run(lambda b: b.hm_grad_to_nhwc(g, o2, NB=2, K=24, Kp=32, OH=96, OW=96, up=2))
    xr = torch.zeros(2, 24, 96, 96, device=dev(), requires_grad=True)
    F.interpolate(xr, size=(48, 48), mode="bilinear", align_corners=False).backward(g)
    assert rel(o2[:, :24].float(), xr.grad.permute(0, 2, 3, 1).reshape(-1, 24)) < 1e-2
    # mean over tokens and its adjoint
    feat = rnd(3, 256, 384, dtype=BF)
    m = torch.zeros(3, 384, device=dev())
    run(lambda b: b.mean_tokens(feat, m, B=3, N=256, D=384))
    assert rel(m, feat.float().mean(1)) < 1e-5
    dfeat = rnd(3, 256, 384, seed=2, dtype=BF)
    want = (dfeat.float() + (m / 256)[:, None, :]).to(BF)
    run(lambda b: b.mean_tokens_bwd(dfeat, m, B=3, N=256, D=384))
    assert rel(dfeat.float(), want.float()) < 1e-2
    # colsum
    xx = rnd(5000, 24 + 8, seed=3, dtype=BF)
    cs = torch.zeros(24, device=dev())
    run(lambda b: b.colsum(xx, cs, P=5000, C=24, ld=32))
    assert rel(cs, xx.float()[:, :24].sum(0)) < 1e-3


def test_sgemm_small():
    M, K, N = 64, 384, 1024
    x, w, bias = rnd(M, K), rnd(N, K, scale=0.05, seed=1), rnd(N, seed=2)
    y = torch.zeros(M, N, device=dev())
    run(lambda b: b.sgemm_small(x, K, 1, w, 1, K, y, N, M=M, N=N, K=K, bias=bias, relu=True))
    ref = F.relu(x @ w.t() + bias)
    assert rel(y, ref) < 1e-5
    # backward forms: dX = (dY*mask) W ; dW = (dY*mask)^T X
    dy = rnd(M, N, seed=3)
    dpre = torch.zeros(M, N, device=dev())
    run(lambda b: b.relu_mask(dy, y, dpre, n=M * N, keep_scale=1.25))
    assert rel(dpre, 1.25 * dy * (ref > 0)) < 1e-6
    run(lambda b: b.relu_mask(dy, y, dpre, n=M * N))
    dx = torch.zeros(M, K, device=dev())
    run(lambda b: b.sgemm_small(dpre, N, 1, w, K, 1, dx, K, M=M, N=K, K=N))
    assert rel(dx, dpre @ w) < 1e-5
    dw = torch.zeros(N, K, device=dev())
    run(lambda b: b.sgemm_small(dpre, 1, N, x, K, 1, dw, K, M=N, N=K, K=M))
    assert rel(dw, dpre.t() @ x) < 1e-5


# ---------------------------------------------------------------------------------- training-step kernels
def test_pose_loss_and_seeds_vs_torch():
    """dp_pose_loss vs the reference losses (train.py:89-120) + DynamicLossWeighting (:17-69) through autograd,
    over two steps (the second exercises the EMA state kept on the device)."""
    from oracle import pose_oracle
    B, K, HW = 5, 24, 48 * 48
    sums = torch.zeros(2, device=dev(), dtype=torch.float64)
    state = torch.tensor([0.0, 0.0, 0.0, 0.1], device=dev())
    out, scales = torch.zeros(3, device=dev()), torch.zeros(2, device=dev())
    w = pose_oracle.DynamicLossWeighting()
    for step in range(2):
        hm = rnd(B, K, 48, 48, seed=step)
        thm = rnd(B, K, 48, 48, seed=10 + step).abs()
        kps = torch.cat([rnd(B, K, 2, seed=20 + step), torch.randint(0, 3, (B, K, 1), device=dev()).float()], -1).contiguous()
        z, tz = rnd(B, K, seed=30 + step), rnd(B, K, seed=40 + step)
        dhm, dz = torch.zeros_like(hm), torch.zeros_like(z)
        run(lambda b: b.pose_loss(hm, thm, kps, z, tz, sums, state, out, scales, dhm, dz, B=B, K=K, HW=HW))
        hm_r, z_r = hm.clone().requires_grad_(True), z.clone().requires_grad_(True)
        conf = kps[..., 2]
        kp, zl = pose_oracle.keypoint_loss(hm_r, thm, conf), pose_oracle.z_loss(z_r, tz, conf)
        w.update(kp.item(), zl.item())
        loss = w.balanced(kp, zl)
        g_hm, g_z = torch.autograd.grad(loss, (hm_r, z_r))
        assert abs(out[0].item() - loss.item()) < 1e-5 * abs(loss.item())
        assert abs(out[1].item() - kp.item()) < 1e-5 * abs(kp.item())
        assert abs(out[2].item() - zl.item()) < 1e-5 * abs(zl.item())
        assert rel(dhm, g_hm) < 1e-4
        assert rel(dz, g_z) < 1e-5
        assert sums.abs().max().item() == 0
    assert abs(state[3].item() - w.weight) < 1e-5


def test_adamw_vs_torch():
    n = 4096 * 5 + 64
    p0, lr, wd, eps = rnd(n, seed=1), 3e-3, 1e-2, 1e-8
    p = p0.clone()
    m, v = torch.zeros(n, device=dev()), torch.zeros(n, device=dev())
    step = torch.zeros((), dtype=torch.int64, device=dev())
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=lr, weight_decay=wd, eps=eps)
    for s in range(4):
        g = rnd(n, seed=50 + s, scale=0.1)
        run(lambda b: b.adamw(p, g, m, v, step, n=n, lr=lr, beta1=0.9, beta2=0.999, eps=eps, weight_decay=wd,
                              grad_scale=0.5))
        ref.grad = g * 0.5
        opt.step()
        assert rel(p, ref.detach()) < 2e-6, s
    assert int(step.item()) == 4


def test_pack_weights_and_counters():
    """dp_pack_weights_bf16: permuted / tap-mirrored bf16 GEMM layouts of fp32 conv weights, all jobs in one launch."""
    w1 = rnd(96, 64, 3, 3)
    w2 = rnd(128, 32, 4, 4, seed=1)
    w3 = rnd(24, 64, 1, 1, seed=2)
    d1 = torch.zeros(96, 9 * 64, device=dev(), dtype=BF)
    d2 = torch.zeros(64, 9 * 96, device=dev(), dtype=BF)
    d3 = torch.zeros(16 * 32, 128, device=dev(), dtype=BF)
    d4 = torch.zeros(64, 32, device=dev(), dtype=BF)
    jobs = [(d1, w1, (0, 2, 3, 1), (), None), (d2, w1, (1, 2, 3, 0), (2, 3), None), (d3, w2, (2, 3, 1, 0), (), None),
            (d4, w3, (1, 0, 2, 3), (), (32, 1, 1, 1))]
    c = [torch.full((), 5, device=dev(), dtype=torch.int64) for _ in range(3)]

    def fn(b):
        b.pack_weights(jobs)
        b.add_i64(c, 2)
    run(fn)
    assert torch.equal(d1, w1.permute(0, 2, 3, 1).reshape(96, -1).to(BF))
    assert torch.equal(d2, w1.flip(2, 3).permute(1, 2, 3, 0).reshape(64, -1).to(BF))
    assert torch.equal(d3, w2.permute(2, 3, 1, 0).reshape(-1, 128).to(BF))
    assert torch.equal(d4[:, :24], w3.reshape(24, 64).t().to(BF)) and d4[:, 24:].abs().max().item() == 0
    assert all(int(t.item()) == 7 for t in c)
