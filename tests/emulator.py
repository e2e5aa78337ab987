"""Plain-torch emulation of the C-ABI op vocabulary (TEST INFRASTRUCTURE ONLY).

``TorchEmulator`` has the same methods as ``dino_pose_b200.backend.CudaBackend`` but executes each op
with torch on whatever device the tensors live on (CPU in this container).  It exists so the engine's
host logic -- weight packing layouts, tap / index conventions, the forward and backward op graphs -- can
be validated against the oracle WITHOUT a GPU.  It mirrors the kernel SPECS (include/dinopose.h), it is
not a fallback: the package never imports it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BF = torch.bfloat16


class EmuProgram:
    def __init__(self):
        self.calls = []

    def __len__(self):
        return len(self.calls)

    def run(self, on_mark=None):
        for fn in self.calls:
            if isinstance(fn, tuple):
                if on_mark is not None:
                    on_mark(fn[1])
            else:
                fn()


class TorchEmulator:
    name = "emulator"

    def __init__(self):
        self.prog = None

    def begin(self):
        self.prog = EmuProgram()
        return self.prog

    def host(self, name, fn):
        self.prog.calls.append(fn)

    def fork(self, k=1):     # stream fork / join of the CUDA backend: the emulator runs everything in program order
        pass

    def side(self, flag):
        pass

    def join(self, k=1):
        pass

    def sync(self, which, k=1, src=0):
        pass

    def mark(self, tag):
        self.prog.calls.append(("mark", tag))

    def pack_weights(self, jobs):
        def fn():
            with torch.no_grad():
                for dst, w, order, flips, dst_strides in jobs:
                    src = w.detach()
                    if flips:
                        src = src.flip(*sorted(flips))
                    src = src.permute(*order)
                    if dst_strides is None:
                        dst.view(-1)[:src.numel()].view(src.shape).copy_(src)
                    else:
                        torch.as_strided(dst, tuple(src.shape), tuple(dst_strides)).copy_(src)
        self.prog.calls.append(fn)

    def add_i64(self, tensors, inc=1):
        def fn():
            for t in tensors:
                t.add_(inc)
        self.prog.calls.append(fn)

    # ------------------------------------------------------------------ training step around the model
    def pose_loss(self, hm, thm, kps, z, tz, sums, state, out, scales, dhm, dz, *, B, K, HW, momentum=0.9, rate=0.1):
        def fn():
            mask = (kps[..., 2] > 1).float()
            d = hm.view(B, K, HW) - thm.view(B, K, HW)
            q = d * d
            kp = (torch.exp(-q) * q * mask.view(B, K, 1)).double().sum() / (B * K * HW)
            zd = z * mask.view(B, K) - tz * mask.view(B, K)
            zl = zd.abs().double().sum() / (B * K)
            kp, zl = kp.float(), zl.float()
            if state[2] == 0:
                kp_avg, z_avg = kp, zl
            else:
                kp_avg = momentum * state[0] + (1 - momentum) * kp
                z_avg = momentum * state[1] + (1 - momentum) * zl
            w = ((1 - rate) * state[3] + rate * (kp + 1e-8) / (zl + 1e-8)).clamp(1e-3, 10.0)
            state[0], state[1], state[2], state[3] = kp_avg, z_avg, 1.0, w
            out[0], out[1], out[2] = kp / (kp_avg + 1e-8) + zl / (z_avg + 1e-8), kp, zl
            s0 = 2.0 / (B * K * HW * (kp_avg + 1e-8))
            s1 = 1.0 / (B * K * (z_avg + 1e-8))
            dhm.view(B, K, HW).copy_(s0 * torch.exp(-q) * d * mask.view(B, K, 1))
            dz.view(B, K).copy_(s1 * mask.view(B, K) * torch.sign(zd))
        self.prog.calls.append(fn)

    def adamw(self, p, g, m, v, step_dev, *, n, lr, beta1, beta2, eps, weight_decay, grad_scale, hyper=None, bump=True):
        lr0, wd0 = lr, weight_decay

        def fn():
            lr, weight_decay = (float(hyper[0]), float(hyper[1])) if hyper is not None else (lr0, wd0)
            t = int(step_dev.item()) + 1
            gg = g[:n] * grad_scale
            p[:n].mul_(1 - lr * weight_decay)
            m[:n].mul_(beta1).add_(gg, alpha=1 - beta1)
            v[:n].mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
            denom = v[:n].sqrt() / (1 - beta2 ** t) ** 0.5 + eps
            p[:n].addcdiv_(m[:n], denom, value=-lr / (1 - beta1 ** t))
            if bump:
                step_dev.add_(1)
        self.prog.calls.append(fn)

    # ------------------------------------------------------------------ GEMM family
    def gemm(self, A, W, out, *, M, N, K, lda=None, ldw=None, ldo=None, out_dtype="bf16", bias=None, scale=None,
             ls=None, residual=None, ldr=None, aux_out=None, aux_in=None, ld_aux=0, act="none", row_map="identity",
             n_valid=0, map_a=0, map_b=0, conv=None, OH=0, OW=0, NB=0, block_n=0, stats=None, stats_c=0, cta_pair=0,
             ln=None, lora=None, name="gemm"):
        def fn():
            Wf = W.float()[:N, :K]
            if conv is not None:
                nb, ih, iw, c = A.shape
                k, pad, oh, ow = conv["KH"], conv["pad"], conv["OH"], conv["OW"]
                x = A.float().permute(0, 3, 1, 2)
                x = F.pad(x, (pad, ow + k - 1 - pad - iw, pad, oh + k - 1 - pad - ih))
                w4 = Wf.view(N, k, k, c).permute(0, 3, 1, 2)
                acc = F.conv2d(x, w4).permute(0, 2, 3, 1).reshape(nb * oh * ow, N)
            else:
                acc = A.float()[:M, :K] @ Wf.t()
            nv = n_valid if n_valid > 0 else N
            v = acc[:, :nv]
            if scale is not None:
                v = v * scale[:nv]
            if bias is not None:
                v = v + bias[:nv]
            if stats is not None:
                sc = stats_c if stats_c > 0 else nv
                s1, s2 = v.double().sum(0), (v.double() ** 2).sum(0)
                if row_map == "shuffle2x2":
                    s1, s2 = s1.view(-1, map_a).sum(0), s2.view(-1, map_a).sum(0)
                stats[:sc] += s1
                stats[sc:2 * sc] += s2
            if aux_out is not None:
                side = v
                if act == "gelu":       # what the backward multiplies with: gelu'(v)
                    pre = v.detach().clone().requires_grad_(True)
                    with torch.enable_grad():
                        F.gelu(pre).sum().backward()
                    side = pre.grad
                aux_out.view(-1, ld_aux)[:M, :nv] = side.to(aux_out.dtype)
            if act == "relu":
                v = F.relu(v)
            elif act == "gelu":
                v = F.gelu(v)
            if aux_in is not None:
                v = v * aux_in.view(-1, ld_aux)[:M, :nv].float()
            if lora is not None:       # fused LoRA adapter: v + scaling * dropout(v A B); y and u saved for the backward
                u = v @ lora["A"].float()
                if lora.get("y_out") is not None:
                    lora["y_out"][:M, :nv] = v
                if lora.get("u_out") is not None:
                    lora["u_out"][:M] = u
                d = u @ lora["B"].float()
                assert float(lora.get("p_drop", 0.0)) == 0.0, "emulator: dropout parity is tested statistically on the GPU only"
                v = v + d * float(lora["scaling"])
            if ls is not None:
                v = v * ls[:nv]
            if row_map == "identity":
                if residual is not None:
                    v = v + residual.float()[:M, :nv]
                out[:M, :nv] = v.to(out.dtype)
                if ln is not None:      # fused LayerNorm of the output rows (row-owning kernel)
                    y = F.layer_norm(v.float(), (nv,), ln["gamma"], ln["beta"], ln.get("eps", 1e-6))
                    ln["out"][:M, :nv] = y.to(ln["out"].dtype)
            elif row_map == "patch_tokens":
                Bn = M // map_a
                v = v.view(Bn, map_a, nv) + residual[1:1 + map_a, :nv]
                out.view(Bn, map_b, -1)[:, 1:, :nv] = v.to(out.dtype)
            elif row_map == "nchw":
                hw = OH * OW if conv is None else conv["OH"] * conv["OW"]
                oh_, ow_ = (OH, OW) if conv is None else (conv["OH"], conv["OW"])
                out.copy_(v.view(-1, oh_, ow_, nv).permute(0, 3, 1, 2).to(out.dtype))
            elif row_map == "shuffle2x2":
                co = map_a
                nb = M // (OH * OW)
                v = v.view(nb, OH, OW, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(nb * 2 * OH * 2 * OW, co)
                out[:, :co] = v.to(out.dtype)
        self.prog.calls.append(fn)

    def wgrad(self, A, B, out, *, Mc, Nc, so_m, so_n, so_t=0, so_mo=0, so_no=0, m_inner=0, n_inner=0, conv=None,
              P=0, lda=None, ldb=None, block_n=0, splits=0, workspace=None, name="wgrad"):
        def fn():
            flat = out.view(-1)
            shift = 0
            if conv is not None and so_t < 0:
                # mirrored taps: `out` points at the LAST tap and the kernel walks backwards (so_t = -1)
                shift = (conv["KH"] * conv["KW"] - 1) * (-so_t)
                flat = torch.as_strided(out, (out.numel() + shift,), (1,), out.storage_offset() - shift)
            m = torch.arange(Mc, device=out.device)
            n = torch.arange(Nc, device=out.device)
            mi = m_inner if m_inner > 0 else 1 << 30
            ni = n_inner if n_inner > 0 else 1 << 30
            offm = (m % mi) * so_m + (m // mi) * so_mo
            offn = (n % ni) * so_n + (n // ni) * so_no
            if conv is None:
                g = A.float()[:P, :Mc].t() @ B.float()[:P, :Nc]
                flat.index_put_(((offm[:, None] + offn[None, :]).reshape(-1),), g.reshape(-1), accumulate=True)
            else:
                k, pad = conv["KH"], conv["pad"]
                nb, oh, ow, _ = A.shape
                _, ih, iw, _ = B.shape
                xb = F.pad(B.float().permute(0, 3, 1, 2), (pad, ow + k - 1 - pad - iw, pad, oh + k - 1 - pad - ih))
                a2 = A.float().reshape(nb * oh * ow, -1)[:, :Mc]
                for ky in range(k):
                    for kx in range(k):
                        sh = xb[:, :Nc, ky:ky + oh, kx:kx + ow].permute(0, 2, 3, 1).reshape(nb * oh * ow, Nc)
                        g = a2.t() @ sh
                        idx = (offm[:, None] + offn[None, :] + (ky * k + kx) * so_t + shift).reshape(-1)
                        flat.index_put_((idx,), g.reshape(-1), accumulate=True)
        self.prog.calls.append(fn)

    # ------------------------------------------------------------------ backbone
    def layernorm_fwd(self, x, gamma, beta, y_bf16, y_f32, *, rows, D, T=0, drop_cls=False, eps=1e-6):
        def fn():
            y = F.layer_norm(x.float()[:rows], (D,), gamma, beta, eps)
            if drop_cls:
                y = y.view(-1, T, D)[:, 1:].reshape(-1, D)
            if y_bf16 is not None:
                y_bf16.copy_(y.to(y_bf16.dtype))
            if y_f32 is not None:
                y_f32.copy_(y)
        self.prog.calls.append(fn)

    def layernorm_bwd(self, dy, x, gamma, add_in, dx, *, rows, D, T=0, drop_cls=False, eps=1e-6, ls=None,
                      dx_scaled=None):
        def fn():
            g = dy.float()
            if drop_cls:
                full = torch.zeros(rows // T, T, D, device=g.device)
                full[:, 1:] = g.view(-1, T - 1, D)
                g = full.view(rows, D)
            xr = x.detach().clone().float().requires_grad_(True)
            with torch.enable_grad():
                F.layer_norm(xr, (D,), gamma.detach(), None, eps).backward(g)
            o = xr.grad
            if add_in is not None:
                o = o + add_in
            dx.copy_(o)
            if dx_scaled is not None:
                dx_scaled.copy_((o * ls).to(dx_scaled.dtype))
        self.prog.calls.append(fn)

    def patch_im2col(self, px, out, *, B, H, W, Kp):
        def fn():
            cols = F.unfold(px, 14, stride=14).transpose(1, 2).reshape(-1, 588)
            out.zero_()
            out[:, :588] = cols.to(out.dtype)
        self.prog.calls.append(fn)

    def fill_cls(self, x, cls_row, *, B, T, D):
        def fn():
            x.view(B, T, D)[:, 0] = cls_row
        self.prog.calls.append(fn)

    def lora_fwd(self, y, A, Bm, lambda1, x_in, x_out, u_save, *, rows, D, R, scaling, p_drop, seed):
        def fn():
            assert p_drop == 0.0, "emulator: dropout parity is tested statistically on the GPU only"
            u = y @ A.detach()
            if u_save is not None:
                u_save.copy_(u)
            x_out.copy_(x_in + (y + (u @ Bm.detach()) * scaling) * lambda1.detach())
        self.prog.calls.append(fn)

    def lora_bwd(self, g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, *, rows, D, R, scaling, p_drop, seed):
        def fn():
            gv = g * lambda1.detach() * scaling
            dB.add_(u_saved.t() @ gv)
            dA.add_(y.t() @ (gv @ Bm.detach().t()))
        self.prog.calls.append(fn)

    def attention_fwd(self, qkv, ctx, *, B, T, heads, scale):
        def fn():
            D = heads * 64
            q, k, v = [t.view(B, T, heads, 64).transpose(1, 2) for t in qkv.float().view(B, T, 3 * D).split(D, -1)]
            p = torch.softmax(q @ k.transpose(2, 3) * scale, -1)
            ctx.copy_((p.to(ctx.dtype).float() @ v).transpose(1, 2).reshape(B * T, D).to(ctx.dtype))
        self.prog.calls.append(fn)

    def attention_bwd(self, qkv, ctx, dctx, dqkv, stats, *, B, T, heads, scale):
        def fn():
            D = heads * 64
            q, k, v = [t.reshape(B, T, heads, 64).transpose(1, 2).detach().clone().requires_grad_(True)
                       for t in qkv.float().view(B, T, 3 * D).split(D, -1)]
            with torch.enable_grad():
                o = (torch.softmax(q @ k.transpose(2, 3) * scale, -1) @ v).transpose(1, 2).reshape(B * T, D)
                o.backward(dctx.float())
            g = torch.cat([t.grad.transpose(1, 2).reshape(B * T, D) for t in (q, k, v)], -1)
            dqkv.copy_(g.to(dqkv.dtype))
        self.prog.calls.append(fn)

    def layernorm_bwd_params(self, dy, x, dgamma, dbeta, *, rows, D, eps=1e-6):
        def fn():
            xf = x.float()[:rows]
            xhat = (xf - xf.mean(-1, keepdim=True)) * torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + eps)
            g = dy.float()[:rows]
            dgamma.add_((g * xhat).sum(0))
            dbeta.add_(g.sum(0))
        self.prog.calls.append(fn)

    def colsum_prod(self, g, a, out, *, P, C):
        def fn():
            out.view(-1)[:C].add_((g.float().view(P, C) * a.float().view(P, C)).sum(0))
        self.prog.calls.append(fn)

    def pred1x1_fwd(self, a, w, bias, out, *, P, HW, C, K):
        def fn():
            y = a.float().view(P, C) @ w.detach().view(K, C).t() + bias.detach()
            out.copy_(y.view(P // HW, HW, K).permute(0, 2, 1).reshape(out.shape))
        self.prog.calls.append(fn)

    def pred1x1_bwd(self, g, a, w, d, dW, db, *, P, HW, C, K):
        def fn():
            gp = g.reshape(P // HW, K, HW).permute(0, 2, 1).reshape(P, K)
            d.copy_((gp @ w.detach().view(K, C)).to(d.dtype))
            dW.view(K, C).add_(gp.t() @ a.float().view(P, C))
            db.add_(gp.sum(0))
        self.prog.calls.append(fn)

    # ------------------------------------------------------------------ heads
    def im2col(self, x, col, *, NB, IH, IW, C, OH, OW, KH, KW, stride, pad):
        def fn():
            xx = x.float().reshape(NB, IH, IW, C).permute(0, 3, 1, 2)
            need_h = (OH - 1) * stride + KH
            need_w = (OW - 1) * stride + KW
            xx = F.pad(xx, (pad, need_w - pad - IW, pad, need_h - pad - IH))
            u = F.unfold(xx, (KH, KW), stride=stride)            # [NB, C*KH*KW, OH*OW]
            u = u.view(NB, C, KH * KW, OH * OW).permute(0, 3, 2, 1).reshape(NB * OH * OW, KH * KW * C)
            col.copy_(u.to(col.dtype))
        self.prog.calls.append(fn)

    def col2im(self, col, bias, big, *, NB, SH, SW, C, BH, BW, KH, KW, stride, pad):
        def fn():
            c = col.float().view(NB, SH * SW, KH * KW, C).permute(0, 3, 2, 1).reshape(NB, C * KH * KW, SH * SW)
            full_h = (SH - 1) * stride + KH
            full_w = (SW - 1) * stride + KW
            o = F.fold(c, (full_h, full_w), (KH, KW), stride=stride)   # [NB, C, full_h, full_w]
            canvas = torch.zeros(NB, C, max(full_h, pad + BH), max(full_w, pad + BW), device=o.device)
            canvas[:, :, :full_h, :full_w] = o
            o = canvas[:, :, pad:pad + BH, pad:pad + BW]
            if bias is not None:
                o = o + bias.detach().view(1, C, 1, 1)
            big.view(NB, BH, BW, C).copy_(o.permute(0, 2, 3, 1).to(big.dtype))
        self.prog.calls.append(fn)

    def dwconv3x3(self, x, w, bias, add, out, *, NB, H, W, C, flip=False):
        def fn():
            xx = x.float().reshape(NB, H, W, C).permute(0, 3, 1, 2)
            ww = w.detach().flip(2, 3) if flip else w.detach()
            o = F.conv2d(xx, ww, None if bias is None else bias.detach(), 1, 1, 1, C).permute(0, 2, 3, 1)
            if add is not None:
                o = o + add.float().reshape(NB, H, W, C)
            out.view(NB, H, W, C).copy_(o.to(out.dtype))
        self.prog.calls.append(fn)

    def dwconv3x3_wgrad(self, x, dout, dw, *, NB, H, W, C):
        def fn():
            xx = x.float().reshape(NB, H, W, C).permute(0, 3, 1, 2)
            w = torch.zeros(C, 1, 3, 3, requires_grad=True, device=x.device)
            with torch.enable_grad():
                F.conv2d(xx, w, None, 1, 1, 1, C).backward(dout.float().reshape(NB, H, W, C).permute(0, 3, 1, 2))
            dw.add_(w.grad)
        self.prog.calls.append(fn)

    def bn_stats(self, raw, sums, *, P, C):
        def fn():
            r = raw.float().view(P, C).double()
            sums[:C] += r.sum(0)
            sums[C:2 * C] += (r * r).sum(0)
        self.prog.calls.append(fn)

    def bn_finalize(self, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, *, C, count, eps=1e-5, momentum=0.1):
        def fn():
            m = sums[:C] / count
            var = (sums[C:2 * C] / count - m * m).clamp_min(0)
            inv = (1.0 / torch.sqrt(var + eps)).float()
            scale.copy_(gamma.detach() * inv)
            shift.copy_(beta.detach() - m.float() * scale)
            mean.copy_(m.float())
            invstd.copy_(inv)
            if rm is not None:
                rm.mul_(1 - momentum).add_(momentum * m.float())
                rv.mul_(1 - momentum).add_(momentum * (var * count / (count - 1)).float())
            sums.zero_()
        self.prog.calls.append(fn)

    def bn_finalize_apply(self, raw, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, add1, add2, out, *, P, C,
                          relu=True, mode=0, eps=1e-5, momentum=0.1):
        self.bn_finalize(sums, gamma, beta, rm, rv, scale, shift, mean, invstd, C=C, count=P, eps=eps, momentum=momentum)
        self.bn_apply(raw, scale, shift, add1, add2, out, P=P, C=C, relu=relu, mode=mode)

    def bn_fold_eval(self, gamma, beta, rm, rv, conv_bias, scale, shift, *, C, eps=1e-5, mean=None, invstd=None):
        def fn():
            sc = gamma.detach() * torch.rsqrt(rv + eps)
            scale.copy_(sc)
            cb = conv_bias.detach() if conv_bias is not None else 0.0
            shift.copy_(beta.detach() + (cb - rm) * sc)
            if mean is not None:
                mean.copy_(rm)
            if invstd is not None:
                invstd.copy_(torch.rsqrt(rv + eps))
        self.prog.calls.append(fn)

    def bn_apply(self, raw, scale, shift, add1, add2, out, *, P, C, relu=True, mode=0):
        def fn():
            y = raw.float().view(P, C) * scale + shift
            if mode == 1:
                y = F.relu(y + add1.float().view(P, C))
            else:
                if relu:
                    y = F.relu(y)
                if add1 is not None:
                    y = y + add1.float().view(P, C)
                if add2 is not None:
                    y = y + add2.float().view(P, C)
            out.view(P, C).copy_(y.to(out.dtype))
        self.prog.calls.append(fn)

    def _masked_dy(self, dout, raw, add1, scale, shift, P, C, relu, mode):
        r = raw.float().view(P, C)
        y = r * scale + shift
        if mode == 1:
            y = y + add1.float().view(P, C)
        g = dout.float().view(P, C)
        if relu or mode == 1:
            g = g * (y > 0)
        return r, g

    def bn_bwd_reduce(self, dout, raw, add1, scale, shift, mean, invstd, sums, *, P, C, relu=True, mode=0):
        def fn():
            r, g = self._masked_dy(dout, raw, add1, scale, shift, P, C, relu, mode)
            xhat = (r - mean) * invstd
            sums[:C] += g.double().sum(0)
            sums[C:2 * C] += (g * xhat).double().sum(0)
        self.prog.calls.append(fn)

    def bn_bwd_apply(self, dout, raw, add1, gamma, scale, shift, mean, invstd, sums, draw, dres, dgamma, dbeta, *, P, C,
                     relu=True, mode=0, eval_mode=False, shuffle_oh=0, shuffle_ow=0):
        def fn():
            r, g = self._masked_dy(dout, raw, add1, scale, shift, P, C, relu, mode)
            if eval_mode:
                o = g * scale
            else:
                xhat = (r - mean) * invstd
                s1 = (sums[:C] / P).float()
                s2 = (sums[C:2 * C] / P).float()
                o = gamma.detach() * invstd * (g - s1 - xhat * s2)
            if shuffle_oh > 0:
                nb = P // (shuffle_oh * shuffle_ow)
                ih, iw = shuffle_oh // 2, shuffle_ow // 2
                o = o.view(nb, ih, 2, iw, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(nb * ih * iw, 4 * C)
                draw.view(nb * ih * iw, 4 * C).copy_(o.to(draw.dtype))
            else:
                draw.view(P, C).copy_(o.to(draw.dtype))
            if dres is not None:
                dres.view(P, C).copy_(g.to(dres.dtype))
            if dgamma is not None:
                dbeta.copy_(sums[:C].float())
                dgamma.copy_(sums[C:2 * C].float())
            sums.zero_()      # dp_bn_bwd_apply re-zeroes the accumulators
        self.prog.calls.append(fn)

    def avgpool2(self, x, out, *, planes, OH, OW):
        def fn():
            out.copy_(F.avg_pool2d(x.view(planes, 1, 2 * OH, 2 * OW), 2).view(out.shape))
        self.prog.calls.append(fn)

    def hm_grad_to_nhwc(self, g, out, *, NB, K, Kp, OH, OW, up):
        def fn():
            gg = g
            if up == 2:
                gg = 0.25 * g.repeat_interleave(2, 2).repeat_interleave(2, 3)
            out.zero_()
            out.view(NB, OH, OW, Kp)[..., :K] = gg.permute(0, 2, 3, 1).to(out.dtype)
        self.prog.calls.append(fn)

    def mean_tokens(self, feat, out, *, B, N, D):
        def fn():
            out.copy_(feat.float().view(B, N, D).mean(1))
        self.prog.calls.append(fn)

    def mean_tokens_bwd(self, dfeat, dmean, *, B, N, D):
        def fn():
            v = dfeat.float().view(B, N, D) + (dmean / N)[:, None, :]
            dfeat.view(B, N, D).copy_(v.to(dfeat.dtype))
        self.prog.calls.append(fn)

    def sgemm_small(self, A, sa_m, sa_k, Bm, sb_k, sb_n, Cm, ldc, *, M, N, K, bias=None, relu=False, mask_ref=None,
                    ld_ref=0, p_drop=0.0, seed=None, accumulate=False):
        def fn():
            assert p_drop == 0.0
            a = torch.as_strided(A.detach(), (M, K), (sa_m, sa_k))
            b = torch.as_strided(Bm.detach(), (K, N), (sb_k, sb_n))
            v = a @ b
            if bias is not None:
                v = v + bias.detach()
            if relu:
                v = F.relu(v)
            if mask_ref is not None:
                v = v * (torch.as_strided(mask_ref, (M, N), (ld_ref, 1)) > 0)
            c = torch.as_strided(Cm, (M, N), (ldc, 1))
            if accumulate:
                c.add_(v)
            else:
                c.copy_(v)
        self.prog.calls.append(fn)

    def relu_mask(self, d, ref, out, *, n, keep_scale=1.0):
        def fn():
            out.copy_(torch.where(ref > 0, d * keep_scale, torch.zeros_like(d)))
        self.prog.calls.append(fn)

    def colsum(self, x, out, *, P, C, ld):
        def fn():
            out.view(-1)[:C].add_(torch.as_strided(x, (P, C), (ld, 1)).float().sum(0))
        self.prog.calls.append(fn)

    def decode(self, hm, idx, xy, conf, *, maps, H, W, target_w, target_h):
        from oracle import decode_oracle

        def fn():
            i, p = decode_oracle.decode_batch(hm.view(1, maps, H, W).cpu().numpy(), (target_w, target_h))
            idx.copy_(torch.from_numpy(i[0]).to(idx.dtype))
            xy.copy_(torch.from_numpy(p[0]))
        self.prog.calls.append(fn)
