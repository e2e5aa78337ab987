"""Host-side engine: lowers the dino_pose forward / backward onto the sm_100a kernels.

The reference executes this path as a chain of ATen calls driven by Python (``train.py:144,169`` ->
``model/dinov2_pose.py:292-306`` -> HF ``Dinov2Model`` -> ``model/pose_heads.py:395-400``).  Here the same
chain is described ONCE per plan -- (batch, height, width, training) -- as a recorded program of
C-ABI launches on statically allocated device buffers (see ``backend.py``); a step replays it.

Data layout in HBM (B images, T = 1 + N tokens, D channels):
  residual stream x            fp32 [B*T, D]
  LayerNorm outputs / qkv / ctx / MLP hidden / head activations     bf16, rows = tokens or NHWC pixels
  weights                      bf16 [N, K] K-major copies of the fp32 nn.Parameters (packed here)
  gradients of parameters      fp32, one flat buffer, views per parameter
  heat-maps / z                fp32 [B,K,48,48] / [B,K]

What autograd reaches in the reference: heads, final LayerNorm, MLP branch of the last
block, LoRA adapter.  The backward program implements exactly that path; everything earlier is frozen
(``model/dinov2_pose.py:193-194``).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
F32 = torch.float32
PATCH = 14
FROZEN_HEADS = 2                     # `training` value of a training step whose pose heads are in eval mode
PATCH_K = 3 * PATCH * PATCH          # 588
PATCH_KP = 640                       # padded to a multiple of the 64-wide k-block
LN_EPS = 1e-6
HM_PAD = 32                          # heat-map gradient channels padded 24 -> 32 (16-byte TMA pitch)


def _ceil(a, b):
    return (a + b - 1) // b


class ConvLayer:
    """One conv(+BN+ReLU) unit of the heads: geometry, parameter names, packed weights, saved tensors."""

    def __init__(self, name, kind, cin, cout, k, ih, iw, oh, ow, bn=None, relu=True, stride=1, pad=0):
        self.name, self.kind = name, kind
        self.cin, self.cout, self.k = cin, cout, k
        self.ih, self.iw, self.oh, self.ow = ih, iw, oh, ow
        self.bn, self.relu = bn, relu
        self.stride, self.pad = stride, pad
        self.t = {}   # tensors


class PoseEngine:
    def __init__(self, params, buffers, cfg, backend, device):
        """params / buffers: dict reference-state_dict-name -> tensor (fp32 parameters / BN buffers).
        cfg: dict(D, L, heads, num_keypoints, heatmap_size, lora=None|dict(rank, alpha, dropout),
                  z_hidden, z_dropout)."""
        self.P, self.Bufs, self.cfg = params, buffers, cfg
        self.be = backend
        self.device = device
        self.D, self.L, self.heads = cfg["D"], cfg["L"], cfg["heads"]
        self.K = cfg["num_keypoints"]
        self.hm_size = cfg["heatmap_size"]
        self.lora = cfg.get("lora")
        # Dinov2PoseModel(unfreeze_last_n_layers=n) (reference model/dinov2_pose.py:25-39): the last n encoder layers train
        self.unfreeze = max(0, min(int(cfg.get("unfreeze", 0) or 0), cfg["L"]))
        self.train_layers = list(range(cfg["L"] - self.unfreeze, cfg["L"]))
        # storage dtype of activations / packed weights.  The CUDA kernels are bf16-only; fp32 is accepted only by
        # the torch emulator (tests) to separate logic errors from bf16 rounding.
        self.adt = cfg.get("act_dtype", BF16)
        if self.adt != BF16 and getattr(backend, "name", "") == "cuda":
            raise ValueError("the sm_100a kernels store activations in bf16 only")
        # storage dtype of pre-BatchNorm conv outputs in training ("raw")
        self.rdt = cfg.get("raw_dtype", F32)
        self.plans = {}
        self.frozen = None
        self.seed = None
        self.step_count = 0
        last = self.L - 1
        self.att_prefix = [f"backbone.encoder.layer.{i}.attention." for i in range(self.L)]
        if self.lora:
            self.att_prefix[last] = f"backbone.encoder.layer.{last}.attention.original_attention."
        self.lora_prefix = f"backbone.encoder.layer.{last}.attention.lora_output."

    # ------------------------------------------------------------------ helpers
    def new(self, shape, dtype, fill=None):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        if fill is None:
            t.zero_()
        else:
            t.fill_(fill)
        return t

    def p(self, name):
        t = self.P.get(name)
        if t is None:
            t = self.Bufs[name]
        return t

    def trainable_names(self):
        out = []
        for n, t in self.P.items():
            if t.requires_grad:
                out.append(n)
        return out

    # ------------------------------------------------------------------ flat gradient / parameter layout
    BWD_LAYER_ORDER = ("pred3", "pred0", "ups1", "ups0", "fr4", "up2", "up1", "bt2", "bt1", "down2", "down1", "pw", "dw",
                       "skip", "fr0")
    GRAD_ALIGN = 64   # elements (256 B): every parameter's slice can be used as a 16-byte vector / TMA base

    def layout(self):
        """Order and offsets of the trainable parameters in the flat gradient buffer (and in the trainer's flat
        parameter / AdamW-state buffers): the order in which the backward program FINISHES them, so that a prefix
        of the buffer is final at every ``mark`` and can be all-reduced while the rest is still being computed."""
        if getattr(self, "_layout", None) is not None:
            return self._layout
        trainable = set(self.trainable_names())
        Ls = self.build_head_layers(16)
        order = []
        # the z-head backward (a chain of small launches on the second stream) is the FIRST thing the backward finishes:
        # its 4 MB of gradients ride in the first all-reduce bucket instead of an exposed one after the last kernel
        nz = len(self.cfg["z_hidden"]) + 1
        zg = []
        for j in reversed(range(nz)):
            zg += [f"pose_heads.z_head.mlp.{3 * j}.weight", f"pose_heads.z_head.mlp.{3 * j}.bias"]
        order.append(("z_head", [n for n in zg if n in trainable]))
        for key in self.BWD_LAYER_ORDER:
            L = Ls[key]
            group = ([L.bn + ".weight", L.bn + ".bias"] if L.bn else []) + [L.name + ".weight", L.name + ".bias"]
            order.append((key, [n for n in group if n in trainable]))
        lg = [self.lora_prefix + "lora_A", self.lora_prefix + "lora_B"] if self.lora else []
        order.append(("lora", [n for n in lg if n in trainable]))
        for i in reversed(self.train_layers):
            order.append((f"layer{i}", [n for n in self.layer_param_names(i) if n in trainable]))
        seen = {n for _k, g in order for n in g}
        rest = [n for n in self.trainable_names() if n not in seen]
        if rest:
            order.append(("other", rest))
        offsets, ends, off = {}, {}, 0
        for key, group in order:
            for n in group:
                k = self.P[n].numel()
                offsets[n] = (off, k)
                off += _ceil(k, self.GRAD_ALIGN) * self.GRAD_ALIGN
            ends[key] = off
        self._layout = {"names": [n for _k, g in order for n in g], "offsets": offsets, "group_end": ends, "total": off}
        return self._layout

    def layer_param_names(self, i):
        """Parameters of encoder layer i in the order the backward program finishes them.  The q/k/v biases and the q/k/v
        weights are adjacent (and D, D*D are multiples of GRAD_ALIGN), so each triple is ONE contiguous slice of the flat
        gradient buffer: the fused [3D, D] QKV weight gradient is written by a single launch."""
        lp, ap = f"backbone.encoder.layer.{i}.", self.att_prefix[i]
        return [lp + "layer_scale2.lambda1", lp + "mlp.fc2.bias", lp + "mlp.fc2.weight", lp + "mlp.fc1.bias",
                lp + "mlp.fc1.weight", lp + "norm2.weight", lp + "norm2.bias", lp + "layer_scale1.lambda1",
                ap + "output.dense.bias", ap + "output.dense.weight",
                ap + "attention.query.bias", ap + "attention.key.bias", ap + "attention.value.bias",
                ap + "attention.query.weight", ap + "attention.key.weight", ap + "attention.value.weight",
                lp + "norm1.weight", lp + "norm1.bias"]

    # ------------------------------------------------------------------ frozen weight packing
    def pack_frozen(self):
        """bf16 K-major copies of the frozen backbone weights (plain torch layout ops, run once and
        whenever a frozen parameter's version counter changes)."""
        D, L = self.D, self.L
        fz = {}
        w = self.p("backbone.embeddings.patch_embeddings.projection.weight").detach()
        wpe = torch.zeros(D, PATCH_KP, dtype=self.adt, device=self.device)
        wpe[:, :PATCH_K] = w.reshape(D, PATCH_K).to(self.adt)
        fz["wpe"] = wpe
        for i in range(L):
            lp = f"backbone.encoder.layer.{i}."
            ap = self.att_prefix[i]
            q, k, v = (self.p(ap + f"attention.{n}.weight").detach() for n in ("query", "key", "value"))
            fz[f"wqkv{i}"] = torch.cat([q, k, v], 0).to(self.adt).contiguous()
            fz[f"bqkv{i}"] = torch.cat([self.p(ap + f"attention.{n}.bias").detach() for n in ("query", "key", "value")]).float().contiguous()
            fz[f"wo{i}"] = self.p(ap + "output.dense.weight").detach().to(self.adt).contiguous()
            fz[f"bo{i}"] = self.p(ap + "output.dense.bias").detach().float().contiguous()
            fz[f"w1{i}"] = self.p(lp + "mlp.fc1.weight").detach().to(self.adt).contiguous()
            fz[f"w2{i}"] = self.p(lp + "mlp.fc2.weight").detach().to(self.adt).contiguous()
        lp = f"backbone.encoder.layer.{L - 1}."
        # transposed copies for the last block's MLP input-gradient GEMMs
        fz["w2T"] = self.p(lp + "mlp.fc2.weight").detach().t().to(self.adt).contiguous()   # [4D, D]
        fz["w1T"] = self.p(lp + "mlp.fc1.weight").detach().t().to(self.adt).contiguous()   # [D, 4D]
        self.frozen = fz
        self.frozen_versions = self._versions(frozen=True)
        self.pos_cache = {}

    def _versions(self, frozen):
        return tuple(t._version for n, t in self.P.items() if (not t.requires_grad) == frozen)

    def pos_tables(self, H, W):
        """Position-embedding table for an HxW input: rows 1.. = bicubic-resized patch position
        embeddings (HF modeling_dinov2.py:57-95, constant per resolution for frozen parameters) + the
        patch-conv bias; row 0 = cls_token + position_embeddings[0] (HF:108-112)."""
        key = (H, W)
        if key in self.pos_cache:
            return self.pos_cache[key]
        pos = self.p("backbone.embeddings.position_embeddings").detach().float()
        D = self.D
        gh, gw = H // PATCH, W // PATCH
        npos = pos.shape[1] - 1
        s = int(math.sqrt(npos))
        if gh * gw == npos and gh == gw:
            patch = pos[0, 1:]
        else:
            grid = pos[0, 1:].reshape(1, s, s, D).permute(0, 3, 1, 2)
            grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
            patch = grid.permute(0, 2, 3, 1).reshape(gh * gw, D)
        table = torch.empty(1 + gh * gw, D, dtype=F32, device=self.device)
        bias = self.p("backbone.embeddings.patch_embeddings.projection.bias").detach().float()
        table[1:] = patch + bias
        table[0] = self.p("backbone.embeddings.cls_token").detach().float().reshape(D) + pos[0, 0]
        self.pos_cache[key] = table
        return table

    # ------------------------------------------------------------------ head layer table
    def build_head_layers(self, g):
        """g = side of the token grid (16 at 224^2, 32 at 448^2).  Geometry follows
        model/pose_heads.py:212-343 with spatial_input_size=16, heatmap_size=48 (dinov2_pose.py:45-54)."""
        hp = "pose_heads.heatmap_head."
        fr = hp + "feature_refine."
        hg = fr + "3."
        D = self.D
        g2, g4 = g // 2, g // 4
        s47 = (g - 1) * 3 - 2 + 4          # ConvT(k4,s3,p1): 16 -> 47, 32 -> 95
        s48 = s47 + 1                      # ConvT(k4,s1,p1): 47 -> 48, 95 -> 96
        Ls = {}

        def add(key, *a, **k):
            Ls[key] = ConvLayer(*a, **k)

        add("fr0", fr + "0", "conv", D, 512, 3, g, g, g, g, bn=fr + "1", pad=1)
        add("skip", hg + "skip.0", "conv", 512, 512, 1, g, g, g, g, bn=hg + "skip.1")
        add("dw", hg + "depthwise_conv.0", "dw", 512, 512, 3, g, g, g, g, bn=hg + "depthwise_conv.1", pad=1)
        add("pw", hg + "depthwise_conv.3", "conv", 512, 512, 1, g, g, g, g, bn=hg + "depthwise_conv.4")
        add("down1", hg + "down1.0", "conv_s2", 512, 256, 3, g, g, g2, g2, bn=hg + "down1.1", stride=2, pad=1)
        add("down2", hg + "down2.0", "conv_s2", 256, 128, 3, g2, g2, g4, g4, bn=hg + "down2.1", stride=2, pad=1)
        add("bt1", hg + "bottleneck.0", "conv", 128, 128, 3, g4, g4, g4, g4, bn=hg + "bottleneck.1", pad=1)
        add("bt2", hg + "bottleneck.3", "conv", 128, 128, 3, g4, g4, g4, g4, bn=hg + "bottleneck.4", pad=1, relu=False)
        add("up1", hg + "up1.0", "convT2", 128, 256, 2, g4, g4, g2, g2, bn=hg + "up1.1", stride=2)
        add("up2", hg + "up2.0", "convT2", 256, 512, 2, g2, g2, g, g, bn=hg + "up2.1", stride=2)
        add("fr4", fr + "4", "conv", 512, 256, 3, g, g, g, g, bn=fr + "5", pad=1)
        add("ups0", hp + "upsampling.0.0", "convT", 256, 128, 4, g, g, s47, s47, bn=hp + "upsampling.0.1", stride=3, pad=1)
        add("ups1", hp + "upsampling.1.0", "convT_s1", 128, 128, 4, s47, s47, s48, s48, bn=hp + "upsampling.1.1", pad=1)
        add("pred0", hp + "prediction.0", "conv", 128, 64, 3, s48, s48, s48, s48, bn=hp + "prediction.1", pad=1)
        add("pred3", hp + "prediction.3", "conv", 64, self.K, 1, s48, s48, s48, s48, bn=None, relu=False)
        return Ls

    # ------------------------------------------------------------------ trainable weight packing (recorded)
    def record_pack_heads(self, plan):
        """bf16 GEMM-layout copies of the trainable head weights; re-run every step in training (the
        optimizer just changed them) and on version change in eval."""
        Ls = plan["layers"]
        be = self.be

        def alloc(L):
            k, ci, co = L.k, L.cin, L.cout
            kk = k * k
            if L.kind == "dw":
                return
            if L.kind in ("conv", "conv_s2"):
                L.t["wf"] = self.new((co, kk * ci), self.adt)                 # [Cout, (ky,kx,ci)]
                L.t["wd"] = self.new((ci, kk * co), self.adt) if L.kind == "conv" else self.new((kk * ci, co), self.adt)
            elif L.kind in ("convT2", "convT"):
                L.t["wf"] = self.new((kk * co, ci), self.adt)                 # [(ky,kx,co), ci]
                L.t["wd"] = self.new((ci, kk * co), self.adt)                 # [ci, (ky,kx,co)]
            elif L.kind == "convT_s1":
                L.t["wf"] = self.new((co, kk * ci), self.adt)                 # conv form, flipped taps
                L.t["wd"] = self.new((ci, kk * co), self.adt)
            if L.name.endswith("prediction.3"):
                L.t["wd"] = self.new((ci, HM_PAD), self.adt)                  # [64, 32] zero padded K

        for L in Ls.values():
            alloc(L)

        # one launch re-packs every layer (dp_pack_weights_bf16): (dst, parameter, dim order, mirrored dims, dst strides)
        jobs = []
        for L in Ls.values():
            if L.kind == "dw":
                continue
            w = self.p(L.name + ".weight")
            ci, co = L.cin, L.cout
            if L.kind == "conv":
                jobs.append((L.t["wf"], w, (0, 2, 3, 1), (), None))                    # [co, (ky,kx,ci)]
                if L.name.endswith("prediction.3"):
                    jobs.append((L.t["wd"], w, (1, 0, 2, 3), (), (HM_PAD, 1, 1, 1)))   # [ci, co | zero padding]
                else:
                    # dgrad: dIn[y,x,ci] = sum dRaw[y+ky'-pad, x+kx'-pad, co] * W[co,ci,k-1-ky',k-1-kx']
                    jobs.append((L.t["wd"], w, (1, 2, 3, 0), (2, 3), None))            # [ci, (ky',kx',co)] mirrored taps
            elif L.kind == "conv_s2":
                jobs.append((L.t["wf"], w, (0, 2, 3, 1), (), None))
                jobs.append((L.t["wd"], w, (2, 3, 1, 0), (), None))                    # wf^T: [(ky,kx,ci), co]
            elif L.kind in ("convT2", "convT"):
                jobs.append((L.t["wf"], w, (2, 3, 1, 0), (), None))                    # [(ky,kx,co), ci]
                jobs.append((L.t["wd"], w, (0, 2, 3, 1), (), None))                    # [ci, (ky,kx,co)]
            elif L.kind == "convT_s1":
                # out[y,x,co] = sum in[y+ky'-2, x+kx'-2, ci] * Wt[ci,co,3-ky',3-kx']
                jobs.append((L.t["wf"], w, (1, 2, 3, 0), (2, 3), None))
                # dIn[iy,ix,ci] = sum dOut[iy-1+ky, ix-1+kx, co] * Wt[ci,co,ky,kx]
                jobs.append((L.t["wd"], w, (0, 2, 3, 1), (), None))
        be.pack_weights(jobs)


    def record_pack_layers(self, plan, training):
        """bf16 GEMM-layout copies of the UN-FROZEN encoder layers' weights (the optimizer changes them every step, so the
        re-packing is part of the forward program): [3D, D] fused QKV, projection, fc1, fc2 and -- for the input-gradient
        GEMMs of the backward -- their transposes; plus the concatenated fp32 QKV bias."""
        D = self.D
        lw = plan["lw"] = {}
        jobs, cats = [], []
        ident, transp = (0, 1, 2, 3), (1, 0, 2, 3)
        for i in self.train_layers:
            lp, ap = f"backbone.encoder.layer.{i}.", self.att_prefix[i]
            w = lw[i] = {}
            w["wqkv"] = self.new((3 * D, D), self.adt)
            w["bqkv"] = self.new((3 * D,), F32)
            w["wo"] = self.new((D, D), self.adt)
            w["w1"] = self.new((4 * D, D), self.adt)
            w["w2"] = self.new((D, 4 * D), self.adt)
            if training:
                w["wqkvT"] = self.new((D, 3 * D), self.adt)
                w["woT"] = self.new((D, D), self.adt)
                w["w1T"] = self.new((D, 4 * D), self.adt)
                w["w2T"] = self.new((4 * D, D), self.adt)
            qkv_w = [self.p(ap + f"attention.{n}.weight").detach() for n in ("query", "key", "value")]
            for j, pw in enumerate(qkv_w):
                jobs.append((w["wqkv"][j * D:(j + 1) * D], pw.view(D, D, 1, 1), ident, (), None))
                if training:
                    jobs.append((w["wqkvT"][:, j * D:(j + 1) * D], pw.view(D, D, 1, 1), transp, (), (3 * D, 1, 1, 1)))
            for key, name, n_out, n_in in (("wo", ap + "output.dense.weight", D, D), ("w1", lp + "mlp.fc1.weight", 4 * D, D),
                                           ("w2", lp + "mlp.fc2.weight", D, 4 * D)):
                pw = self.p(name).detach().view(n_out, n_in, 1, 1)
                jobs.append((w[key], pw, ident, (), None))
                if training:
                    jobs.append((w[key + "T"], pw, transp, (), None))
            cats.append((w["bqkv"], [self.p(ap + f"attention.{n}.bias") for n in ("query", "key", "value")]))
        if not jobs:
            return

        def cat_biases():
            with torch.no_grad():
                for dst, parts in cats:
                    torch.cat([t.detach() for t in parts], out=dst)
        self.be.pack_weights(jobs)
        self.be.host("cat_qkv_bias", cat_biases)

    def tile(self, which, M):
        """Tile shape of the backbone GEMMs (tools/gemm_tune.py, B200).  K = D = 384 (ViT-S): single-CTA 128-row tiles,
        192 columns for fc1 (fewer waves at N = 1536); the CTA-pair kernel is 5-20 % slower there (6 k-blocks per tile do
        not amortise the cross-CTA handshakes).  K >= 768 (ViT-B / L): 256 x 256 CTA-pair tiles (`cta_group::2`, each
        CTA stages half of the weight tile) -- fc2 69.0 vs 79.8 us, fc1 68.2 vs 72.5 us at ViT-B, M = 16448; the pair
        kernel halves the weight traffic from L2, which is what bounds the long-K shapes.  DP_PAIR_WIDE=0: off;
        DP_PAIR_QKV=0 keeps qkv / proj on the single-CTA kernel."""
        D = self.D
        wide = D >= 768 and M >= 4096 and bool(int(os.environ.get("DP_PAIR_WIDE", "1")))
        if which in ("fc1", "fc2"):
            if wide:
                return dict(block_n=256, cta_pair=1)
            return dict(block_n=192 if (which == "fc1" and (4 * D) % 192 == 0) else 0)
        if wide and bool(int(os.environ.get("DP_PAIR_QKV", "1"))):
            return dict(block_n=256, cta_pair=1)
        return {}

    def head_tile(self, M, N, K):
        """Tile shape of the heads' implicit-GEMM convolutions (forward and input gradients).  DP_HEAD_TILE: "pair" = CTA-pair
        256-row tiles (cta_group::2) where the shape allows, "wide" = single-CTA tiles 256 columns wide, "default" = 128 x
        128.  The long-K convolutions (K = 9 C) are bound by L2 -> SM operand traffic at 128 x 128 (64 FLOP/B)."""
        mode = os.environ.get("DP_HEAD_TILE", "pair")
        kmin = int(os.environ.get("DP_HEAD_TILE_KMIN", "1024"))
        if mode == "default" or M < 4096 or K < kmin:
            return {}
        if mode == "pair":
            for bn in (256, 192, 128):
                if N % bn == 0:
                    return dict(block_n=bn, cta_pair=1)
            return {}
        if mode == "wide" and N % 256 == 0:
            return dict(block_n=256)
        return {}

    # ------------------------------------------------------------------ plans
    def get_plan(self, B, H, W, training, scope="model"):
        """training: False (inference program), True (training step), or FROZEN_HEADS = 2: a training step whose pose
        heads are in eval mode -- `model.train(); model.pose_heads.eval()` in the reference's terms: BatchNorm
        normalises with the running statistics and leaves them alone, the z-head's Dropout is off, gradients still
        flow to every trainable tensor."""
        mode = int(training)
        key = (B, H, W, mode) if scope == "model" else (B, H, W, mode, scope)
        if key not in self.plans:
            if scope in ("model", "backbone"):
                plan = self.build_plan(B, H, W, mode > 0, scope, bn_frozen=mode == FROZEN_HEADS)
            else:
                plan = self.build_scope_plan(scope, B, H // PATCH, mode > 0, bn_frozen=mode == FROZEN_HEADS)
            self.plans[key] = plan
        return self.plans[key]

    # layers of the head table that belong to each stand-alone module scope (reference model/pose_heads.py)
    HOURGLASS_KEYS = ("skip", "dw", "pw", "down1", "down2", "bt1", "bt2", "up1", "up2")

    def scope_layers(self, scope, g):
        Ls = self.build_head_layers(g)
        if scope == "hourglass":
            return {k: Ls[k] for k in Ls if k in self.HOURGLASS_KEYS}
        if scope == "z_head":
            return {}
        return Ls

    def build_scope_plan(self, scope, B, g, training, bn_frozen=False):
        """Plan for a stand-alone sub-module of the heads (reference model/pose_heads.py:268-285 HourglassModule,
        :345-361 SpatialAwareHeatmapHead, :161-162 ZCoordinateHead, :395-400 SpatialAwarePoseHeads): the same recorded
        launches as inside the full model, starting from the module's own input buffer."""
        be = self.be
        if self.seed is None:
            self.seed = torch.zeros(1, dtype=torch.int64, device=self.device)
        N = g * g
        plan = {"B": B, "H": g * PATCH, "W": g * PATCH, "training": training, "g": g, "N": N, "T": N + 1, "M": B * (N + 1),
                "scope": scope, "lw": {}, "saved": {}, "bn_frozen": bool(bn_frozen and training)}
        t = plan["t"] = {}
        if scope in ("pose_heads", "heatmap_head"):
            t["feat"] = self.new((B * N, self.D), self.adt)        # NHWC rows, what the final LayerNorm writes in the model
        elif scope == "hourglass":
            t["hg_in"] = self.new((B * N, 512), self.adt)
        elif scope == "z_head":
            t["zin"] = self.new((B, self.D), F32)
        else:
            raise ValueError(scope)
        prog_f = be.begin()
        plan["layers"] = self.scope_layers(scope, g)
        be.fork()
        if plan["layers"]:
            self.record_pack_heads(plan)
        be.side(False)
        self.record_heads_forward(plan)
        plan["fwd"] = prog_f
        if training:
            plan["bwd"] = be.begin()
            self.record_backward(plan)
        return plan

    def build_plan(self, B, H, W, training, scope="model", bn_frozen=False):
        if H % PATCH or W % PATCH or H != W:
            raise ValueError(f"pixel_values must be square with sides a multiple of {PATCH} "
                             f"(reference model/dinov2_pose.py:151 assumes H = W = sqrt(N)); got {H}x{W}")
        if self.frozen is None:
            self.pack_frozen()
        if self.seed is None:
            self.seed = torch.zeros(1, dtype=torch.int64, device=self.device)
        be = self.be
        D, L, heads = self.D, self.L, self.heads
        g = H // PATCH
        N = g * g
        T = N + 1
        M = B * T
        plan = {"B": B, "H": H, "W": W, "training": training, "g": g, "N": N, "T": T, "M": M, "scope": scope,
                "bn_frozen": bool(bn_frozen and training)}
        fz = self.frozen
        t = plan["t"] = {}
        t["px"] = self.new((B, 3, H, W), F32)
        t["acol"] = self.new((B * N, PATCH_KP), self.adt)
        t["x"] = self.new((M, D), F32)
        t["xn"] = self.new((M, D), self.adt)
        t["qkv"] = self.new((M, 3 * D), self.adt)
        t["ctx"] = self.new((M, D), self.adt)
        t["h"] = self.new((M, 4 * D), self.adt)
        t["feat"] = self.new((B * N, D), self.adt)
        pos = self.pos_tables(H, W)
        use_lora = self.lora is not None
        lora_train = use_lora and training
        if training:
            t["x_mid"] = self.new((M, D), F32)     # residual stream after the last block's attention branch
            t["x_last"] = self.new((M, D), F32)    # residual stream entering the final LayerNorm
            t["pre"] = self.new((M, 4 * D), self.adt)  # gelu'(fc1 pre-activation) of the last block: the fc2 input gradient's multiplier
        if lora_train:
            t["y"] = self.new((M, D), F32)
            t["u"] = self.new((M, self.lora["rank"]), F32)

        # ---------------- forward program
        prog_f = be.begin()
        # un-frozen encoder layers: their bf16 weight copies are refreshed by the program itself
        self.record_pack_layers(plan, training)
        lw = plan["lw"]
        # the bf16 re-packing of the trainable head weights only has to be done before the first head convolution: it
        # runs on the second stream underneath the backbone (gather-bound, ~50 us, a fraction of the SMs)
        plan["layers"] = self.build_head_layers(g) if scope == "model" else {}
        be.fork()
        if scope == "model":
            self.record_pack_heads(plan)
        be.side(False)
        be.patch_im2col(t["px"], t["acol"], B=B, H=H, W=W, Kp=PATCH_KP)
        be.fill_cls(t["x"], pos[0], B=B, T=T, D=D)
        be.gemm(t["acol"], fz["wpe"], t["x"], M=B * N, N=D, K=PATCH_KP, out_dtype="f32", residual=pos,
                row_map="patch_tokens", map_a=N, map_b=T, name="patch_embed")
        scale = 1.0 / math.sqrt(D // heads)
        sv = plan["saved"] = {}      # per un-frozen layer: everything its backward reads
        # DP_SPLIT_BATCH=1 switches the two-stream backbone on (A/B); DP_SPLIT_MIN_BATCH lowers the batch threshold (tests)
        # Measured (gpurun_out/ab3): ViT-S step 3.974 -> 3.944 ms, ViT-B step and ViT-L inference 1-5 % SLOWER (half-size GEMMs
        # lose more to wave quantisation than the overlap returns) -> off by default
        split = bool(int(os.environ.get("DP_SPLIT_BATCH", "0"))) and B >= int(os.environ.get("DP_SPLIT_MIN_BATCH", "8"))
        B0 = B // 2
        halves = ((0, B0 * T, B0), (B0 * T, M, B - B0))
        split_open = False
        x_cur = t["x"]               # residual stream entering the current layer
        # LayerNorm fused into the producing projection's epilogue (row-owning kernel, gemm_rowln.cu): N = D <= 384 (ViT-S;
        # D = 768 / 1024 exceed the 512 accumulator columns of one SM).  OPT-IN (DP_FUSE_LN=1): measured on the benchmark
        # step it is SLOWER, 3.97 vs 3.87 ms (gpurun_out/rowln_bench*.log; kernel level 29.5 vs 17.9 + 6.3 us for proj + LN):
        # one tile per CTA means the 63 MB residual-stream epilogue of all 129 CTAs lands in one HBM burst with nothing to
        # overlap, and a fifth of the SMs idle in the single wave.  The 25 LayerNorm launches stay.
        fuse_ln = D in (128, 256, 384) and bool(int(os.environ.get("DP_FUSE_LN", "0")))
        ln1_done = ln2_done = False
        for i in range(L):
            lp = f"backbone.encoder.layer.{i}."
            last = i == L - 1
            if i in lw and split_open:
                be.sync("main_wait")
                split_open = False
            if i in lw:
                # ---- un-frozen layer (reference model/dinov2_pose.py:25-39): same arithmetic, weights from the per-step
                # packed copies; in training every intermediate the backward needs gets its own buffer
                w = lw[i]
                if training:
                    s_ = sv[i] = {"x_in": x_cur, "xn1": self.new((M, D), self.adt), "qkv": self.new((M, 3 * D), self.adt),
                                  "ctx": self.new((M, D), self.adt), "a": self.new((M, D), self.adt),
                                  "x_mid": self.new((M, D), F32), "xn2": self.new((M, D), self.adt),
                                  "pre": self.new((M, 4 * D), self.adt), "h": self.new((M, 4 * D), self.adt),
                                  "m": self.new((M, D), self.adt),
                                  "x_out": t["x_last"] if last else self.new((M, D), F32)}
                else:
                    s_ = {"x_in": x_cur, "xn1": t["xn"], "qkv": t["qkv"], "ctx": t["ctx"], "a": None, "x_mid": x_cur,
                          "xn2": t["xn"], "pre": None, "h": t["h"], "m": None, "x_out": x_cur}
                be.layernorm_fwd(x_cur, self.p(lp + "norm1.weight"), self.p(lp + "norm1.bias"), s_["xn1"], None, rows=M, D=D,
                                 eps=LN_EPS)
                be.gemm(s_["xn1"], w["wqkv"], s_["qkv"], M=M, N=3 * D, K=D, bias=w["bqkv"], name=f"qkv{i}")
                be.attention_fwd(s_["qkv"], s_["ctx"], B=B, T=T, heads=heads, scale=scale)
                be.gemm(s_["ctx"], w["wo"], s_["x_mid"], M=M, N=D, K=D, bias=self.p(self.att_prefix[i] + "output.dense.bias"),
                        out_dtype="f32", ls=self.p(lp + "layer_scale1.lambda1"), residual=x_cur, aux_out=s_["a"], ld_aux=D,
                        name=f"proj{i}")
                be.layernorm_fwd(s_["x_mid"], self.p(lp + "norm2.weight"), self.p(lp + "norm2.bias"), s_["xn2"], None, rows=M,
                                 D=D, eps=LN_EPS)
                be.gemm(s_["xn2"], w["w1"], s_["h"], M=M, N=4 * D, K=D, bias=self.p(lp + "mlp.fc1.bias"), act="gelu",
                        aux_out=s_["pre"], ld_aux=4 * D, name=f"fc1_{i}", **self.tile("fc1", M))
                be.gemm(s_["h"], w["w2"], s_["x_out"], M=M, N=D, K=4 * D, bias=self.p(lp + "mlp.fc2.bias"), out_dtype="f32",
                        ls=self.p(lp + "layer_scale2.lambda1"), residual=s_["x_mid"], aux_out=s_["m"], ld_aux=D,
                        name=f"fc2_{i}", **(self.tile("fc2", M) if s_["m"] is None else {}))   # no pair variant with aux_out
                x_cur = s_["x_out"]
                continue
            if split and not (last and (use_lora or training)):
                # ---- frozen layer, two half-batches on two streams: the HBM-bound LayerNorms and the MUFU-bound attention
                # of one half overlap the tensor-bound GEMMs of the other (images are independent in the backbone)
                if not split_open:
                    be.sync("side_wait")          # second stream: wait for the embeddings
                    split_open = True
                for hb, (r0, r1, Bh) in enumerate(halves):
                    be.side(hb == 1)
                    xs_, xn_, qkv_, ctx_, h_ = (t[k][r0:r1] for k in ("x", "xn", "qkv", "ctx", "h"))
                    Mh = r1 - r0
                    be.layernorm_fwd(xs_, self.p(lp + "norm1.weight"), self.p(lp + "norm1.bias"), xn_, None, rows=Mh, D=D,
                                     eps=LN_EPS)
                    be.gemm(xn_, fz[f"wqkv{i}"], qkv_, M=Mh, N=3 * D, K=D, bias=fz[f"bqkv{i}"], name=f"qkv{i}")
                    be.attention_fwd(qkv_, ctx_, B=Bh, T=T, heads=heads, scale=scale)
                    be.gemm(ctx_, fz[f"wo{i}"], xs_, M=Mh, N=D, K=D, bias=fz[f"bo{i}"], out_dtype="f32",
                            ls=self.p(lp + "layer_scale1.lambda1"), residual=xs_, name=f"proj{i}")
                    be.layernorm_fwd(xs_, self.p(lp + "norm2.weight"), self.p(lp + "norm2.bias"), xn_, None, rows=Mh, D=D,
                                     eps=LN_EPS)
                    be.gemm(xn_, fz[f"w1{i}"], h_, M=Mh, N=4 * D, K=D, bias=self.p(lp + "mlp.fc1.bias"), act="gelu",
                            name=f"fc1_{i}", **self.tile("fc1", Mh))
                    be.gemm(h_, fz[f"w2{i}"], xs_, M=Mh, N=D, K=4 * D, bias=self.p(lp + "mlp.fc2.bias"), out_dtype="f32",
                            ls=self.p(lp + "layer_scale2.lambda1"), residual=xs_, name=f"fc2_{i}", **self.tile("fc2", Mh))
                be.side(False)
                continue
            if split_open:
                be.sync("main_wait")              # both halves done: the remaining layers run on the whole batch
                split_open = False
            if not ln1_done:
                be.layernorm_fwd(t["x"], self.p(lp + "norm1.weight"), self.p(lp + "norm1.bias"), t["xn"], None, rows=M, D=D,
                                 eps=LN_EPS)
            ln1_done = False
            be.gemm(t["xn"], fz[f"wqkv{i}"], t["qkv"], M=M, N=3 * D, K=D, bias=fz[f"bqkv{i}"], name=f"qkv{i}",
                    **self.tile("qkv", M))
            be.attention_fwd(t["qkv"], t["ctx"], B=B, T=T, heads=heads, scale=scale)
            x_in = t["x"]
            x_att = t["x_mid"] if (last and training) else t["x"]
            if last and use_lora:
                lora_fused = (lora_train and D in (128, 256, 384) and self.lora["rank"] == 8
                              and bool(int(os.environ.get("DP_LORA_FUSED", "1"))))
                if lora_fused:
                    # the adapter INSIDE the projection GEMM (row-owning tcgen05 kernel, gemm_rowln.cu MODE 1): the rank-8 side
                    # product, the dropout mask, LayerScale and the residual are applied to the accumulator tile in tensor
                    # memory; y and u are saved for dp_lora_bwd (reference model/lora.py:26-28,53-59, dinov2_pose.py:197-204)
                    be.gemm(t["ctx"], fz[f"wo{i}"], x_att, M=M, N=D, K=D, bias=fz[f"bo{i}"], out_dtype="f32",
                            ls=self.p(lp + "layer_scale1.lambda1"), residual=x_in,
                            lora=dict(A=self.p(self.lora_prefix + "lora_A"), B=self.p(self.lora_prefix + "lora_B"),
                                      scaling=self.lora["alpha"] / self.lora["rank"],
                                      p_drop=float(self.lora.get("dropout", 0.0)), seed=self.seed, y_out=t["y"], u_out=t["u"]),
                            name="proj_last_lora")
                elif lora_train:
                    # explicit adapter (dropout + saved activations for the backward): ViT-B / L (D > 384) or rank != 8
                    be.gemm(t["ctx"], fz[f"wo{i}"], t["y"], M=M, N=D, K=D, bias=fz[f"bo{i}"], out_dtype="f32",
                            name="proj_last")
                    be.lora_fwd(t["y"], self.p(self.lora_prefix + "lora_A"), self.p(self.lora_prefix + "lora_B"),
                                self.p(lp + "layer_scale1.lambda1"), x_in, x_att, t["u"], rows=M, D=D,
                                R=self.lora["rank"], scaling=self.lora["alpha"] / self.lora["rank"],
                                p_drop=float(self.lora.get("dropout", 0.0)), seed=self.seed)
                else:
                    # eval: fold the adapter into the projection, W' = (I + s A B)^T W  (dropout inactive)
                    t["wo_m"] = self.new((D, D), self.adt)
                    t["bo_m"] = self.new((D,), F32)
                    be.host("merge_lora", self._merge_lora_fn(t["wo_m"], t["bo_m"], i))
                    be.gemm(t["ctx"], t["wo_m"], x_att, M=M, N=D, K=D, bias=t["bo_m"], out_dtype="f32",
                            ls=self.p(lp + "layer_scale1.lambda1"), residual=x_in, name="proj_last_merged")
            else:
                # row-owning projection with the sub-block's NEXT LayerNorm fused into its epilogue (gemm_rowln.cu)
                ln2 = dict(ln=dict(gamma=self.p(lp + "norm2.weight"), beta=self.p(lp + "norm2.bias"), out=t["xn"],
                                   eps=LN_EPS)) if fuse_ln else {}
                be.gemm(t["ctx"], fz[f"wo{i}"], x_att, M=M, N=D, K=D, bias=fz[f"bo{i}"], out_dtype="f32",
                        ls=self.p(lp + "layer_scale1.lambda1"), residual=x_in, name=f"proj{i}",
                        **(ln2 if fuse_ln else self.tile("proj", M)))
                ln2_done = fuse_ln
            if not ln2_done:
                be.layernorm_fwd(x_att, self.p(lp + "norm2.weight"), self.p(lp + "norm2.bias"), t["xn"], None, rows=M, D=D,
                                 eps=LN_EPS)
            ln2_done = False
            be.gemm(t["xn"], fz[f"w1{i}"], t["h"], M=M, N=4 * D, K=D, bias=self.p(lp + "mlp.fc1.bias"), act="gelu",
                    aux_out=t["pre"] if (last and training) else None, ld_aux=4 * D, name=f"fc1_{i}",
                    **self.tile("fc1", M))
            x_out = t["x_last"] if (last and training) else t["x"]
            # fc2 + LayerScale + residual, and -- when the next layer is a plain frozen one -- ITS norm1 in the same epilogue
            nxt = f"backbone.encoder.layer.{i + 1}."
            fuse_next = fuse_ln and not last and (i + 1) not in lw and not split
            ln1 = dict(ln=dict(gamma=self.p(nxt + "norm1.weight"), beta=self.p(nxt + "norm1.bias"), out=t["xn"],
                               eps=LN_EPS)) if fuse_next else {}
            be.gemm(t["h"], fz[f"w2{i}"], x_out, M=M, N=D, K=4 * D, bias=self.p(lp + "mlp.fc2.bias"), out_dtype="f32",
                    ls=self.p(lp + "layer_scale2.lambda1"), residual=x_att, name=f"fc2_{i}",
                    **(ln1 if fuse_next else self.tile("fc2", M)))
            ln1_done = fuse_next
        if split_open:
            be.sync("main_wait")
        x_fin = t["x_last"] if training else t["x"]
        plan["x_final"] = x_fin
        if scope == "backbone":
            # stand-alone Dinov2Model.forward (HF:473-478): last_hidden_state = LayerNorm(all tokens), fp32
            be.join()
            t["lhs"] = self.new((M, D), F32)
            be.layernorm_fwd(x_fin, self.p("backbone.layernorm.weight"), self.p("backbone.layernorm.bias"), None, t["lhs"],
                             rows=M, D=D, eps=LN_EPS)
            plan["fwd"] = prog_f
            return plan
        be.layernorm_fwd(x_fin, self.p("backbone.layernorm.weight"), self.p("backbone.layernorm.bias"), t["feat"], None,
                         rows=M, D=D, T=T, drop_cls=True, eps=LN_EPS)
        self.record_heads_forward(plan)
        plan["fwd"] = prog_f
        if training:
            plan["bwd"] = be.begin()
            self.record_backward(plan)
        return plan

    def _merge_lora_fn(self, wo_m, bo_m, i):
        def fn():
            with torch.no_grad():
                A = self.p(self.lora_prefix + "lora_A").detach().float()
                Bm = self.p(self.lora_prefix + "lora_B").detach().float()
                s = self.lora["alpha"] / self.lora["rank"]
                wo = self.p(self.att_prefix[i] + "output.dense.weight").detach().float()
                bo = self.p(self.att_prefix[i] + "output.dense.bias").detach().float()
                mix = torch.eye(self.D, device=wo.device) + s * (A @ Bm)      # y' = y @ mix
                wo_m.copy_(mix.t() @ wo)
                bo_m.copy_(bo @ mix)
        return fn

    # ------------------------------------------------------------------ heads forward
    def _bn_tensors(self, L):
        c = L.cout
        for n in ("scale", "shift", "mean", "invstd"):
            L.t[n] = self.new((c,), F32)
        # fp64 statistic accumulators: [:2c] for the forward (fused GEMM statistics / bn_stats -> bn_finalize),
        # 8 replicas of 2c for the backward reduce (DP_BN_BWD_REPLICAS) + one 2c block of coefficient scratch
        L.t["sums"] = self.new((2 * c * 9,), torch.float64)

    def _conv_forward(self, L, x, NB, training, out_override=None, bn_frozen=False):
        """Records conv (no BN) of layer L on NHWC input x; returns raw (train) or activated (eval) output.
        x: 4-D NHWC view for implicit convs, 2-D [P, C] otherwise."""
        be = self.be
        P_out = NB * L.oh * L.ow
        fold = (not training) and L.bn is not None
        if fold:
            be_scale, be_shift = L.t["scale"], L.t["shift"]
            act = "relu" if L.relu else "none"
        bias = self.p(L.name + ".bias")
        # train-mode BatchNorm statistics are accumulated by the producing GEMM's epilogue (no separate pass)
        st = dict(stats=L.t["sums"], stats_c=L.cout) if (training and L.bn is not None and not bn_frozen
                                                         and L.kind not in ("convT", "dw")) else {}
        L.t["stats_fused"] = bool(st)
        out = out_override if out_override is not None else self.new((P_out, L.cout), self.rdt if (training and L.bn is not None) else self.adt)
        kk = L.k * L.k
        if L.kind == "conv" and L.k == 1:
            be.gemm(x.reshape(-1, L.cin), L.t["wf"], out, M=P_out, N=L.cout, K=L.cin,
                    bias=be_shift if fold else bias, scale=be_scale if fold else None, act=act if fold else "none",
                    name=L.name, **st, **(self.head_tile(P_out, L.cout, L.cin) if (training and L.bn is not None) else {}))
        elif L.kind in ("conv", "convT_s1"):
            pad = L.pad if L.kind == "conv" else L.k - 1 - L.pad
            xin = x.view(NB, L.ih, L.iw, L.cin) if x.dim() == 2 else x
            be.gemm(xin, L.t["wf"], out, M=P_out, N=L.cout, K=kk * L.cin, bias=be_shift if fold else bias,
                    scale=be_scale if fold else None, act=act if fold else "none",
                    conv=dict(KH=L.k, KW=L.k, pad=pad, OH=L.oh, OW=L.ow), name=L.name, **st,
                    **(self.head_tile(P_out, L.cout, kk * L.cin) if (training and L.bn is not None) else {}))
        elif L.kind == "conv_s2":
            L.t["col"] = self.new((P_out, kk * L.cin), self.adt)
            be.im2col(x, L.t["col"], NB=NB, IH=L.ih, IW=L.iw, C=L.cin, OH=L.oh, OW=L.ow, KH=L.k, KW=L.k, stride=L.stride,
                      pad=L.pad)
            be.gemm(L.t["col"], L.t["wf"], out, M=P_out, N=L.cout, K=kk * L.cin, bias=be_shift if fold else bias,
                    scale=be_scale if fold else None, act=act if fold else "none", name=L.name, **st,
                    **(self.head_tile(P_out, L.cout, kk * L.cin) if (training and L.bn is not None) else {}))
        elif L.kind == "convT2":
            P_in = NB * L.ih * L.iw
            L.t["bias4"] = self.new((kk * L.cout,), F32)
            if fold:
                L.t["scale4"] = self.new((kk * L.cout,), F32)
            be.host("bias4", self._tile4_fn(L, fold))
            be.gemm(x.reshape(P_in, L.cin), L.t["wf"], out, M=P_in, N=kk * L.cout, K=L.cin, bias=L.t["bias4"],
                    scale=L.t["scale4"] if fold else None, act=act if fold else "none", row_map="shuffle2x2",
                    map_a=L.cout, OH=L.ih, OW=L.iw, NB=NB, ldo=L.cout, name=L.name, **st)
        elif L.kind == "convT":
            P_in = NB * L.ih * L.iw
            L.t["colT"] = self.new((P_in, kk * L.cout), self.adt)
            be.gemm(x.reshape(P_in, L.cin), L.t["wf"], L.t["colT"], M=P_in, N=kk * L.cout, K=L.cin, name=L.name)
            raw = self.new((P_out, L.cout), self.adt) if fold else out
            be.col2im(L.t["colT"], bias, raw, NB=NB, SH=L.ih, SW=L.iw, C=L.cout, BH=L.oh, BW=L.ow, KH=L.k, KW=L.k,
                      stride=L.stride, pad=L.pad)
            if fold:
                L.t["shift_nb"] = self.new((L.cout,), F32)
                be.host("fold_nobias", self._fold_nobias_fn(L))
                be.bn_apply(raw, L.t["scale"], L.t["shift_nb"], None, None, out, P=P_out, C=L.cout, relu=L.relu)
        elif L.kind == "dw":
            raw = self.new((P_out, L.cout), self.adt) if fold else out
            be.dwconv3x3(x, self.p(L.name + ".weight"), bias, None, raw, NB=NB, H=L.ih, W=L.iw, C=L.cin)
            if fold:
                L.t["shift_nb"] = self.new((L.cout,), F32)
                be.host("fold_nobias", self._fold_nobias_fn(L))
                be.bn_apply(raw, L.t["scale"], L.t["shift_nb"], None, None, out, P=P_out, C=L.cout, relu=L.relu)
        else:
            raise ValueError(L.kind)
        return out

    def _tile4_fn(self, L, fold):
        def fn():
            with torch.no_grad():
                kk = L.k * L.k
                if fold:
                    L.t["bias4"].copy_(L.t["shift"].repeat(kk))
                    L.t["scale4"].copy_(L.t["scale"].repeat(kk))
                else:
                    L.t["bias4"].copy_(self.p(L.name + ".bias").detach().repeat(kk))
        return fn

    def _fold_nobias_fn(self, L):
        # col2im / dwconv already added the conv bias: remove it from the folded shift
        def fn():
            with torch.no_grad():
                L.t["shift_nb"].copy_(L.t["shift"] - self.p(L.name + ".bias").detach() * L.t["scale"])
        return fn

    def _bn_forward(self, L, raw, NB, add1=None, add2=None, mode=0, out=None, bn_frozen=False):
        """train-mode BatchNorm (+ReLU, + fused adds) on raw [P, C]."""
        be = self.be
        P = NB * L.oh * L.ow
        bn = L.bn
        if bn_frozen:
            # heads in eval mode inside a training step: running statistics, re-derived every step because gamma / beta
            # train; mean / invstd feed the backward where the batch statistics would
            be.bn_fold_eval(self.p(bn + ".weight"), self.p(bn + ".bias"), self.p(bn + ".running_mean"),
                            self.p(bn + ".running_var"), None, L.t["scale"], L.t["shift"], C=L.cout, mean=L.t["mean"],
                            invstd=L.t["invstd"])
            out = out if out is not None else self.new((P, L.cout), self.adt)
            be.bn_apply(raw, L.t["scale"], L.t["shift"], add1, add2, out, P=P, C=L.cout, relu=L.relu, mode=mode)
            return out
        if not L.t.get("stats_fused"):
            be.bn_stats(raw, L.t["sums"], P=P, C=L.cout)
        out = out if out is not None else self.new((P, L.cout), self.adt)
        if L.cout <= 512:
            be.bn_finalize_apply(raw, L.t["sums"], self.p(bn + ".weight"), self.p(bn + ".bias"), self.p(bn + ".running_mean"),
                                 self.p(bn + ".running_var"), L.t["scale"], L.t["shift"], L.t["mean"], L.t["invstd"],
                                 add1, add2, out, P=P, C=L.cout, relu=L.relu, mode=mode)
            return out
        be.bn_finalize(L.t["sums"], self.p(bn + ".weight"), self.p(bn + ".bias"), self.p(bn + ".running_mean"),
                       self.p(bn + ".running_var"), L.t["scale"], L.t["shift"], L.t["mean"], L.t["invstd"], C=L.cout,
                       count=P)
        be.bn_apply(raw, L.t["scale"], L.t["shift"], add1, add2, out, P=P, C=L.cout, relu=L.relu, mode=mode)
        return out

    def record_heads_forward(self, plan):
        be = self.be
        B, g, training = plan["B"], plan["g"], plan["training"]
        bn_frozen = bool(plan.get("bn_frozen"))
        t = plan["t"]
        Ls = plan["layers"]
        scope = plan.get("scope", "model")
        has_z = scope in ("model", "pose_heads", "z_head")
        has_hm = scope in ("model", "pose_heads", "heatmap_head")
        be.join()                # packed head weights (second stream, forked at the start of the program)
        for L in Ls.values():
            if L.bn is not None:
                self._bn_tensors(L)
                if not training:
                    bn = L.bn
                    be.bn_fold_eval(self.p(bn + ".weight"), self.p(bn + ".bias"), self.p(bn + ".running_mean"),
                                    self.p(bn + ".running_var"), self.p(L.name + ".bias"), L.t["scale"], L.t["shift"],
                                    C=L.cout)
        K = self.K
        # The z head (5 small launches on a [B, D] matrix) only needs the final LayerNorm output: it runs on the second
        # stream underneath the heat-map head and is joined at the end of the program.
        be.fork()
        if has_z:
            # z head: mean over patch tokens -> MLP (pose_heads.py:397-398, :148-159)
            zh = self.cfg["z_hidden"]
            dims = [self.D] + list(zh) + [K]
            if scope != "z_head":    # stand-alone ZCoordinateHead: the [B, D] feature vector IS the input
                t["zin"] = self.new((B, self.D), F32)
                be.mean_tokens(t["feat"], t["zin"], B=B, N=plan["N"], D=self.D)
            zp = "pose_heads.z_head.mlp."
            cur = t["zin"]
            p_drop = float(self.cfg.get("z_dropout", 0.0)) if (training and not bn_frozen) else 0.0   # heads.eval(): Dropout off
            t["zact"] = [cur]
            for j in range(len(dims) - 1):
                lastl = j == len(dims) - 2
                out = self.new((B, dims[j + 1]), F32)
                be.sgemm_small(cur, dims[j], 1, self.p(zp + f"{3 * j}.weight"), 1, dims[j], out, dims[j + 1], M=B,
                               N=dims[j + 1], K=dims[j], bias=self.p(zp + f"{3 * j}.bias"), relu=not lastl,
                               p_drop=0.0 if lastl else p_drop, seed=self.seed if (p_drop > 0 and not lastl) else None)
                cur = out
                t["zact"].append(cur)
            t["z"] = cur
            plan["zdims"] = dims
        be.side(False)
        a = plan["a"] = {}    # activations (bf16 [P, C])
        r = plan["raw"] = {}  # pre-BN conv outputs (training only)
        if scope == "z_head":
            be.join()
            return
        feat4 = t["feat"].view(B, g, g, self.D) if has_hm else None

        def unit(key, x, **kw):
            L = Ls[key]
            if training:
                r[key] = self._conv_forward(L, x, B, True, bn_frozen=bn_frozen)
                a[key] = self._bn_forward(L, r[key], B, bn_frozen=bn_frozen, **kw)
            else:
                a[key] = self._conv_forward(L, x, B, False)
            return a[key]

        if has_hm:
            a1 = unit("fr0", feat4)
        else:                       # stand-alone HourglassModule: its input takes the place of feature_refine's first unit
            a1 = a["fr0"] = t["hg_in"]
        a1_4 = a1.view(B, g, g, 512)
        # The hourglass has three independent branches (pose_heads.py:268-285).  The down/up branch is a chain of small
        # launches on 8x8 / 4x4 maps (a handful of CTAs each, latency bound): it runs on side stream 2 underneath the
        # skip and depthwise branches and is joined before the 3-way sum.  DP_HG_STREAMS=0 keeps it in line (A/B).
        hg2 = bool(int(os.environ.get("DP_HG_STREAMS", "1")))
        plan["hg2"] = hg2
        if training:
            if hg2:
                be.fork(2)
            unit("down1", a1_4)
            unit("down2", a["down1"].view(B, g // 2, g // 2, 256))
            unit("bt1", a["down2"].view(B, g // 4, g // 4, 128))
            unit("bt2", a["bt1"].view(B, g // 4, g // 4, 128), add1=a["down2"], mode=1)
            unit("up1", a["bt2"])
            be.side(False)
            unit("skip", a1)
            unit("dw", a1_4)
            unit("pw", a["dw"])
            if hg2:
                be.join(2)
            # hourglass output = up2 + skip + depthwise branch (pose_heads.py:285), fused into up2's BN apply
            unit("up2", a["up1"], add1=a["skip"], add2=a["pw"])
            hgout = a["up2"]
        else:
            if hg2:
                be.fork(2)
            unit("down1", a1_4)
            unit("down2", a["down1"].view(B, g // 2, g // 2, 256))
            unit("bt1", a["down2"].view(B, g // 4, g // 4, 128))
            # bottleneck residual + relu, 3-way sum: small element-wise passes through bn_apply with identity BN
            L = Ls["bt2"]
            L.t["one"] = self.new((L.cout,), F32, 1.0)
            L.t["zero"] = self.new((L.cout,), F32)
            bt2 = unit("bt2", a["bt1"].view(B, g // 4, g // 4, 128))
            a["bt2r"] = self.new(tuple(bt2.shape), self.adt)
            be.bn_apply(bt2, L.t["one"], L.t["zero"], a["down2"], None, a["bt2r"], P=bt2.shape[0], C=L.cout, relu=True, mode=1)
            unit("up1", a["bt2r"])
            up2 = unit("up2", a["up1"])
            be.side(False)
            unit("skip", a1)
            unit("dw", a1_4)
            unit("pw", a["dw"])
            if hg2:
                be.join(2)
            L2 = Ls["up2"]
            L2.t["one"] = self.new((L2.cout,), F32, 1.0)
            L2.t["zero"] = self.new((L2.cout,), F32)
            hgout = a["hg"] = self.new(tuple(up2.shape), self.adt)
            be.bn_apply(up2, L2.t["one"], L2.t["zero"], a["skip"], a["pw"], hgout, P=up2.shape[0], C=L2.cout, relu=False, mode=0)
        plan["hgout"] = hgout
        if not has_hm:
            be.join()
            return
        unit("fr4", hgout.view(B, g, g, 512))
        unit("ups0", a["fr4"])
        s47, s48 = Ls["ups0"].oh, Ls["ups1"].oh
        unit("ups1", a["ups0"].view(B, s47, s47, 128))
        unit("pred0", a["ups1"].view(B, s48, s48, 128))
        # prediction.3: 1x1 conv 64 -> K with bias, written straight to fp32 NCHW
        L = Ls["pred3"]
        K = self.K
        t["hm_full"] = self.new((B, K, s48, s48), F32)
        # 0.45 GFLOP on 33 MB: a CUDA-core kernel with the fp32 parameter (csrc/pred_ops.cu) instead of 1152 one-k-block
        # tensor-core tiles.  DP_PRED_SIMT=0 keeps the GEMM (A/B); K > 32 has no SIMT variant.
        plan["pred_simt"] = K <= 32 and L.cin == 64 and bool(int(os.environ.get("DP_PRED_SIMT", "1")))
        if plan["pred_simt"]:
            be.pred1x1_fwd(a["pred0"], self.p(L.name + ".weight"), self.p(L.name + ".bias"), t["hm_full"], P=B * s48 * s48,
                           HW=s48 * s48, C=64, K=K)
        else:
            be.gemm(a["pred0"], L.t["wf"], t["hm_full"], M=B * s48 * s48, N=K, K=64, bias=self.p(L.name + ".bias"),
                    out_dtype="f32", row_map="nchw", n_valid=K, map_a=K, OH=s48, OW=s48, NB=B, block_n=32, name="pred3")
        if s48 == self.hm_size:
            t["hm"] = t["hm_full"]       # bilinear resize to the same size is the identity (pose_heads.py:353-359)
        elif s48 == 2 * self.hm_size:
            t["hm"] = self.new((B, K, self.hm_size, self.hm_size), F32)
            be.avgpool2(t["hm_full"], t["hm"], planes=B * K, OH=self.hm_size, OW=self.hm_size)
        else:
            raise NotImplementedError(f"heat-map resize {s48} -> {self.hm_size} (only 1x and 2x reductions occur "
                                      "for 224^2 / 448^2 inputs)")
        be.join()                # z head (second stream)
        if training and not bn_frozen:
            # torch BatchNorm bookkeeping (momentum is fixed, the value is unused): one launch for the 14 counters
            counters = [self.Bufs[L.bn + ".num_batches_tracked"] for L in Ls.values() if L.bn is not None]
            be.add_i64(counters, 1)

    # ------------------------------------------------------------------ backward
    def record_backward(self, plan):
        be = self.be
        B, g, N, T, M = plan["B"], plan["g"], plan["N"], plan["T"], plan["M"]
        D, K = self.D, self.K
        t, a, r, Ls = plan["t"], plan["a"], plan["raw"], plan["layers"]
        scope = plan.get("scope", "model")
        has_z = scope in ("model", "pose_heads", "z_head")
        has_hm = scope in ("model", "pose_heads", "heatmap_head")
        lay = self.layout()
        flat = plan["gflat"] = self.new((lay["total"],), F32)
        G = plan["grads"] = {}
        for n in lay["names"]:
            off, k = lay["offsets"][n]
            G[n] = flat[off:off + k].view(self.P[n].shape)

        # Weight-gradient GEMMs depend only on (draw, saved activation) and nothing downstream needs them before the
        # gradient buffer is handed over: they go to the second stream (opened below for the z head) and overlap the
        # HBM-bound BatchNorm-backward / col2im kernels of the following layers on the main stream.  DP_BWD_OVERLAP=0
        # keeps them in line (A/B switch).
        overlap = bool(int(os.environ.get("DP_BWD_OVERLAP", "1")))

        chain = {"ws": None}     # set while the hourglass' down/up branch is being recorded on side stream 2

        def wg(*a_, **k_):
            if chain["ws"] is not None:
                k_["workspace"] = chain["ws"]
                be.wgrad(*a_, **k_)
            elif overlap:
                be.sync("side_wait")
                be.side(True)
                be.wgrad(*a_, **k_)
                be.side(False)
            else:
                be.wgrad(*a_, **k_)

        def done(key):
            # every gradient in flat[:group_end[key]] is final from here on: the second stream (z-head chain, weight
            # gradients) has caught up
            be.sync("main_wait")
            be.mark(("grads_final", lay["group_end"][key]))
        if has_hm:
            t["dhm"] = self.new(tuple(t["hm"].shape), F32)
        if has_z:
            t["dz"] = self.new((B, K), F32)
        # split-K workspace shared by all weight-gradient launches (they run one after the other): partial tiles are
        # stored without atomics and reduced by a second kernel (deterministic, no contention on the gradient buffer)
        ws = t["wgrad_ws"] = self.new((16 << 20,), F32)
        be.host("zero_grads", flat.zero_)
        # the z-head backward (12 small launches, M = batch) depends only on dz: second stream, under the heat-map head
        be.fork()
        dims = plan.get("zdims", [0])
        zp = "pose_heads.z_head.mlp."
        p_drop = 0.0 if plan.get("bn_frozen") else float(self.cfg.get("z_dropout", 0.0))
        dcur = t.get("dz")
        nl = len(dims) - 1
        for j in reversed(range(nl if has_z else 0)):
            xin, yout = t["zact"][j], t["zact"][j + 1]
            if j < nl - 1:
                # through dropout + relu of layer j: mask by the saved (post-dropout) activation
                dm = self.new((B, dims[j + 1]), F32)
                be.relu_mask(dcur, yout, dm, n=B * dims[j + 1], keep_scale=1.0 / (1.0 - p_drop) if p_drop > 0 else 1.0)
                dcur = dm
            be.sgemm_small(dcur, 1, dims[j + 1], xin, dims[j], 1, G[zp + f"{3 * j}.weight"], dims[j], M=dims[j + 1],
                           N=dims[j], K=B)
            be.colsum(dcur, G[zp + f"{3 * j}.bias"], P=B, C=dims[j + 1], ld=dims[j + 1])
            dx = self.new((B, dims[j]), F32)
            be.sgemm_small(dcur, dims[j + 1], 1, self.p(zp + f"{3 * j}.weight"), dims[j], 1, dx, dims[j], M=B, N=dims[j],
                           K=dims[j + 1])
            dcur = dx
        be.side(False)
        if scope == "z_head":
            be.join()
            plan["d_in"] = dcur            # gradient w.r.t. the [B, D] feature vector
            be.mark(("grads_final", lay["total"]))
            return

        def bn_bwd(key, dact, add1=None, mode=0, dres=None, shuffle=False):
            L = Ls[key]
            P = B * L.oh * L.ow
            bn = L.bn
            be.bn_bwd_reduce(dact, r[key], add1, L.t["scale"], L.t["shift"], L.t["mean"], L.t["invstd"], L.t["sums"], P=P,
                             C=L.cout, relu=L.relu, mode=mode)
            draw = self.new((P, L.cout), self.adt)
            be.bn_bwd_apply(dact, r[key], add1, self.p(bn + ".weight"), L.t["scale"], L.t["shift"], L.t["mean"],
                            L.t["invstd"], L.t["sums"], draw, dres, G[bn + ".weight"], G[bn + ".bias"], P=P, C=L.cout,
                            relu=L.relu, mode=mode, eval_mode=2 if plan.get("bn_frozen") else 0,
                            shuffle_oh=L.oh if shuffle else 0, shuffle_ow=L.ow if shuffle else 0)
            if plan.get("bn_frozen") and (L.name + ".bias") in G:
                # frozen statistics: the conv bias is no longer cancelled by the batch mean.  y = scale * (conv + b) + shift,
                # so db = sum draw = scale * sum dy = scale * dbeta (one [C] product; this mode is not the benchmarked one)
                gb, gbeta, sc = G[L.name + ".bias"], G[bn + ".bias"], L.t["scale"]
                be.host("conv_bias_grad", lambda gb=gb, gbeta=gbeta, sc=sc: torch.mul(sc, gbeta, out=gb))
            return draw

        def conv_bwd(key, draw, x_in, want_dx=True, dx_residual=None):
            """weight gradient of layer `key` (+ input gradient).  x_in = the layer's forward input."""
            L = Ls[key]
            k, ci, co = L.k, L.cin, L.cout
            kk = k * k
            gw = G[L.name + ".weight"]
            P_out = B * L.oh * L.ow
            P_in = B * L.ih * L.iw
            dx = None
            if L.kind == "conv" and k == 1:
                wg(draw, x_in.reshape(P_in, ci), gw, Mc=co, Nc=ci, so_m=ci, so_n=1, P=P_out, name=L.name + ".wgrad",
                         workspace=ws)
                if want_dx:
                    dx = self.new((P_in, ci), self.adt)
                    be.gemm(draw, L.t["wd"], dx, M=P_out, N=ci, K=draw.shape[1], residual=dx_residual,
                            name=L.name + ".dgrad", **self.head_tile(P_out, ci, draw.shape[1]))
            elif L.kind == "conv":
                d4 = draw.view(B, L.oh, L.ow, co)
                x4 = x_in.view(B, L.ih, L.iw, ci) if x_in.dim() == 2 else x_in
                if co <= 64 < ci and (L.oh, L.ow) == (L.ih, L.iw):
                    # few output channels (prediction.0: 128 -> 64): put the INPUT channels on the 128-row MMA side.
                    # dW[co,ci,ky,kx] = sum_q x[q,ci] * dRaw[q - (ky,kx) + pad, co]: the shift moves to the N operand
                    # with mirrored taps (tap' = kk-1-tap, pad' = k-1-pad), hence so_t = -1 from the last tap
                    wg(x4, d4, gw.view(-1)[kk - 1:], Mc=ci, Nc=co, so_m=kk, so_n=ci * kk, so_t=-1,
                             conv=dict(KH=k, KW=k, pad=k - 1 - L.pad), block_n=64, name=L.name + ".wgrad", workspace=ws)
                else:
                    wg(d4, x4, gw, Mc=co, Nc=ci, so_m=ci * kk, so_n=kk, so_t=1, conv=dict(KH=k, KW=k, pad=L.pad),
                             name=L.name + ".wgrad", workspace=ws)
                if want_dx:
                    dx = self.new((P_in, ci), self.adt)
                    be.gemm(d4, L.t["wd"], dx, M=P_in, N=ci, K=kk * co, residual=dx_residual,
                            conv=dict(KH=k, KW=k, pad=k - 1 - L.pad, OH=L.ih, OW=L.iw), name=L.name + ".dgrad",
                            **self.head_tile(P_in, ci, kk * co))
            elif L.kind == "convT_s1":
                x4 = x_in.view(B, L.ih, L.iw, ci)
                d4 = draw.view(B, L.oh, L.ow, co)
                wg(x4, d4, gw, Mc=ci, Nc=co, so_m=co * kk, so_n=kk, so_t=1, conv=dict(KH=k, KW=k, pad=L.pad),
                         name=L.name + ".wgrad", workspace=ws)
                if want_dx:
                    dx = self.new((P_in, ci), self.adt)
                    be.gemm(d4, L.t["wd"], dx, M=P_in, N=ci, K=kk * co, residual=dx_residual,
                            conv=dict(KH=k, KW=k, pad=L.pad, OH=L.ih, OW=L.iw), name=L.name + ".dgrad",
                            **self.head_tile(P_in, ci, kk * co))
            elif L.kind == "conv_s2":
                wg(draw, L.t["col"], gw, Mc=co, Nc=kk * ci, so_m=ci * kk, so_n=kk, so_no=1, n_inner=ci, P=P_out,
                         name=L.name + ".wgrad", workspace=ws)
                if want_dx:
                    dcol = self.new((P_out, kk * ci), self.adt)
                    be.gemm(draw, L.t["wd"], dcol, M=P_out, N=kk * ci, K=co, name=L.name + ".dgrad")
                    dx = self.new((P_in, ci), self.adt)
                    be.col2im(dcol, None, dx, NB=B, SH=L.oh, SW=L.ow, C=ci, BH=L.ih, BW=L.iw, KH=k, KW=k, stride=L.stride,
                              pad=L.pad)
            elif L.kind == "convT2":
                # draw arrives in the un-shuffled [P_in, 4*Cout] layout (bn_bwd(..., shuffle=True))
                dcol = draw.view(P_in, kk * co)
                wg(x_in.reshape(P_in, ci), dcol, gw, Mc=ci, Nc=kk * co, so_m=co * kk, so_n=kk, so_no=1, n_inner=co,
                         P=P_in, name=L.name + ".wgrad", workspace=ws)
                if want_dx:
                    dx = self.new((P_in, ci), self.adt)
                    be.gemm(dcol, L.t["wd"], dx, M=P_in, N=ci, K=kk * co, residual=dx_residual, name=L.name + ".dgrad")
            elif L.kind == "convT":
                dcol = self.new((P_in, kk * co), self.adt)
                be.im2col(draw, dcol, NB=B, IH=L.oh, IW=L.ow, C=co, OH=L.ih, OW=L.iw, KH=k, KW=k, stride=L.stride,
                          pad=L.pad)
                wg(x_in.reshape(P_in, ci), dcol, gw, Mc=ci, Nc=kk * co, so_m=co * kk, so_n=kk, so_no=1, n_inner=co,
                         P=P_in, name=L.name + ".wgrad", workspace=ws)
                if want_dx:
                    dx = self.new((P_in, ci), self.adt)
                    be.gemm(dcol, L.t["wd"], dx, M=P_in, N=ci, K=kk * co, residual=dx_residual, name=L.name + ".dgrad")
            # conv biases feeding train-mode BN have an identically-zero gradient (the batch mean removes any
            # constant); gflat was zeroed, nothing to write.
            return dx

        # ---- heat-map head
        def tail_backward():
            """prediction.3 .. feature_refine.4: heat-map gradient -> gradient w.r.t. the hourglass output"""
            s47, s48 = Ls["ups0"].oh, Ls["ups1"].oh
            P48 = B * s48 * s48
            up = s48 // self.hm_size
            L = Ls["pred3"]
            d = self.new((P48, 64), self.adt)
            if plan.get("pred_simt") and up == 1:
                # one launch: input gradient, weight gradient and bias gradient straight from the fp32 NCHW heat-map gradient
                be.pred1x1_bwd(t["dhm"], a["pred0"], self.p(L.name + ".weight"), d, G[L.name + ".weight"],
                               G[L.name + ".bias"], P=P48, HW=s48 * s48, C=64, K=K)
            else:
                t["ghm"] = self.new((P48, HM_PAD), self.adt)
                be.hm_grad_to_nhwc(t["dhm"], t["ghm"], NB=B, K=K, Kp=HM_PAD, OH=s48, OW=s48, up=up)
                be.colsum(t["ghm"], G[L.name + ".bias"], P=P48, C=K, ld=HM_PAD)
                wg(t["ghm"], a["pred0"], G[L.name + ".weight"], Mc=K, Nc=64, so_m=64, so_n=1, P=P48, block_n=64,
                   name="pred3.wgrad", workspace=ws)
                be.gemm(t["ghm"], L.t["wd"], d, M=P48, N=64, K=HM_PAD, name="pred3.dgrad")
            d = conv_bwd("pred0", bn_bwd("pred0", d), a["ups1"])
            d = conv_bwd("ups1", bn_bwd("ups1", d), a["ups0"])
            d = conv_bwd("ups0", bn_bwd("ups0", d), a["fr4"])
            d = conv_bwd("fr4", bn_bwd("fr4", d), plan["hgout"])
            done("fr4")
            return d

        if has_hm:
            d_hg = tail_backward()
        else:
            # stand-alone HourglassModule: the seed is the gradient w.r.t. its output, [P, 512] NHWC
            d_hg = t["d_hg"] = self.new((B * g * g, 512), self.adt)
        # ---- hourglass (three consumers of d_hg: up2, skip, depthwise branch).  As in the forward, the down/up branch (a
        # latency-bound chain of ~25 small launches) runs on side stream 2, its weight gradients in line with their own
        # split-K workspace; the depthwise and skip branches proceed on the main stream and need its result (d_a1) last.
        hg2 = plan.get("hg2", False)
        if hg2:
            be.fork(2)
            chain["ws"] = t["wgrad_ws2"] = self.new((8 << 20,), F32)
        d = conv_bwd("up2", bn_bwd("up2", d_hg, shuffle=True), a["up1"])
        d = conv_bwd("up1", bn_bwd("up1", d, shuffle=True), a["bt2"])
        P4 = B * (g // 4) ** 2
        dres = self.new((P4, 128), self.adt)
        d = conv_bwd("bt2", bn_bwd("bt2", d, add1=a["down2"], mode=1, dres=dres), a["bt1"])
        d = conv_bwd("bt1", bn_bwd("bt1", d), a["down2"], dx_residual=dres)
        d = conv_bwd("down2", bn_bwd("down2", d), a["down1"])
        d_a1 = conv_bwd("down1", bn_bwd("down1", d), a["fr0"])
        chain["ws"] = None
        be.side(False)
        # depthwise branch
        d = conv_bwd("pw", bn_bwd("pw", d_hg), a["dw"])
        ddw = bn_bwd("dw", d)
        Ldw = Ls["dw"]
        # like the GEMM weight gradients: nothing on the main stream needs it, second stream (DP_DW_WGRAD_SIDE=0: in line)
        dw_side = overlap and chain["ws"] is None and bool(int(os.environ.get("DP_DW_WGRAD_SIDE", "1")))
        if dw_side:
            be.sync("side_wait")
            be.side(True)
        be.dwconv3x3_wgrad(a["fr0"], ddw, G[Ldw.name + ".weight"], NB=B, H=g, W=g, C=512)
        if dw_side:
            be.side(False)
        dskip = bn_bwd("skip", d_hg)
        if hg2:
            be.join(2)
        done("down1")
        d_a1b = self.new((B * g * g, 512), self.adt)
        be.dwconv3x3(ddw, self.p(Ldw.name + ".weight"), None, d_a1, d_a1b, NB=B, H=g, W=g, C=512, flip=True)
        # skip branch, accumulating into the running gradient of a1
        d_a1c = conv_bwd("skip", dskip, a["fr0"], dx_residual=d_a1b)
        if not has_hm:
            be.join()
            plan["d_in"] = d_a1c           # gradient w.r.t. the hourglass input, [P, 512] NHWC
            be.mark(("grads_final", lay["total"]))
            return
        dfeat = conv_bwd("fr0", bn_bwd("fr0", d_a1c), t["feat"].view(B, g, g, D))
        done("fr0")
        be.join()                # z-head chain (second stream): its input gradient dcur is needed now
        if has_z:
            be.mean_tokens_bwd(dfeat, dcur, B=B, N=N, D=D)
        plan["d_in"] = dfeat               # gradient w.r.t. the NHWC feature map (stand-alone head modules)
        if scope != "model":
            be.mark(("grads_final", lay["total"]))
            return
        if self.train_layers:
            self.record_backward_layers(plan, dfeat, G, flat, ws, done)
            return
        if not self.lora:
            return
        # ---- backbone: final LayerNorm, last block's MLP branch, LoRA adapter
        lp = f"backbone.encoder.layer.{self.L - 1}."
        fz = self.frozen
        t["gx"] = self.new((M, D), F32)
        t["gxs"] = self.new((M, D), self.adt)
        be.layernorm_bwd(dfeat, t["x_last"], self.p("backbone.layernorm.weight"), None, t["gx"], rows=M, D=D, T=T,
                         drop_cls=True, eps=LN_EPS, ls=self.p(lp + "layer_scale2.lambda1"), dx_scaled=t["gxs"])
        t["dpre"] = self.new((M, 4 * D), self.adt)
        be.gemm(t["gxs"], fz["w2T"], t["dpre"], M=M, N=4 * D, K=D, aux_in=t["pre"], ld_aux=4 * D, name="fc2.dgrad")
        t["dxn2"] = self.new((M, D), self.adt)
        be.gemm(t["dpre"], fz["w1T"], t["dxn2"], M=M, N=D, K=4 * D, name="fc1.dgrad")
        t["gmid"] = self.new((M, D), F32)
        be.layernorm_bwd(t["dxn2"], t["x_mid"], self.p(lp + "norm2.weight"), t["gx"], t["gmid"], rows=M, D=D, eps=LN_EPS)
        t["gu"] = self.new((M, self.lora["rank"]), F32)
        be.lora_bwd(t["gmid"], t["y"], t["u"], self.p(self.lora_prefix + "lora_B"), self.p(lp + "layer_scale1.lambda1"),
                    G[self.lora_prefix + "lora_A"], G[self.lora_prefix + "lora_B"], t["gu"], rows=M, D=D, R=self.lora["rank"],
                    scaling=self.lora["alpha"] / self.lora["rank"], p_drop=float(self.lora.get("dropout", 0.0)),
                    seed=self.seed)
        be.mark(("grads_final", lay["total"]))

    def record_backward_layers(self, plan, dfeat, G, flat, ws, done):
        """Backward through the final LayerNorm and the un-frozen encoder layers (autograd of HF:367-386 for
        Dinov2PoseModel(unfreeze_last_n_layers=n), reference model/dinov2_pose.py:25-39).  Per layer, with g = dL/dx_out:
          MLP branch   x_out = x_mid + l2 * m,  m = fc2(gelu(fc1(LN2(x_mid))))
          attention    x_mid = x_in  + l1 * a,  a = dense(attn(qkv(LN1(x_in))))
        Linear weight gradients go through the MN-major tcgen05 wgrad kernel, input gradients through the K-major GEMM
        with the transposed weight copies, bias / LayerScale / LayerNorm-parameter gradients through column reductions."""
        be = self.be
        B, T, M = plan["B"], plan["T"], plan["M"]
        D, L, heads = self.D, self.L, self.heads
        lay = self.layout()
        sv, lw = plan["saved"], plan["lw"]
        scale = 1.0 / math.sqrt(D // heads)
        first = self.train_layers[0]

        def gslice(names, shape):
            # q / k / v gradients as ONE tensor: their flat-buffer slices are adjacent (layer_param_names)
            off0, k0 = lay["offsets"][names[0]]
            for j, n in enumerate(names):
                assert lay["offsets"][n] == (off0 + j * k0, k0), "q/k/v gradient slices must be contiguous"
            return flat[off0:off0 + len(names) * k0].view(shape)

        # weight gradients on the second stream, the activation-gradient chain on the main one (see record_backward)
        overlap = bool(int(os.environ.get("DP_BWD_OVERLAP", "1")))
        if overlap:
            be.fork()
            be.side(False)

        def wg(*a_, **k_):
            if overlap:
                be.sync("side_wait")
                be.side(True)
                be.wgrad(*a_, **k_)
                be.side(False)
            else:
                be.wgrad(*a_, **k_)

        g = self.new((M, D), F32)              # dL/dx_out of the current layer
        gs = self.new((M, D), self.adt)        # bf16(g * layer_scale2)
        top = f"backbone.encoder.layer.{L - 1}."
        be.layernorm_bwd(dfeat, plan["x_final"], self.p("backbone.layernorm.weight"), None, g, rows=M, D=D, T=T,
                         drop_cls=True, eps=LN_EPS, ls=self.p(top + "layer_scale2.lambda1"), dx_scaled=gs)
        stats = self.new((2 * B * heads * T,), F32)
        dpre = self.new((M, 4 * D), self.adt)
        dxn = self.new((M, D), self.adt)
        gmid = self.new((M, D), F32)
        gmids = self.new((M, D), self.adt)
        dctx = self.new((M, D), self.adt)
        dqkv = self.new((M, 3 * D), self.adt)
        for i in reversed(self.train_layers):
            lp, ap = f"backbone.encoder.layer.{i}.", self.att_prefix[i]
            s_, w = sv[i], lw[i]
            # ---- MLP branch
            be.colsum_prod(g, s_["m"], G[lp + "layer_scale2.lambda1"], P=M, C=D)
            be.colsum(gs, G[lp + "mlp.fc2.bias"], P=M, C=D, ld=D)
            wg(gs, s_["h"], G[lp + "mlp.fc2.weight"], Mc=D, Nc=4 * D, so_m=4 * D, so_n=1, P=M, name=f"fc2_{i}.wgrad",
                     workspace=ws)
            be.gemm(gs, w["w2T"], dpre, M=M, N=4 * D, K=D, aux_in=s_["pre"], ld_aux=4 * D, name=f"fc2_{i}.dgrad")
            be.colsum(dpre, G[lp + "mlp.fc1.bias"], P=M, C=4 * D, ld=4 * D)
            wg(dpre, s_["xn2"], G[lp + "mlp.fc1.weight"], Mc=4 * D, Nc=D, so_m=D, so_n=1, P=M, name=f"fc1_{i}.wgrad",
                     workspace=ws)
            be.gemm(dpre, w["w1T"], dxn, M=M, N=D, K=4 * D, name=f"fc1_{i}.dgrad")
            be.layernorm_bwd_params(dxn, s_["x_mid"], G[lp + "norm2.weight"], G[lp + "norm2.bias"], rows=M, D=D, eps=LN_EPS)
            be.layernorm_bwd(dxn, s_["x_mid"], self.p(lp + "norm2.weight"), g, gmid, rows=M, D=D, eps=LN_EPS,
                             ls=self.p(lp + "layer_scale1.lambda1"), dx_scaled=gmids)
            # ---- attention branch
            be.colsum_prod(gmid, s_["a"], G[lp + "layer_scale1.lambda1"], P=M, C=D)
            be.colsum(gmids, G[ap + "output.dense.bias"], P=M, C=D, ld=D)
            wg(gmids, s_["ctx"], G[ap + "output.dense.weight"], Mc=D, Nc=D, so_m=D, so_n=1, P=M, name=f"proj{i}.wgrad",
                     workspace=ws)
            be.gemm(gmids, w["woT"], dctx, M=M, N=D, K=D, name=f"proj{i}.dgrad")
            be.attention_bwd(s_["qkv"], s_["ctx"], dctx, dqkv, stats, B=B, T=T, heads=heads, scale=scale)
            qkv_names = [ap + f"attention.{n}." for n in ("query", "key", "value")]
            be.colsum(dqkv, gslice([n + "bias" for n in qkv_names], (3 * D,)), P=M, C=3 * D, ld=3 * D)
            wg(dqkv, s_["xn1"], gslice([n + "weight" for n in qkv_names], (3 * D, D)), Mc=3 * D, Nc=D, so_m=D, so_n=1,
                     P=M, name=f"qkv{i}.wgrad", workspace=ws)
            be.gemm(dqkv, w["wqkvT"], dxn, M=M, N=D, K=3 * D, name=f"qkv{i}.dgrad")
            be.layernorm_bwd_params(dxn, s_["x_in"], G[lp + "norm1.weight"], G[lp + "norm1.bias"], rows=M, D=D, eps=LN_EPS)
            if i > first:
                below = f"backbone.encoder.layer.{i - 1}."
                be.layernorm_bwd(dxn, s_["x_in"], self.p(lp + "norm1.weight"), gmid, g, rows=M, D=D, eps=LN_EPS,
                                 ls=self.p(below + "layer_scale2.lambda1"), dx_scaled=gs)
            if overlap:
                be.sync("main_wait")
            done(f"layer{i}")
        if overlap:
            be.join()
        be.mark(("grads_final", lay["total"]))

    # ------------------------------------------------------------------ running
    def check_frozen(self):
        if self.frozen is not None and self._versions(frozen=True) != self.frozen_versions:
            # frozen parameters were modified (load_state_dict, manual edit): repack and drop plans
            self.frozen = None
            self.plans = {}

    def forward(self, pixel_values, training):
        self.check_frozen()
        B, C, H, W = pixel_values.shape
        if C != 3:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                             f"configuration. Expected 3 but got {C}.")   # HF modeling_dinov2.py:143-147
        plan = self.get_plan(B, H, W, training)
        plan["t"]["px"].copy_(pixel_values, non_blocking=True)
        if training:
            self.seed.add_(1)
            plan["generation"] = plan.get("generation", 0) + 1   # stamps the contents of the saved-activation buffers
            plan["fwd"].run()
            return plan
        # inference: the ~130 launches of the forward program are replayed as ONE CUDA graph (batch-1 latency is
        # launch-bound otherwise); first call runs eagerly (lazy kernel attributes), second call captures
        if getattr(self.be, "name", "") != "cuda" or not self.cfg.get("infer_graph", True):
            plan["fwd"].run()
            return plan
        calls = plan["calls"] = plan.get("calls", 0) + 1
        if calls == 1:
            plan["fwd"].run()
        else:
            if plan.get("graph") is None:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side), torch.no_grad():
                    with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                        plan["fwd"].run()
                torch.cuda.current_stream().wait_stream(side)
                plan["graph"] = g
            plan["graph"].replay()
        return plan

    def backward(self, plan, dhm, dz, on_mark=None):
        """dhm / dz: gradients w.r.t. the outputs (copied into the plan's static seed buffers), or the strings
        "static" when a kernel (dp_pose_loss) already wrote plan["t"]["dhm"] / ["dz"] in place."""
        t = plan["t"]
        if not isinstance(dhm, str):
            if dhm is None:
                t["dhm"].zero_()
            else:
                t["dhm"].copy_(dhm)
        if not isinstance(dz, str):
            if dz is None:
                t["dz"].zero_()
            else:
                t["dz"].copy_(dz)
        plan["bwd"].run(on_mark=on_mark)
        return plan["grads"]
