"""Plain-torch fp32 restatement of the reference forward path (TEST INFRASTRUCTURE).

Functional (no nn.Module): every function takes the reference-named ``state_dict``
``sd`` and tensors, and follows the cited reference lines.  ``HF:`` =
transformers 5.5.0 ``models/dinov2/modeling_dinov2.py`` (third-party dependency of the
reference, not vendored under /root/reference; the reference requires
``transformers>=4.30.0`` un-pinned, requirements.txt:6).  Gradients come from
autograd over these same functions.

Pinned against the real reference by ``tests/golden/*.npz`` (``make_golden.py``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .weights import ARCHS

PATCH = 14
LN_EPS = 1e-6   # HF configuration_dinov2.py layer_norm_eps
BN_EPS = 1e-5   # torch BatchNorm2d default used by model/pose_heads.py
BN_MOM = 0.1


# --------------------------------------------------------------------------- backbone
def interpolate_pos_encoding(pos, gh, gw):
    """HF:57-95.  pos [1, 1+S*S, D] -> [1, 1+gh*gw, D] (bicubic, align_corners=False, fp32)."""
    npos = pos.shape[1] - 1
    if gh * gw == npos and gh == gw:
        return pos
    D = pos.shape[-1]
    cls_pos, patch_pos = pos[:, :1], pos[:, 1:]
    s = int(npos ** 0.5)
    patch_pos = patch_pos.reshape(1, s, s, D).permute(0, 3, 1, 2)
    patch_pos = F.interpolate(patch_pos.float(), size=(gh, gw), mode="bicubic", align_corners=False)
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, D)
    return torch.cat((cls_pos, patch_pos), dim=1)


def embeddings(sd, px, pfx="backbone.embeddings."):
    """HF:97-116 (+ patch conv HF:141-149).  px [B,3,H,W] -> tokens [B,T,D]."""
    B, _, H, W = px.shape
    w = sd[pfx + "patch_embeddings.projection.weight"]
    b = sd[pfx + "patch_embeddings.projection.bias"]
    x = F.conv2d(px, w, b, stride=PATCH).flatten(2).transpose(1, 2)
    cls = sd[pfx + "cls_token"].expand(B, -1, -1)
    x = torch.cat((cls, x), dim=1)
    return x + interpolate_pos_encoding(sd[pfx + "position_embeddings"], H // PATCH, W // PATCH)


def attention_block(sd, xn, att, heads):
    """HF:203-234 self-attention (softmax(q k^T / sqrt(dh)) v) + HF:249-252 output dense."""
    B, T, D = xn.shape
    dh = D // heads

    def proj(n):
        return F.linear(xn, sd[att + f"attention.{n}.weight"], sd[att + f"attention.{n}.bias"]) \
            .view(B, T, heads, dh).transpose(1, 2)

    q, k, v = proj("query"), proj("key"), proj("value")
    s = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5)
    p = torch.softmax(s, dim=-1)
    ctx = torch.matmul(p, v).transpose(1, 2).reshape(B, T, D)
    return F.linear(ctx, sd[att + "output.dense.weight"], sd[att + "output.dense.bias"])


def lora_delta(sd, y, lpfx, alpha, rank, p_drop=0.0, training=False):
    """reference model/lora.py:26-28: dropout(y @ A @ B) * (alpha / rank)."""
    d = y @ sd[lpfx + "lora_A"] @ sd[lpfx + "lora_B"]
    d = F.dropout(d, p_drop, training)
    return d * (alpha / rank)


def encoder_layer(sd, x, p, heads, lora=None, training=False, aux=None):
    """HF:367-386; on the LoRA layer the attention block is wrapped per model/lora.py:53-59."""
    D = x.shape[-1]
    xn = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    if lora is not None:
        y = attention_block(sd, xn, p + "attention.original_attention.", heads)
        if aux is not None:
            aux["lora_in"] = y
        y = y + lora_delta(sd, y, p + "attention.lora_output.", lora["alpha"], lora["rank"],
                           lora.get("dropout", 0.0), training)
    else:
        y = attention_block(sd, xn, p + "attention.", heads)
    x = x + y * sd[p + "layer_scale1.lambda1"]
    xn2 = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    h = F.linear(xn2, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.gelu(h)  # HF ACT2FN["gelu"] = erf GELU
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h * sd[p + "layer_scale2.lambda1"]


def backbone(sd, px, arch, lora=None, training=False, aux=None, pfx="backbone."):
    """HF:459-485: embeddings -> L layers -> final LayerNorm.  Returns [B,T,D]."""
    D, L, heads = ARCHS[arch]
    x = embeddings(sd, px, pfx + "embeddings.")
    if aux is not None:
        aux["tokens_embed"] = x
        aux["hidden"] = []
    for i in range(L):
        use_lora = lora if (lora is not None and i >= L - 1) else None  # dinov2_pose.py:197-204
        x = encoder_layer(sd, x, pfx + f"encoder.layer.{i}.", heads, use_lora, training, aux)
        if aux is not None:
            aux["hidden"].append(x)
    x = F.layer_norm(x, (D,), sd[pfx + "layernorm.weight"], sd[pfx + "layernorm.bias"], LN_EPS)
    if aux is not None:
        aux["tokens_final"] = x
    return x


# --------------------------------------------------------------------------- heads
def _bn(sd, x, p, training, relu=True):
    y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                     sd[p + ".bias"], training, BN_MOM, BN_EPS)
    return F.relu(y) if relu else y


def _conv(sd, x, p, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride, padding, 1, groups)


def _convT(sd, x, p, stride, padding=0):
    return F.conv_transpose2d(x, sd[p + ".weight"], sd[p + ".bias"], stride, padding)


def hourglass(sd, x, p, training):
    """reference model/pose_heads.py:268-285 (ctor :212-266)."""
    C = x.shape[1]
    skip = _bn(sd, _conv(sd, x, p + "skip.0"), p + "skip.1", training)
    d = _bn(sd, _conv(sd, x, p + "depthwise_conv.0", 1, 1, C), p + "depthwise_conv.1", training)
    d = _bn(sd, _conv(sd, d, p + "depthwise_conv.3"), p + "depthwise_conv.4", training)
    d1 = _bn(sd, _conv(sd, x, p + "down1.0", 2, 1), p + "down1.1", training)
    d2 = _bn(sd, _conv(sd, d1, p + "down2.0", 2, 1), p + "down2.1", training)
    b = _bn(sd, _conv(sd, d2, p + "bottleneck.0", 1, 1), p + "bottleneck.1", training)
    b = _bn(sd, _conv(sd, b, p + "bottleneck.3", 1, 1), p + "bottleneck.4", training, relu=False)
    b = F.relu(b + d2)
    u1 = _bn(sd, _convT(sd, b, p + "up1.0", 2), p + "up1.1", training)
    u2 = _bn(sd, _convT(sd, u1, p + "up2.0", 2), p + "up2.1", training)
    return u2 + skip + d


def heatmap_head(sd, fmap, training, heatmap_size=48, spatial_input_size=16,
                 p="pose_heads.heatmap_head.", aux=None):
    """reference model/pose_heads.py:345-361 (ctor :293-343)."""
    fr = p + "feature_refine."
    x = _bn(sd, _conv(sd, fmap, fr + "0", 1, 1), fr + "1", training)
    if aux is not None:
        aux["fr0"] = x
    x = hourglass(sd, x, fr + "3.", training)
    if aux is not None:
        aux["hg"] = x
    x = _bn(sd, _conv(sd, x, fr + "4", 1, 1), fr + "5", training)
    if aux is not None:
        aux["fr4"] = x
    cur, i = spatial_input_size, 0
    while cur < heatmap_size:
        stride = heatmap_size // cur
        x = _bn(sd, _convT(sd, x, p + f"upsampling.{i}.0", stride, 1), p + f"upsampling.{i}.1", training)
        if aux is not None:
            aux[f"up{i}"] = x
        cur *= 2
        i += 1
    x = _bn(sd, _conv(sd, x, p + "prediction.0", 1, 1), p + "prediction.1", training)
    if aux is not None:
        aux["pred0"] = x
    x = _conv(sd, x, p + "prediction.3")
    if cur != heatmap_size:   # use_interpolation (:343,:353-359)
        x = F.interpolate(x, size=(heatmap_size, heatmap_size), mode="bilinear", align_corners=False)
    return x


def z_head(sd, feats, training, p_drop=0.0, p="pose_heads.z_head.mlp.", n_hidden=3):
    """reference model/pose_heads.py:161-162 (ctor :134-159)."""
    x = feats
    for j in range(n_hidden):
        x = F.relu(F.linear(x, sd[p + f"{3 * j}.weight"], sd[p + f"{3 * j}.bias"]))
        x = F.dropout(x, p_drop, training)
    return F.linear(x, sd[p + f"{3 * n_hidden}.weight"], sd[p + f"{3 * n_hidden}.bias"])


def pose_heads(sd, fmap, training, heatmap_size=48, z_dropout=0.0, aux=None):
    """reference model/pose_heads.py:395-400."""
    hm = heatmap_head(sd, fmap, training, heatmap_size, aux=aux)
    z = z_head(sd, fmap.mean(dim=(2, 3)), training, z_dropout)
    return hm, z


def model_forward(sd, px, arch, lora=None, training=False, heatmap_size=48, z_dropout=0.0, aux=None,
                  heads_training=None):
    """reference model/dinov2_pose.py:143-157 / :292-306.

    ``sd`` BatchNorm running statistics are updated IN PLACE when ``training``
    (torch semantics) -- pass clones if that matters.  ``heads_training=False`` with
    ``training=True`` is ``model.train(); model.pose_heads.eval()`` (nn.Module.train is
    per sub-module): BatchNorm2d / Dropout of the heads in eval mode, LoRA dropout active.
    """
    heads_training = training if heads_training is None else heads_training
    tok = backbone(sd, px, arch, lora, training, aux)
    patch = tok[:, 1:, :]
    B, N, D = patch.shape
    H = W = int(N ** 0.5)
    fmap = patch.contiguous().view(B, H, W, D).permute(0, 3, 1, 2).contiguous()
    return pose_heads(sd, fmap, heads_training, heatmap_size, z_dropout, aux)


# --------------------------------------------------------------------------- losses (train.py:89-120)
def keypoint_loss(pred, target, conf):
    mask = (conf > 1).float()[:, :, None, None].expand_as(pred)
    diff = (pred - target) ** 2
    w = torch.exp(-diff.detach())
    return (w * diff * mask).mean()


def z_loss(pred_z, target_z, conf):
    mask = (conf > 1).float()
    return torch.abs(pred_z * mask - target_z * mask).mean()


class DynamicLossWeighting:
    """reference train.py:17-87 (host-side scalar state)."""

    def __init__(self, initial_weight=0.1, adjustment_rate=0.1):
        self.weight = initial_weight
        self.adjustment_rate = adjustment_rate
        self.kp_avg = None
        self.z_avg = None
        self.momentum = 0.9

    def update(self, kp, z):
        if self.kp_avg is None:
            self.kp_avg, self.z_avg = kp, z
        else:
            self.kp_avg = self.momentum * self.kp_avg + (1 - self.momentum) * kp
            self.z_avg = self.momentum * self.z_avg + (1 - self.momentum) * z
        tgt = (kp + 1e-8) / (z + 1e-8)
        self.weight = (1 - self.adjustment_rate) * self.weight + self.adjustment_rate * tgt
        self.weight = max(1e-3, min(10.0, self.weight))
        return self.weight

    def balanced(self, kp_loss, zl):
        if self.kp_avg is None or self.z_avg is None:
            return kp_loss + self.weight * zl
        return kp_loss / (self.kp_avg + 1e-8) + zl / (self.z_avg + 1e-8)


def trainable_names(sd, lora, unfreeze=0, arch=None):
    """Parameters that receive gradients: every head parameter, plus lora_A/B in the LoRA model (reference
    model/dinov2_pose.py:193-204: backbone frozen) or every parameter of the last ``unfreeze`` encoder layers in
    ``Dinov2PoseModel(unfreeze_last_n_layers=unfreeze)`` (reference model/dinov2_pose.py:25-39)."""
    out = []
    layers = ()
    if unfreeze:
        L = ARCHS[arch][1]
        layers = tuple(f"backbone.encoder.layer.{i}." for i in range(L - unfreeze, L))
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            continue
        if "running_" in k:
            continue
        if k.startswith("pose_heads.") or ".lora_output." in k or (layers and k.startswith(layers)):
            out.append(k)
    return out


def loss_and_grads(sd, batch, arch, lora, training=True, z_dropout=0.0, unfreeze=0, heads_training=None):
    """One forward + reference losses + backward.  Loss = kp + 0.1 * z on the first step
    (``DynamicLossWeighting.get_balanced_loss`` falls back to ``kp + weight*z`` with
    weight 0.1 until averages exist -- but ``update`` is called first in train.py:154-163,
    so the first-step loss is the normalised form; both are exposed)."""
    names = trainable_names(sd, lora, unfreeze, arch)
    for n in names:
        sd[n].requires_grad_(True)
    hm, z = model_forward(sd, batch["pixel_values"], arch, lora, training, z_dropout=z_dropout,
                          heads_training=heads_training)
    conf = batch["keypoints"][..., 2]
    kp = keypoint_loss(hm, batch["heatmaps"], conf)
    zl = z_loss(z, batch["z"], conf)
    w = DynamicLossWeighting()
    w.update(kp.item(), zl.item())
    loss = w.balanced(kp, zl)
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    for n in names:
        sd[n].requires_grad_(False)
    return {"heatmaps": hm.detach(), "z": z.detach(), "kp_loss": kp.detach(), "z_loss": zl.detach(),
            "loss": loss.detach(), "grads": {n: g for n, g in zip(names, grads)}}
