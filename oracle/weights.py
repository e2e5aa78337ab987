"""Deterministic synthetic weights for the dino_pose models (TEST INFRASTRUCTURE).

There is no network and the reference ships only git-LFS stubs, so every parity
test runs on random weights of the named architecture.  Shipping 22-300 M floats as
fixtures is not possible; instead every tensor is a pure function of
``(seed, state_dict key, shape)`` -- the real reference (in ``make_golden.py``), the
oracle and the CUDA path are all loaded from the same generated ``state_dict``.

Key names / shapes follow the reference's ``state_dict`` (SURVEY.md section 8b;
reference model/dinov2_pose.py:13-54,180-219, model/lora.py:14-15,
model/pose_heads.py:212-343,134-159 and HF modeling_dinov2.py:44-52,139,199-201,250,275,318-322,354-361,449).
``make_golden.py`` asserts that the generated keys and shapes equal the real
reference's ``state_dict`` exactly.

Unlike the reference's default init (LoRA ``B = 0``, biases 0, LayerScale 1), values
are chosen so that every term on the path is numerically visible: non-zero biases,
non-trivial LayerNorm/BatchNorm affine parameters and running statistics, and a
non-zero ``lora_B`` so the adapter contributes to the output.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

ARCHS = {
    # name: (hidden D, layers L, heads h)   -- HF configuration_dinov2.py + hub configs
    "facebook/dinov2-small": (384, 12, 6),
    "facebook/dinov2-base": (768, 12, 12),
    "facebook/dinov2-large": (1024, 24, 16),
    # tiny configuration used by fast CPU tests (not a reference checkpoint)
    "test/dinov2-tiny": (128, 2, 2),
}

NUM_POS = 37 * 37 + 1  # hub checkpoints are image_size 518 / patch 14


def _bn(prefix, c):
    return [
        (prefix + ".weight", (c,), "bn_w"),
        (prefix + ".bias", (c,), "bn_b"),
        (prefix + ".running_mean", (c,), "bn_rm"),
        (prefix + ".running_var", (c,), "bn_rv"),
        (prefix + ".num_batches_tracked", (), "bn_n"),
    ]


def _conv(prefix, cout, cin_per_group, k):
    fan_in = cin_per_group * k * k
    return [
        (prefix + ".weight", (cout, cin_per_group, k, k), ("w", 1.0 / math.sqrt(fan_in))),
        (prefix + ".bias", (cout,), ("b", 0.05)),
    ]


def _convT(prefix, cin, cout, k, stride):
    # effective fan-in of a transposed conv output element ~ cin * (k/stride)^2
    fan_in = cin * max(1.0, (k / stride) ** 2)
    return [
        (prefix + ".weight", (cin, cout, k, k), ("w", 1.0 / math.sqrt(fan_in))),
        (prefix + ".bias", (cout,), ("b", 0.05)),
    ]


def _linear(prefix, cout, cin, std=None, bstd=0.05):
    std = 1.0 / math.sqrt(cin) if std is None else std
    return [
        (prefix + ".weight", (cout, cin), ("w", std)),
        (prefix + ".bias", (cout,), ("b", bstd)),
    ]


def heads_spec(feat_dim, num_keypoints=24, heatmap_size=48, spatial_input_size=16,
               z_hidden=(1024, 512, 256), prefix="pose_heads."):
    """Spec of ``SpatialAwarePoseHeads`` (reference model/pose_heads.py:287-400)."""
    s = []
    hh = prefix + "heatmap_head."
    fr = hh + "feature_refine."
    s += _conv(fr + "0", 512, feat_dim, 3) + _bn(fr + "1", 512)
    hg = fr + "3."
    s += _conv(hg + "depthwise_conv.0", 512, 1, 3) + _bn(hg + "depthwise_conv.1", 512)
    s += _conv(hg + "depthwise_conv.3", 512, 512, 1) + _bn(hg + "depthwise_conv.4", 512)
    s += _conv(hg + "down1.0", 256, 512, 3) + _bn(hg + "down1.1", 256)
    s += _conv(hg + "down2.0", 128, 256, 3) + _bn(hg + "down2.1", 128)
    s += _conv(hg + "bottleneck.0", 128, 128, 3) + _bn(hg + "bottleneck.1", 128)
    s += _conv(hg + "bottleneck.3", 128, 128, 3) + _bn(hg + "bottleneck.4", 128)
    s += _convT(hg + "up1.0", 128, 256, 2, 2) + _bn(hg + "up1.1", 256)
    s += _convT(hg + "up2.0", 256, 512, 2, 2) + _bn(hg + "up2.1", 512)
    s += _conv(hg + "skip.0", 512, 512, 1) + _bn(hg + "skip.1", 512)
    s += _conv(fr + "4", 256, 512, 3) + _bn(fr + "5", 256)
    # upsampling stages (pose_heads.py:316-332)
    cur, cin, i = spatial_input_size, 256, 0
    while cur < heatmap_size:
        cout = max(128, cin // 2)
        stride = heatmap_size // cur
        s += _convT(hh + f"upsampling.{i}.0", cin, cout, 4, stride) + _bn(hh + f"upsampling.{i}.1", cout)
        cur *= 2
        cin = cout
        i += 1
    final_c = cin if i else 256
    s += _conv(hh + "prediction.0", 64, final_c, 3) + _bn(hh + "prediction.1", 64)
    s += _conv(hh + "prediction.3", num_keypoints, 64, 1)
    # z head (pose_heads.py:134-159)
    zin, j = feat_dim, 0
    for hd in z_hidden:
        s += _linear(prefix + f"z_head.mlp.{j}", hd, zin)
        zin, j = hd, j + 3
    s += _linear(prefix + f"z_head.mlp.{j}", num_keypoints, zin)
    return s


def backbone_spec(arch, lora_rank=0, prefix="backbone."):
    """Spec of HF ``Dinov2Model`` (+ the reference's LoRAAttention on the last layer)."""
    D, L, _ = ARCHS[arch]
    s = []
    e = prefix + "embeddings."
    s += [(e + "cls_token", (1, 1, D), ("w", 0.02)),
          (e + "mask_token", (1, D), "zero"),
          (e + "position_embeddings", (1, NUM_POS, D), ("w", 0.02)),
          (e + "patch_embeddings.projection.weight", (D, 3, 14, 14), ("w", 0.02)),
          (e + "patch_embeddings.projection.bias", (D,), ("b", 0.02))]
    for i in range(L):
        p = prefix + f"encoder.layer.{i}."
        att = p + "attention."
        if lora_rank and i >= L - 1:      # reference model/dinov2_pose.py:197-204
            att = p + "attention.original_attention."
        s += [(p + "norm1.weight", (D,), "ln_w"), (p + "norm1.bias", (D,), "ln_b")]
        for n in ("query", "key", "value"):
            s += _linear(att + "attention." + n, D, D, std=0.02, bstd=0.02)
        s += _linear(att + "output.dense", D, D, std=0.02, bstd=0.02)
        if lora_rank and i >= L - 1:
            s += [(p + "attention.lora_output.lora_A", (D, lora_rank), ("u", 1.0 / math.sqrt(lora_rank))),
                  (p + "attention.lora_output.lora_B", (lora_rank, D), ("w", 0.02))]
        s += [(p + "layer_scale1.lambda1", (D,), "ls")]
        s += [(p + "norm2.weight", (D,), "ln_w"), (p + "norm2.bias", (D,), "ln_b")]
        s += _linear(p + "mlp.fc1", 4 * D, D, std=0.02, bstd=0.02)
        s += _linear(p + "mlp.fc2", D, 4 * D, std=0.02, bstd=0.02)
        s += [(p + "layer_scale2.lambda1", (D,), "ls")]
    s += [(prefix + "layernorm.weight", (D,), "ln_w"), (prefix + "layernorm.bias", (D,), "ln_b")]
    return s


def model_spec(arch, lora_rank=0, num_keypoints=24, heatmap_size=48):
    D = ARCHS[arch][0]
    return backbone_spec(arch, lora_rank) + heads_spec(D, num_keypoints, heatmap_size)


def _gen(name, shape, kind, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) * 2654435761 + seed * 97 + 12345) % (2 ** 63 - 1))
    if isinstance(kind, tuple):
        k, scale = kind
        if k == "u":
            return (torch.rand(shape, generator=g) * 2 - 1) * scale
        return torch.randn(shape, generator=g) * scale
    if kind == "zero":
        return torch.zeros(shape)
    if kind == "ln_w" or kind == "bn_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "ln_b":
        return 0.05 * torch.randn(shape, generator=g)
    if kind == "bn_b" or kind == "bn_rm":
        return 0.1 * torch.randn(shape, generator=g)
    if kind == "bn_rv":
        return 0.5 + torch.rand(shape, generator=g)
    if kind == "bn_n":
        return torch.zeros((), dtype=torch.int64)
    if kind == "ls":
        return 1.0 + 0.05 * torch.randn(shape, generator=g)
    raise ValueError(kind)


def make_state_dict(arch="facebook/dinov2-small", seed=0, lora_rank=0, num_keypoints=24,
                    heatmap_size=48):
    """``OrderedDict`` name -> fp32 CPU tensor, reference ``state_dict`` naming."""
    sd = OrderedDict()
    for name, shape, kind in model_spec(arch, lora_rank, num_keypoints, heatmap_size):
        sd[name] = _gen(name, shape, kind, seed).contiguous()
    return sd


from dino_pose_b200.synthetic import make_inputs  # noqa: E402,F401  (one generator for tests, bench and tools)
