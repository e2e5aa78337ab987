"""Pose heads (mirrors reference model/pose_heads.py): parameter containers with the reference's module tree
(so ``state_dict`` keys are identical) whose ``forward`` runs on the sm_100a kernels.

Layer tables are written in a tiny spec language instead of literal ``nn.Sequential`` code:
  ``c<k>[s<stride>][p<pad>][g]:<out>``  Conv2d (g = depthwise)     ``t<k>s<stride>[p<pad>][o<outpad>]:<out>``  ConvTranspose2d
  ``bn``  BatchNorm2d of the running width                         ``relu``  ReLU (no parameters)
  ``fc:<out>``  Linear                                              ``drop:<p>``  Dropout
"""
from __future__ import annotations

import re
from typing import Optional, Tuple

import torch
import torch.nn as nn

_TOKEN = re.compile(r"^(?P<op>[ct])(?P<k>\d+)(s(?P<s>\d+))?(p(?P<p>\d+))?(o(?P<op_>\d+))?(?P<g>g)?:(?P<o>\d+)$")


def _build(spec: str, width: int):
    """spec string -> (nn.Sequential, output width)."""
    mods = []
    for tok in spec.split():
        if tok == "bn":
            mods.append(nn.BatchNorm2d(width))
        elif tok == "relu":
            mods.append(nn.ReLU(inplace=True))
        elif tok.startswith("fc:"):
            out = int(tok[3:])
            mods.append(nn.Linear(width, out))
            width = out
        elif tok.startswith("drop:"):
            mods.append(nn.Dropout(float(tok[5:])))
        else:
            m = _TOKEN.match(tok)
            if not m:
                raise ValueError(f"bad layer token {tok!r}")
            k, s, p, out = int(m["k"]), int(m["s"] or 1), int(m["p"] or 0), int(m["o"])
            if m["op"] == "c":
                mods.append(nn.Conv2d(width, out, k, s, p, groups=width if m["g"] else 1))
            else:
                mods.append(nn.ConvTranspose2d(width, out, k, s, p, output_padding=int(m["op_"] or 0)))
            width = out
    return nn.Sequential(*mods), width


def _run_heads(module, kind, x):
    from ..functional import run_head_module
    return run_head_module(module, kind, x)


class HourglassModule(nn.Module):
    """reference model/pose_heads.py:211-285: skip + depthwise branch + (down, down, bottleneck, up, up)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        c, h, q = out_channels, out_channels // 2, out_channels // 4
        self.depthwise_conv, _ = _build(f"c3p1g:{in_channels} bn relu c1:{c} bn relu", in_channels)
        self.down1, _ = _build(f"c3s2p1:{h} bn relu", in_channels)
        self.down2, _ = _build(f"c3s2p1:{q} bn relu", h)
        self.bottleneck, _ = _build(f"c3p1:{q} bn relu c3p1:{q} bn", q)
        self.bottleneck_relu = nn.ReLU(inplace=True)
        self.up1, _ = _build(f"t2s2:{h} bn relu", q)
        self.up2, _ = _build(f"t2s2:{c} bn relu", h)
        self.skip, _ = _build(f"c1:{c} bn relu", in_channels)

    def forward(self, x):
        return _run_heads(self, "hourglass", x)


class SpatialAwareHeatmapHead(nn.Module):
    """reference model/pose_heads.py:287-361."""

    def __init__(self, feat_channels: int = 768, num_keypoints: int = 24, heatmap_size: int = 48,
                 spatial_input_size: int = 14):
        super().__init__()
        self.feat_channels, self.num_keypoints = feat_channels, num_keypoints
        self.heatmap_size, self.spatial_input_size = heatmap_size, spatial_input_size
        front, _ = _build("c3p1:512 bn relu", feat_channels)
        back, width = _build("c3p1:256 bn relu", 512)
        self.feature_refine = nn.Sequential(*front, HourglassModule(512, 512), *back)
        stages, size = [], spatial_input_size
        while size < heatmap_size:          # reference :320-330 (kernel 4, stride = heatmap // current, pad 1)
            out = max(128, width // 2)
            stage, width = _build(f"t4s{heatmap_size // size}p1:{out} bn relu", width)
            stages.append(stage)
            size *= 2
        self.upsampling = nn.Sequential(*stages)
        self.prediction, _ = _build(f"c3p1:64 bn relu c1:{num_keypoints}", width)
        self.target_size = heatmap_size
        self.use_interpolation = size != heatmap_size

    def forward(self, feature_map: torch.Tensor) -> torch.Tensor:
        return _run_heads(self, "heatmap_head", feature_map)


class ZCoordinateHead(nn.Module):
    """reference model/pose_heads.py:128-162: Linear/ReLU/Dropout stack + final Linear."""

    def __init__(self, feat_dim: int, num_keypoints: int, hidden_dims: Tuple[int, ...] = (1024, 512),
                 dropout_rate: float = 0.2):
        super().__init__()
        self.feat_dim, self.num_keypoints = feat_dim, num_keypoints
        self.hidden_dims, self.dropout_rate = tuple(hidden_dims), dropout_rate
        spec = " ".join(f"fc:{h} relu drop:{dropout_rate}" for h in hidden_dims) + f" fc:{num_keypoints}"
        self.mlp, _ = _build(spec, feat_dim)
        for m in self.mlp:                   # reference uses non-inplace ReLU here; no numerical difference
            if isinstance(m, nn.ReLU):
                m.inplace = False

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        return _run_heads(self, "z_head", features)


class SpatialAwarePoseHeads(nn.Module):
    """reference model/pose_heads.py:364-400: heat-map head on the feature map + z head on its spatial mean."""

    def __init__(self, feat_channels: int = 768, num_keypoints: int = 24, heatmap_size: int = 48,
                 spatial_input_size: int = 14, z_coord_config: Optional[dict] = None):
        super().__init__()
        self.heatmap_head = SpatialAwareHeatmapHead(feat_channels, num_keypoints, heatmap_size, spatial_input_size)
        self.z_head = ZCoordinateHead(feat_channels, num_keypoints, **(z_coord_config or {}))

    def forward(self, feature_map: torch.Tensor):
        return _run_heads(self, "pose_heads", feature_map)


class HeatmapHead(nn.Module):
    """reference model/pose_heads.py:6-125: MLP projection of a feature VECTOR to a ``spatial_size``^2 map, then stride-2
    transposed-conv stages up to ``heatmap_size``.  Never instantiated by the reference (dead code) and not on
    the hot path: kept as an ordinary torch module (same parameter tree and ``forward``) so imports and old checkpoints
    keep working; it runs on whatever device torch puts it on and uses none of the sm_100a kernels."""

    def __init__(self, feat_dim: int, num_keypoints: int, heatmap_size: int = 48, intermediate_features: int = 512,
                 spatial_size: int = 6):
        super().__init__()
        self.feat_dim, self.num_keypoints, self.heatmap_size = feat_dim, num_keypoints, heatmap_size
        self.spatial_size, self.intermediate_features = spatial_size, intermediate_features
        self.feature_projection, _ = _build(
            f"fc:2048 relu drop:0.1 fc:1024 relu drop:0.1 fc:{spatial_size * spatial_size * intermediate_features} relu", feat_dim)
        for m in self.feature_projection:
            if isinstance(m, nn.ReLU):
                m.inplace = False
        self.num_stages, size = 0, spatial_size
        while size < heatmap_size:
            size, self.num_stages = size * 2, self.num_stages + 1
        stages, size, width, nxt = [], spatial_size * 2, 256, 128
        stages.append(_build("t3s2p1o1:256 bn relu", intermediate_features)[0])
        while size < heatmap_size:                      # reference :76-86
            stages.append(_build(f"t3s2p1o1:{nxt} bn relu", width)[0])
            size, width, nxt = size * 2, nxt, max(64, nxt // 2)
        if size > heatmap_size:                         # overshoot: conv + adaptive pooling to the exact size (:89-96)
            stages.append(nn.Sequential(*_build("c3p1:64 bn relu", width)[0], nn.AdaptiveAvgPool2d(heatmap_size)))
        elif width != 64:                               # :97-103
            stages.append(_build("c3p1:64 bn relu", width)[0])
        for st in stages:
            for m in st:
                if isinstance(m, nn.ReLU):
                    m.inplace = False
        self.upsampling_layers = nn.ModuleList(stages)
        self.prediction_layer = nn.Conv2d(64, num_keypoints, kernel_size=1)

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        x = self.feature_projection(features)
        x = x.view(features.size(0), self.intermediate_features, self.spatial_size, self.spatial_size)
        for layer in self.upsampling_layers:
            x = layer(x)
        return self.prediction_layer(x)


class PoseHeads(nn.Module):
    """reference model/pose_heads.py:165-208 (``HeatmapHead`` + ``ZCoordinateHead`` on a feature vector) -- dead code in
    the reference, see ``HeatmapHead``; the z head is the kernel-backed ``ZCoordinateHead``."""

    def __init__(self, feat_dim: int, num_keypoints: int, heatmap_size: int = 48, heatmap_config: Optional[dict] = None,
                 z_coord_config: Optional[dict] = None):
        super().__init__()
        self.heatmap_head = HeatmapHead(feat_dim=feat_dim, num_keypoints=num_keypoints, heatmap_size=heatmap_size,
                                        **(heatmap_config or {}))
        self.z_head = ZCoordinateHead(feat_dim=feat_dim, num_keypoints=num_keypoints, **(z_coord_config or {}))

    def forward(self, features: torch.Tensor):
        return self.heatmap_head(features), self.z_head(features)

    def count_parameters(self, trainable_only: bool = True) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad or not trainable_only)
