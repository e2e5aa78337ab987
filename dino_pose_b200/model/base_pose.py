"""Common interface of the pose models (mirrors reference model/base_pose.py:6-63)."""
from __future__ import annotations

import abc
from typing import Any, Dict

import torch.nn as nn


class BasePoseModel(nn.Module, abc.ABC):
    """Attributes every caller of the reference reads: ``num_keypoints``, ``heatmap_size``,
    ``backbone_name``, ``backbone`` (reference model/base_pose.py:12-17)."""

    def __init__(self):
        super().__init__()
        self.num_keypoints = self.heatmap_size = self.backbone_name = self.backbone = None

    @classmethod
    @abc.abstractmethod
    def from_config(cls, model_name: str, config: Dict[str, Any]):
        """Build from the reference's ``config_model`` dict (config/config.py:44-54)."""

    @abc.abstractmethod
    def forward(self, pixel_values):
        """pixel_values fp32 [B,3,H,W] -> (heatmaps [B,K,hm,hm], z_coords [B,K])."""

    def _params(self, trainable_only):
        return (p for p in self.parameters() if p.requires_grad or not trainable_only)

    def count_parameters(self, trainable_only=True):
        return sum(p.numel() for p in self._params(trainable_only))

    def print_trainable_parameters(self):
        for name, p in self.named_parameters():
            if p.requires_grad:
                print(f"Trainable: {name}, Shape: {p.shape}, Parameters: {p.numel():,}")
