"""Drop-in replacements for the reference's ``model/`` package (same class names, constructor arguments,
``forward`` signatures and ``state_dict`` keys); compute runs on the sm_100a kernels through ``engine.py``."""
from .base_pose import BasePoseModel
from .dinov2_pose import Dinov2PoseModel, Dinov2PoseModelLoRA
from .lora import LoRAAttention, LoRALayer
from .pose_heads import (HeatmapHead, HourglassModule, PoseHeads, SpatialAwareHeatmapHead, SpatialAwarePoseHeads,
                         ZCoordinateHead)

__all__ = ["BasePoseModel", "Dinov2PoseModel", "Dinov2PoseModelLoRA", "LoRALayer", "LoRAAttention", "HourglassModule",
           "SpatialAwareHeatmapHead", "SpatialAwarePoseHeads", "ZCoordinateHead", "HeatmapHead", "PoseHeads"]
