"""LoRA adapters (mirrors reference model/lora.py).

``LoRALayer`` / ``LoRAAttention`` hold the trainable ``lora_A [in, r]`` / ``lora_B [r, out]`` parameters under
the reference's names.  Inside ``Dinov2PoseModelLoRA`` the adapter is executed by the fused kernel
``dp_lora_fwd`` (training) or folded into the output-projection weights (eval); the standalone ``forward``
of these modules runs the same kernel on the given tensor.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class LoRALayer(nn.Module):
    """reference model/lora.py:5-28: ``dropout(x @ A @ B) * (alpha / rank)``; A ~ kaiming-uniform(a=sqrt 5), B = 0."""

    def __init__(self, in_features, out_features, r=8, alpha=16, dropout=0.1):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.rank, self.alpha = r, alpha
        self.lora_A = nn.Parameter(torch.empty(in_features, r))
        self.lora_B = nn.Parameter(torch.zeros(r, out_features))
        self.dropout = nn.Dropout(dropout)
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))

    @property
    def scaling(self):
        return self.alpha / self.rank

    def forward(self, x):
        from ..functional import lora_delta
        return lora_delta(x, self.lora_A, self.lora_B, self.scaling, self.dropout.p if self.training else 0.0)


class LoRAAttention(nn.Module):
    """reference model/lora.py:31-65: wraps an attention block and adds a LoRA update to its OUTPUT."""

    def __init__(self, original_attention, r=8, alpha=16, dropout=0.1):
        super().__init__()
        self.original_attention = original_attention
        self.rank, self.alpha = r, alpha
        q = original_attention.attention.query
        self.in_dim, self.out_dim = q.in_features, q.out_features
        self.lora_output = LoRALayer(self.in_dim, self.in_dim, r, alpha, dropout)
        for p in self.original_attention.parameters():
            p.requires_grad = False

    @property
    def scaling(self):
        return self.alpha / self.rank

    def forward(self, hidden_states, head_mask=None, output_attentions=False):
        out = self.original_attention(hidden_states)
        out = out[0] if isinstance(out, tuple) else out
        return (out + self.lora_output(out),)


class ConvLoRA(nn.Module):
    """FastViT-only adapter (reference model/lora.py:68-121): out of scope of the DINOv2 hot path; kept
    importable because reference model/fastvit_pose.py imports it."""

    def __init__(self, original_conv, r=8, alpha=16, dropout=0.1):
        super().__init__()
        raise NotImplementedError("ConvLoRA belongs to the FastViT model family, which dino_pose_b200 does not cover")


class FastViTLoRA(nn.Module):
    """FastViT-only (reference model/lora.py:124-149); see ``ConvLoRA``."""

    def __init__(self, original_mlp, r=8, alpha=16, dropout=0.1):
        super().__init__()
        raise NotImplementedError("FastViTLoRA belongs to the FastViT model family, which dino_pose_b200 does not cover")
