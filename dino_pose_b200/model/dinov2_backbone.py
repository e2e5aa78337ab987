"""Parameter container for the DINOv2 ViT backbone with HuggingFace ``Dinov2Model`` naming.

The reference obtains its backbone from ``transformers.Dinov2Model.from_pretrained`` (model/dinov2_pose.py:13,180)
and reaches into ``.config.hidden_size``, ``.encoder.layer[i].attention / .norm1 / .norm2`` and
``.embeddings.{interpolate_pos_encoding, position_embeddings, patch_size}``.  This module keeps that
attribute tree and the ``state_dict`` keys, but holds parameters only -- the arithmetic of HF
modeling_dinov2.py:97-116,141-149,203-234,249-252,272-278,324-328,367-386,473-478 runs in the CUDA engine.
"""
from __future__ import annotations

import os
import warnings
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

HUB_CONFIGS = {
    # hub name: (hidden, layers, heads); image_size 518 / patch 14 => 1 + 37*37 position embeddings
    "facebook/dinov2-small": (384, 12, 6),
    "facebook/dinov2-base": (768, 12, 12),
    "facebook/dinov2-large": (1024, 24, 16),
    "test/dinov2-tiny": (128, 2, 2),
}


def make_config(name):
    if name not in HUB_CONFIGS:
        raise ValueError(f"unknown DINOv2 backbone {name!r}; known: {sorted(HUB_CONFIGS)}")
    d, l, h = HUB_CONFIGS[name]
    return SimpleNamespace(hidden_size=d, num_hidden_layers=l, num_attention_heads=h, patch_size=14, image_size=518,
                           num_channels=3, mlp_ratio=4, layer_norm_eps=1e-6, layerscale_value=1.0,
                           initializer_range=0.02, hidden_act="gelu", qkv_bias=True, use_swiglu_ffn=False,
                           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, drop_path_rate=0.0,
                           name_or_path=name)


class _Holder(nn.Module):
    """Module that only owns sub-modules / parameters; computing through it directly is not supported."""

    def forward(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is a parameter container; run the pose model (or "
                                  "Dinov2Model) forward, which executes the fused sm_100a path")


class PatchEmbeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.projection = nn.Conv2d(cfg.num_channels, cfg.hidden_size, cfg.patch_size, cfg.patch_size)
        self.patch_size = (cfg.patch_size, cfg.patch_size)
        self.num_channels = cfg.num_channels
        self.num_patches = (cfg.image_size // cfg.patch_size) ** 2


class Embeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.hidden_size
        self.cls_token = nn.Parameter(torch.empty(1, 1, d))
        self.mask_token = nn.Parameter(torch.zeros(1, d))
        self.patch_embeddings = PatchEmbeddings(cfg)
        self.position_embeddings = nn.Parameter(torch.empty(1, self.patch_embeddings.num_patches + 1, d))
        self.patch_size = cfg.patch_size

    def interpolate_pos_encoding(self, embeddings, height, width):
        """HF modeling_dinov2.py:57-95 (bicubic, align_corners=False, fp32)."""
        pos = self.position_embeddings
        n_pos = pos.shape[1] - 1
        if embeddings.shape[1] - 1 == n_pos and height == width:
            return pos
        d = embeddings.shape[-1]
        side = int(n_pos ** 0.5)
        grid = pos[:, 1:].reshape(1, side, side, d).permute(0, 3, 1, 2).float()
        grid = F.interpolate(grid, size=(height // self.patch_size, width // self.patch_size), mode="bicubic",
                             align_corners=False).to(pos.dtype)
        return torch.cat((pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, d)), dim=1)


class SelfAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.hidden_size
        self.query, self.key, self.value = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.num_attention_heads = cfg.num_attention_heads


class SelfOutput(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.hidden_size)


class Attention(_Holder):
    """HF ``Dinov2Attention`` (modeling_dinov2.py:237-252): the block ``LoRAAttention`` wraps (reference model/lora.py:53-65)."""

    def __init__(self, cfg):
        super().__init__()
        self.attention = SelfAttention(cfg)
        self.output = SelfOutput(cfg)

    def forward(self, hidden_states, head_mask=None, output_attentions=False):
        if head_mask is not None or output_attentions:
            raise NotImplementedError("head_mask / output_attentions are not supported by the fused attention kernel")
        from ..functional import run_attention
        return run_attention(self, hidden_states)


class LayerScale(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.lambda1 = nn.Parameter(cfg.layerscale_value * torch.ones(cfg.hidden_size))


class MLP(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.hidden_size
        self.fc1 = nn.Linear(d, int(d * cfg.mlp_ratio))
        self.fc2 = nn.Linear(int(d * cfg.mlp_ratio), d)


class Layer(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.hidden_size
        self.norm1 = nn.LayerNorm(d, eps=cfg.layer_norm_eps)
        self.attention = Attention(cfg)
        self.layer_scale1 = LayerScale(cfg)
        self.norm2 = nn.LayerNorm(d, eps=cfg.layer_norm_eps)
        self.mlp = MLP(cfg)
        self.layer_scale2 = LayerScale(cfg)


class Encoder(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList(Layer(cfg) for _ in range(cfg.num_hidden_layers))


class Dinov2Model(nn.Module):
    """``transformers.Dinov2Model`` look-alike: ``embeddings`` / ``encoder.layer[i]`` / ``layernorm`` + ``config``."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = Embeddings(config)
        self.encoder = Encoder(config)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """HF ``_init_weights`` (modeling_dinov2.py:406-422): trunc-normal(0.02) weights, zero biases,
        LayerNorm (1, 0), LayerScale = layerscale_value."""
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.trunc_normal_(m.weight, mean=0.0, std=std)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        nn.init.trunc_normal_(self.embeddings.position_embeddings, mean=0.0, std=std)
        nn.init.trunc_normal_(self.embeddings.cls_token, mean=0.0, std=std)

    @classmethod
    def from_pretrained(cls, name, *args, **kwargs):
        """``transformers.Dinov2Model.from_pretrained`` for the named architecture (reference model/dinov2_pose.py:13,180).

        The checkpoint is loaded through HuggingFace (cache or hub; ``local_files_only`` and the other keyword
        arguments are passed on, ``HF_HUB_OFFLINE`` is honoured by HF itself) with ``strict=True`` and every failure
        PROPAGATES: a fine-tuning run on a silently random backbone is meaningless.  Random initialisation
        (HF ``_init_weights``) is an explicit opt-in -- ``DINO_POSE_RANDOM_INIT=1`` in the environment, or a
        ``test/...`` architecture -- and is announced with a warning; the benchmark and the tests use it because this
        environment has no network and no checkpoints."""
        model = cls(make_config(name))
        if name.startswith("test/"):
            return model
        if os.environ.get("DINO_POSE_RANDOM_INIT", "0") not in ("", "0"):
            warnings.warn(f"Dinov2Model.from_pretrained({name!r}): DINO_POSE_RANDOM_INIT is set -- the backbone is RANDOMLY "
                          "initialised (no pretrained weights were loaded); results are only meaningful for benchmarking "
                          "and parity tests", RuntimeWarning, stacklevel=2)
            return model
        try:
            from transformers import Dinov2Model as HFModel
        except Exception as ex:   # no transformers: say what to do instead of training on noise
            raise RuntimeError(f"loading the pretrained backbone {name!r} needs `transformers` ({ex}); set "
                               "DINO_POSE_RANDOM_INIT=1 to opt in to a randomly initialised backbone") from ex
        try:
            hf = HFModel.from_pretrained(name, *args, **kwargs)
        except Exception as ex:
            raise RuntimeError(f"could not load the pretrained backbone {name!r} ({type(ex).__name__}: {ex}); set "
                               "DINO_POSE_RANDOM_INIT=1 to opt in to a randomly initialised backbone") from ex
        model.load_state_dict(hf.state_dict(), strict=True)
        return model

    def forward(self, pixel_values):
        from ..functional import run_backbone
        return run_backbone(self, pixel_values)
