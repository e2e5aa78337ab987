"""DINOv2 pose models (mirrors reference model/dinov2_pose.py): same constructor arguments, attributes,
``forward(pixel_values) -> (heatmaps, z_coords)`` and ``state_dict`` keys; compute on the sm_100a engine.
"""
from __future__ import annotations

import warnings
from types import SimpleNamespace
from typing import Any, Dict

import torch
import torch.nn.functional as F

from .base_pose import BasePoseModel
from .dinov2_backbone import Dinov2Model
from .lora import LoRAAttention
from .pose_heads import SpatialAwarePoseHeads

_Z_CONFIG = {"hidden_dims": (1024, 512, 256), "dropout_rate": 0.1}   # reference dinov2_pose.py:50-53


def _image_processor(name):
    """The reference's ``AutoImageProcessor.from_pretrained`` (dinov2_pose.py:15,182) resolves to HF BitImageProcessor
    with the DINOv2 preprocessor config; callers call it on PIL images / frames and read ``crop_size``
    (demo.py:80,171,281; data_loader.py:52,137).  Same call form, same values, computed on the GPU."""
    from ..preprocess import GpuBitImageProcessor
    return GpuBitImageProcessor()


class _PoseFunction(torch.autograd.Function):
    """Autograd boundary: one fused forward program, one fused backward program."""

    @staticmethod
    def forward(ctx, model, pixel_values, *trainable):
        eng = model._get_engine(pixel_values.device)
        plan = eng.forward(pixel_values, training=model.engine_mode())
        ctx.eng, ctx.plan, ctx.names = eng, plan, model._trainable_names
        # The saved activations of this graph are the plan's STATIC buffers.  A later train-mode forward with the same
        # (batch, H, W) overwrites them; the stamp lets backward() detect that instead of returning wrong gradients.
        ctx.generation = plan["generation"]
        ctx.mark_non_differentiable()
        return plan["t"]["hm"].clone(), plan["t"]["z"].clone()

    @staticmethod
    def backward(ctx, dhm, dz):
        if ctx.plan["generation"] != ctx.generation or ctx.plan.get("consumed") == ctx.generation:
            raise RuntimeError(
                "dino_pose_b200: backward() of a forward pass whose saved activations are gone -- the model keeps ONE set "
                "of activation buffers per (batch, height, width), so each train-mode forward must be followed by its "
                "backward before the next forward of the same shape, and a graph can be back-propagated once "
                "(no retain_graph double backward).  Run forward/backward pairs in order, or use a different batch size "
                "for the second view.")
        ctx.plan["consumed"] = ctx.generation
        grads = ctx.eng.backward(ctx.plan, dhm, dz)
        return (None, None) + tuple(grads[n].clone() for n in ctx.names)


class _Dinov2PoseBase(BasePoseModel):
    lora_config = None

    def _init_common(self, num_keypoints, backbone, heatmap_size):
        self.backbone = Dinov2Model.from_pretrained(backbone)
        self.backbone_name = backbone
        self.image_processor = _image_processor(backbone)
        self.heatmap_size, self.num_keypoints = heatmap_size, num_keypoints
        self._coreml_patch_applied = False
        for p in self.backbone.parameters():
            p.requires_grad = False

    def _init_heads(self, num_keypoints, heatmap_size):
        self.feat_dim = self.backbone.config.hidden_size
        self.pose_heads = SpatialAwarePoseHeads(feat_channels=self.feat_dim, num_keypoints=num_keypoints,
                                                heatmap_size=heatmap_size, spatial_input_size=16,
                                                z_coord_config=dict(_Z_CONFIG))
        self._engine = None
        self._backend_factory = None

    # ---- engine plumbing
    def _get_engine(self, device):
        if self._engine is not None and self._engine.device == device:
            return self._engine
        from ..engine import PoseEngine
        if self._backend_factory is not None:
            backend = self._backend_factory()
        else:
            if device.type != "cuda":
                raise RuntimeError("dino_pose_b200 runs on CUDA (sm_100a) only: move the model and the input to a "
                                   "B200 (`model.cuda()`); there is no CPU execution path")
            from ..backend import CudaBackend
            backend = CudaBackend()
        cfg = self.backbone.config
        lora = None
        if self.lora_config is not None:
            lora = {"rank": self.lora_config["rank"], "alpha": self.lora_config["alpha"],
                    "dropout": self._lora_layer().dropout.p}
        zh = self.pose_heads.z_head
        ecfg = dict(D=cfg.hidden_size, L=cfg.num_hidden_layers, heads=cfg.num_attention_heads,
                    num_keypoints=self.num_keypoints, heatmap_size=self.heatmap_size, lora=lora,
                    z_hidden=zh.hidden_dims, z_dropout=zh.mlp[2].p, unfreeze=getattr(self, "unfreeze_last_n_layers", 0))
        if getattr(self, "_act_dtype", None) is not None:
            ecfg["act_dtype"] = self._act_dtype
        if getattr(self, "_raw_dtype", None) is not None:
            ecfg["raw_dtype"] = self._raw_dtype
        self._engine = PoseEngine(dict(self.named_parameters()), dict(self.named_buffers()), ecfg, backend, device)
        return self._engine

    def _lora_layer(self):
        return self.backbone.encoder.layer[-1].attention.lora_output

    def _apply(self, fn, *a, **k):
        # .to()/.cuda() replace parameter storage: plans built on the old tensors are stale
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _sync_dropout_config(self):
        if self._engine is None:
            return
        e = self._engine
        zp = self.pose_heads.z_head.mlp[2].p
        lp = self._lora_layer().dropout.p if self.lora_config is not None else 0.0
        if e.cfg.get("z_dropout") != zp or (e.lora is not None and e.lora.get("dropout") != lp):
            self._engine = None    # probabilities are baked into the recorded programs

    def engine_mode(self):
        """False: inference program.  True: training step.  2: training step with `model.pose_heads.eval()` -- the
        reference's nn.Module semantics for that call: BatchNorm2d of the heads normalises with its running statistics
        and does not update them, the z-head's Dropout is the identity, gradients flow as usual."""
        if not self.training:
            return False
        return True if self.pose_heads.training else 2

    def forward(self, pixel_values):
        """reference model/dinov2_pose.py:143-157 / :292-306."""
        self._sync_dropout_config()
        self._trainable_names = [n for n, p in self.named_parameters() if p.requires_grad]
        needs_grad = torch.is_grad_enabled() and self.training and bool(self._trainable_names)
        if not self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # eval-mode forward: the inference program (folded BatchNorm, merged LoRA) records no graph.  The reference is
            # differentiable here; say so loudly instead of silently returning detached outputs.
            needs_grad = False
            if not getattr(self, "_warned_eval_grad", False):
                warnings.warn("dino_pose_b200: forward() in eval mode with autograd enabled returns outputs WITHOUT a graph "
                              "(the inference program is not differentiable); wrap inference in torch.no_grad() or call "
                              "model.train() for a differentiable forward", RuntimeWarning, stacklevel=2)
                self._warned_eval_grad = True
        if needs_grad:
            params = [p for p in self.parameters() if p.requires_grad]
            return _PoseFunction.apply(self, pixel_values, *params)
        eng = self._get_engine(pixel_values.device)
        plan = eng.forward(pixel_values, training=self.engine_mode())
        return plan["t"]["hm"].clone(), plan["t"]["z"].clone()

    # ---- CoreML export hooks (reference :56-131, :221-278) -- export-time only, not on the CUDA path
    def apply_coreml_compatibility_patch(self):
        if self._coreml_patch_applied:
            print("Core ML patch already applied")
            return
        emb = self.backbone.embeddings

        def nearest_pos_encoding(embeddings, width, height):
            pos = emb.position_embeddings
            n_pos = pos.shape[1] - 1
            if embeddings.shape[1] - 1 == n_pos and height == width:
                return pos
            d = embeddings.shape[-1]
            h0, w0 = height // emb.patch_size + 0.1, width // emb.patch_size + 0.1
            side = int(n_pos ** 0.5)
            grid = pos[:, 1:].reshape(1, side, side, d).permute(0, 3, 1, 2)
            grid = F.interpolate(grid, size=(int(h0), int(w0)), mode="nearest")
            return torch.cat((pos[:, 0].unsqueeze(0), grid.permute(0, 2, 3, 1).view(1, -1, d)), dim=1)

        self._orig_interpolate = emb.interpolate_pos_encoding
        emb.interpolate_pos_encoding = nearest_pos_encoding
        self._coreml_patch_applied = True
        print("Core ML compatibility patch applied (bicubic -> nearest position-embedding resize)")

    def remove_coreml_compatibility_patch(self):
        if not self._coreml_patch_applied:
            print("No Core ML patch to remove")
            return
        self.backbone.embeddings.interpolate_pos_encoding = self._orig_interpolate
        self._coreml_patch_applied = False
        print("Core ML patch removed")


class Dinov2PoseModel(_Dinov2PoseBase):
    """Frozen DINOv2 backbone + trainable pose heads (reference model/dinov2_pose.py:10-174)."""

    def __init__(self, num_keypoints=24, backbone="facebook/dinov2-base", unfreeze_last_n_layers=0, heatmap_size=48):
        super().__init__()
        self._init_common(num_keypoints, backbone, heatmap_size)
        # reference :25-39: every parameter of the last n encoder layers (attention, MLP, LayerScale, both LayerNorms)
        layers = self.backbone.encoder.layer
        self.unfreeze_last_n_layers = max(0, min(int(unfreeze_last_n_layers), len(layers)))
        for i in range(1, self.unfreeze_last_n_layers + 1):
            for p in layers[len(layers) - i].parameters():
                p.requires_grad = True
        self._init_heads(num_keypoints, heatmap_size)

    @classmethod
    def from_config(cls, model_name: str, config: Dict[str, Any]):
        return cls(num_keypoints=config["num_keypoints"], backbone=model_name,
                   unfreeze_last_n_layers=config.get("unfreeze_last_n_layers", 0),
                   heatmap_size=config["output_heatmap_size"])


class Dinov2PoseModelLoRA(_Dinov2PoseBase):
    """DINOv2 backbone with a LoRA adapter on the LAST block's attention output + trainable heads
    (reference model/dinov2_pose.py:176-348; adapter placement :197-204)."""

    def __init__(self, num_keypoints=24, backbone="facebook/dinov2-base", heatmap_size=48, lora_rank=8, lora_alpha=16,
                 lora_dropout=0.1):
        super().__init__()
        self._init_common(num_keypoints, backbone, heatmap_size)
        self.lora_config = {"rank": lora_rank, "alpha": lora_alpha, "dropout": lora_dropout}
        layers = self.backbone.encoder.layer
        for i, layer in enumerate(layers):
            if i >= len(layers) - 1:
                layer.attention = LoRAAttention(layer.attention, r=lora_rank, alpha=lora_alpha, dropout=lora_dropout)
        self._init_heads(num_keypoints, heatmap_size)

    @classmethod
    def from_config(cls, model_name: str, config: Dict[str, Any]):
        return cls(num_keypoints=config["num_keypoints"], backbone=model_name,
                   heatmap_size=config["output_heatmap_size"], lora_rank=config.get("lora_rank", 8),
                   lora_alpha=config.get("lora_alpha", 16), lora_dropout=config.get("lora_dropout", 0.1))

    def apply_loading_fixes(self):
        """reference :325-348: re-sync alpha / rank / dropout.p with ``lora_config`` and force eval mode."""
        for _, m in self.named_modules():
            if hasattr(m, "alpha") and hasattr(m, "rank"):
                m.alpha, m.rank = self.lora_config["alpha"], self.lora_config["rank"]
            if hasattr(m, "dropout") and hasattr(m.dropout, "p"):
                if abs(m.dropout.p - self.lora_config["dropout"]) > 1e-6:
                    m.dropout.p = self.lora_config["dropout"]
        self.eval()
        self._engine = None
