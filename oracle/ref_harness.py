"""Import the REAL reference (``/root/reference``) in the build container (TEST INFRASTRUCTURE).

Used only by ``make_golden.py`` (and by ``tests/test_reference_live.py`` when the
reference tree is present).  Nothing here runs on the GPU box: ``/root/reference``
does not exist there.

What is patched, and why:
* ``Dinov2Model.from_pretrained`` / ``AutoImageProcessor.from_pretrained`` need the hub
  (no network) -> replaced by ``Dinov2Model(Dinov2Config(...))`` with the hub configs'
  sizes (image_size 518 so ``position_embeddings`` is [1,1370,D]) and a stub processor.
* reference ``LoRAAttention.forward`` (model/lora.py:53-65) uses the transformers-4.x
  calling convention (positional ``head_mask, output_attentions``; tuple result); with the
  installed transformers 5.5.0 ``Dinov2Attention.forward(hidden_states, **kw)`` returns a
  tensor, so the unmodified method raises ``TypeError``.  The shim below keeps the
  arithmetic of lora.py:57-59 (``out + lora_output(out)``) and only adapts the calling
  convention.
"""
from __future__ import annotations

import sys
import types

REF_ROOT = "/root/reference"


def reference_available():
    import os
    return os.path.isdir(REF_ROOT + "/model")


def import_reference():
    """Returns (dinov2_pose module, lora module, pose_heads module)."""
    import transformers  # noqa: F401  (must be imported before any stubbing)
    from transformers import Dinov2Config, Dinov2Model
    from .weights import ARCHS

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import model.dinov2_pose as dp
    import model.lora as lora
    import model.pose_heads as ph

    def _from_pretrained(name, *a, **k):
        D, L, h = ARCHS[name]
        cfg = Dinov2Config(image_size=518, patch_size=14, hidden_size=D, num_hidden_layers=L,
                           num_attention_heads=h)
        return Dinov2Model(cfg)

    class _Proc:
        crop_size = {"height": 224, "width": 224}

    dp.Dinov2Model.from_pretrained = staticmethod(_from_pretrained)
    dp.AutoImageProcessor.from_pretrained = staticmethod(lambda *a, **k: _Proc())

    def _lora_fwd(self, hidden_states, **kw):
        out = self.original_attention(hidden_states, **kw)
        return out + self.lora_output(out)

    lora.LoRAAttention.forward = _lora_fwd
    return dp, lora, ph


def import_reference_decode():
    """reference src/model_utils.py decode functions, run verbatim (numpy)."""
    import transformers  # noqa: F401
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ("pycocotools", "pycocotools.coco"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.COCO = object
            sys.modules[name] = m
    if "model.model_utils" not in sys.modules:
        m = types.ModuleType("model.model_utils")
        m.resolve_model_name = lambda x: x
        sys.modules["model.model_utils"] = m
    import src.model_utils as smu
    return smu


def import_reference_losses():
    """keypoint_loss / z_loss / DynamicLossWeighting from train.py:17-120 (pure torch code;
    the module itself cannot be imported: matplotlib / pycocotools / timm are missing)."""
    import torch
    import torch.nn as nn
    with open(REF_ROOT + "/train.py") as f:
        lines = f.readlines()
    ns = {"torch": torch, "nn": nn}
    exec("".join(lines[16:120]), ns)
    return ns["keypoint_loss"], ns["z_loss"], ns["DynamicLossWeighting"]
