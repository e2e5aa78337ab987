"""Stand-alone entry points for sub-modules of the drop-in surface (called by the containers' ``forward``)."""
from __future__ import annotations

import torch


def _cuda_only(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: dino_pose_b200 has no CPU execution path; pass CUDA tensors")


def lora_delta(x, lora_A, lora_B, scaling, p_drop):
    """reference model/lora.py:26-28 on an arbitrary [..., in] tensor via ``dp_lora_fwd``
    (x_in = 0, lambda1 = 1, then subtract y): returns dropout(x @ A @ B) * scaling."""
    _cuda_only(x, "LoRALayer.forward")
    from .backend import CudaBackend
    be = CudaBackend()
    prog = be.begin()
    D = x.shape[-1]
    if lora_B.shape[1] != D:
        raise NotImplementedError("LoRALayer kernel path supports in_features == out_features (the reference's only use)")
    y = x.reshape(-1, D).float().contiguous()
    rows = y.shape[0]
    ones = torch.ones(D, device=x.device)
    zeros = torch.zeros_like(y)
    out = torch.empty_like(y)
    seed = torch.randint(0, 2 ** 62, (1,), device=x.device, dtype=torch.int64) if p_drop > 0 else None
    be.lora_fwd(y, lora_A.detach().float().contiguous(), lora_B.detach().float().contiguous(), ones, zeros, out, None,
                rows=rows, D=D, R=lora_A.shape[1], scaling=float(scaling), p_drop=float(p_drop), seed=seed)
    prog.run()
    return (out - y).reshape(x.shape).to(x.dtype)


def run_backbone(backbone, pixel_values):
    raise NotImplementedError("stand-alone Dinov2Model.forward: use the pose model forward (fused path)")


def run_head_module(module, kind, x):
    raise NotImplementedError(f"stand-alone {type(module).__name__}.forward is not wired yet; the heads run inside "
                              "Dinov2PoseModel / Dinov2PoseModelLoRA.forward (fused path)")
