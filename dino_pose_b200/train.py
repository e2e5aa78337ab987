"""Fine-tuning step of the reference's ``train.py`` hot loop (:134-188) on the sm_100a engine, data-parallel.

One ``PoseTrainer.step`` = ``optimizer.zero_grad`` -> ``model(pixel_values)`` -> reference losses
(``keypoint_loss`` train.py:89-102, ``z_loss`` :109-120, ``DynamicLossWeighting`` :17-69) -> backward ->
gradient all-reduce (world_size > 1) -> AdamW(lr 3e-5, wd 1e-6; train.py:280-284).

Differences from the reference loop, all deliberate:
  * the loss-weight EMA state lives on the device -- the reference's 7 ``.item()`` host syncs per step
    (train.py:155-156,173-178) would cap multi-GPU scaling; the arithmetic is unchanged;
  * gradients of all trainable parameters live in ONE flat fp32 buffer written by the backward program, so
    the data-parallel exchange is a single bucketed NCCL all-reduce over NVLink of 31 MB (ViT-S);
  * BatchNorm uses per-replica batch statistics (torch DDP default); running statistics are per rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DeviceLossWeighting:
    """``DynamicLossWeighting`` (reference train.py:17-87) with tensor state (no host round trips)."""

    def __init__(self, device, initial_weight=0.1, adjustment_rate=0.1, momentum=0.9):
        self.weight = torch.tensor(float(initial_weight), device=device)
        self.kp_avg = torch.zeros((), device=device)
        self.z_avg = torch.zeros((), device=device)
        self.started = torch.zeros((), device=device)
        self.rate, self.momentum = adjustment_rate, momentum

    @torch.no_grad()
    def update(self, kp, z):
        m = self.momentum
        first = 1.0 - self.started
        self.kp_avg.copy_(first * kp + self.started * (m * self.kp_avg + (1 - m) * kp))
        self.z_avg.copy_(first * z + self.started * (m * self.z_avg + (1 - m) * z))
        self.started.fill_(1.0)
        target = (kp + 1e-8) / (z + 1e-8)
        self.weight.copy_(((1 - self.rate) * self.weight + self.rate * target).clamp(1e-3, 10.0))

    def balanced(self, kp_loss, z_loss):
        return kp_loss / (self.kp_avg + 1e-8) + z_loss / (self.z_avg + 1e-8)


def keypoint_loss(pred, target, conf):
    """reference train.py:89-102."""
    mask = (conf > 1).to(pred.dtype)[:, :, None, None]
    diff = (pred - target) ** 2
    return (torch.exp(-diff.detach()) * diff * mask).mean()


def z_loss(pred_z, target_z, conf):
    """reference train.py:109-120 (masked L1)."""
    mask = (conf > 1).to(pred_z.dtype)
    return (pred_z * mask - target_z * mask).abs().mean()


class PoseTrainer:
    def __init__(self, model, lr=3e-5, weight_decay=1e-6, bucket_mb=8.0):
        self.model = model
        self.device = next(model.parameters()).device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.weighting = DeviceLossWeighting(self.device)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.names = [n for n, p in model.named_parameters() if p.requires_grad]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, fused=self.device.type == "cuda")
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self._bound_plan = None
        self.comm_stream = torch.cuda.Stream(device=self.device) if (self.world > 1 and self.device.type == "cuda") else None

    def _bind_grads(self, plan):
        """Point every ``param.grad`` at its slice of the plan's flat gradient buffer (no copies)."""
        if self._bound_plan is plan:
            return
        for n, p in zip(self.names, self.params):
            p.grad = plan["grads"][n]
        self._bound_plan = plan

    def _allreduce(self, flat):
        n = flat.numel()
        handles = []
        for s in range(0, n, self.bucket_elems):
            handles.append(dist.all_reduce(flat[s:min(n, s + self.bucket_elems)], op=dist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()
        flat.mul_(1.0 / self.world)

    def step(self, pixel_values, target_heatmaps, keypoints, target_z):
        """One fine-tuning step on this rank's shard.  Returns (loss, kp_loss, z_loss) device scalars."""
        model = self.model
        model.train()
        eng = model._get_engine(self.device)
        plan = eng.forward(pixel_values, training=True)
        hm = plan["t"]["hm"].detach().requires_grad_(True)
        z = plan["t"]["z"].detach().requires_grad_(True)
        conf = keypoints[..., 2]
        kp = keypoint_loss(hm, target_heatmaps, conf)
        zl = z_loss(z, target_z, conf)
        self.weighting.update(kp.detach(), zl.detach())
        loss = self.weighting.balanced(kp, zl)
        dhm, dz = torch.autograd.grad(loss, (hm, z))
        eng.backward(plan, dhm, dz)
        if self.world > 1:
            self._allreduce(plan["gflat"])
        self._bind_grads(plan)
        self.opt.step()
        return loss.detach(), kp.detach(), zl.detach()
