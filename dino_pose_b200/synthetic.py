"""Synthetic batches of the shapes the reference's loaders produce (``data_loader/data_loader.py``: ImageNet-
normalised ``pixel_values`` [B,3,H,W], target heat-maps [B,K,48,48], key-points (x, y, visibility) [B,K,3],
z targets [B,K]).  Used by bench.py / tools / tests; seeded on the CPU so every box sees the same data."""
from __future__ import annotations

import torch


def make_inputs(batch, height=224, width=224, seed=0, num_keypoints=24, heatmap_size=48):
    """N(0,1) pixels, U(0,1) target maps, visibility in {0,1,2} (the loss mask is ``> 1``, reference
    train.py:94,114), N(0,1) z targets."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    px = torch.randn(batch, 3, height, width, generator=g)
    hm = torch.rand(batch, num_keypoints, heatmap_size, heatmap_size, generator=g)
    xy = torch.rand(batch, num_keypoints, 2, generator=g) * height
    vis = torch.randint(0, 3, (batch, num_keypoints, 1), generator=g).float()
    kps = torch.cat([xy, vis], dim=-1)
    z = torch.randn(batch, num_keypoints, generator=g)
    return {"pixel_values": px, "heatmaps": hm, "keypoints": kps, "z": z}
