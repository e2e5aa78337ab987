"""Heat-map -> key-point decode (mirrors the decode part of reference src/model_utils.py:10-51) on the
``dp_decode`` kernel.  Same function names, argument meaning and return types as the reference:
``argmax_ind`` -> (row, col, value); ``weighted_max_loc`` -> (x, y) floats; ``get_keypoints_from_heatmaps`` ->
list of (x, y); ``get_keypoints_from_heatmaps_batch`` -> float64 ndarray [B, K, 2].  Inputs may be CUDA tensors
(no host round trip of the maps) or numpy arrays / CPU tensors (copied to the device first).
"""
from __future__ import annotations

import numpy as np
import torch


def _to_device_maps(heatmaps):
    if isinstance(heatmaps, np.ndarray):
        heatmaps = torch.from_numpy(np.ascontiguousarray(heatmaps, dtype=np.float32))
    if not torch.cuda.is_available():
        raise RuntimeError("dino_pose_b200 decode runs on CUDA (sm_100a) only; there is no CPU execution path")
    return heatmaps.detach().to(device="cuda", dtype=torch.float32).contiguous()


def decode_heatmaps(heatmaps, target_size=(224, 224)):
    """heatmaps [..., H, W] -> (idx int32 [..., 2] = (row, col), xy float64 [..., 2] = (x, y), conf fp32 [...])
    as CUDA tensors.  Bit-exact against the reference numpy decode."""
    from ..backend import CudaBackend
    hm = _to_device_maps(heatmaps)
    lead, (H, W) = hm.shape[:-2], hm.shape[-2:]
    maps = int(np.prod(lead)) if lead else 1
    idx = torch.empty((maps, 2), dtype=torch.int32, device=hm.device)
    xy = torch.empty((maps, 2), dtype=torch.float64, device=hm.device)
    conf = torch.empty((maps,), dtype=torch.float32, device=hm.device)
    be = CudaBackend()
    prog = be.begin()
    be.decode(hm, idx, xy, conf, maps=maps, H=H, W=W, target_w=target_size[0], target_h=target_size[1])
    prog.run()
    return idx.view(*lead, 2), xy.view(*lead, 2), conf.view(*lead)


def argmax_ind(heatmap):
    idx, _, conf = decode_heatmaps(torch.as_tensor(heatmap).reshape((1,) + tuple(np.shape(heatmap)[-2:])))
    r, c = idx.view(-1).tolist()
    return r, c, np.float32(conf.item())


def weighted_max_loc(heatmap, target_size=(224, 224)):
    hm = torch.as_tensor(np.squeeze(heatmap) if isinstance(heatmap, np.ndarray) else heatmap.squeeze())
    _, xy, _ = decode_heatmaps(hm.reshape((1,) + tuple(hm.shape[-2:])), target_size)
    x, y = xy.view(-1).cpu().numpy()
    return x, y


def get_keypoints_from_heatmaps(heatmaps, target_size=(224, 224)):
    hm = torch.as_tensor(heatmaps).squeeze()
    _, xy, _ = decode_heatmaps(hm, target_size)
    arr = xy.cpu().numpy()
    return [(arr[k, 0], arr[k, 1]) for k in range(arr.shape[0])]


def get_keypoints_from_heatmaps_batch(heatmaps_batch, target_size=(224, 224)):
    _, xy, _ = decode_heatmaps(heatmaps_batch, target_size)
    return xy.cpu().numpy()
