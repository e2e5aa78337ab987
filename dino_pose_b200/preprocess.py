"""GPU image processor: drop-in for the reference's ``model.image_processor``.

The reference sets ``self.image_processor = AutoImageProcessor.from_pretrained(backbone)`` (model/dinov2_pose.py:15,182)
and its callers do ``image_processor(image, return_tensors="pt").to(device)`` then read ``.pixel_values``
(demo.py:80,171; benchmark_model.py:35,45; data_loader/data_loader.py:52) and ``.crop_size['width' / 'height']``
(data_loader.py:137,140).  For the DINOv2 checkpoints that object is HF ``BitImageProcessor`` with: shortest edge 256,
bicubic, center crop 224, rescale 1/255, ImageNet mean / std.  ``GpuBitImageProcessor`` keeps that call form and those
attributes and computes bit-identical ``pixel_values`` on the device (``dp_preprocess_u8``, csrc/preprocess.cu): the only
host work is the H2D copy of the uint8 image.  There is no CPU path -- it raises without a CUDA device / the extension.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


class PixelBatch(dict):
    """Minimal ``BatchFeature``: ``["pixel_values"]``, ``.pixel_values`` and ``.to(device)`` (what the callers use)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, *args, **kwargs):
        return PixelBatch({k: v.to(*args, **kwargs) for k, v in self.items()})


class GpuBitImageProcessor:
    def __init__(self, size=None, crop_size=None, image_mean=_MEAN, image_std=_STD, rescale_factor=0.00392156862745098,
                 device=None):
        self.do_resize = self.do_center_crop = self.do_rescale = self.do_normalize = self.do_convert_rgb = True
        self.size = dict(size or {"shortest_edge": 256})
        self.crop_size = dict(crop_size or {"height": 224, "width": 224})
        self.resample = 3                               # PIL.Image.BICUBIC
        self.image_mean, self.image_std = list(image_mean), list(image_std)
        self.rescale_factor = rescale_factor
        self.device = device
        if self.crop_size["height"] != self.crop_size["width"]:
            raise ValueError("GpuBitImageProcessor supports square crops (the DINOv2 preprocessor config)")
        # TorchvisionBackend._fuse_mean_std_and_rescale_factor: float32 tensor * python scalar (a float32 product)
        k = np.float32(1.0 / rescale_factor)
        self._mean255 = (C.c_float * 3)(*(np.asarray(image_mean, np.float32) * k).tolist())
        self._std255 = (C.c_float * 3)(*(np.asarray(image_std, np.float32) * k).tolist())
        self._ws = None

    # ------------------------------------------------------------------ input normalisation
    def _to_u8_hwc(self, image, dev):
        """PIL image / numpy HWC (or CHW) uint8 / torch uint8 tensor -> contiguous uint8 [H, W, 3] on the device."""
        if hasattr(image, "convert") and hasattr(image, "size") and not isinstance(image, (np.ndarray, torch.Tensor)):
            image = np.asarray(image.convert("RGB"))     # do_convert_rgb
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if not isinstance(image, torch.Tensor):
            raise ValueError(f"Unsupported input image type {type(image)}")
        if image.dtype != torch.uint8:
            raise ValueError("GpuBitImageProcessor takes uint8 images (the reference passes PIL images / uint8 frames)")
        if image.ndim != 3:
            raise ValueError(f"expected a 3-D image, got shape {tuple(image.shape)}")
        if image.shape[-1] != 3 and image.shape[0] == 3:   # channels first (infer_channel_dimension_format)
            image = image.permute(1, 2, 0)
        if image.shape[-1] != 3:
            raise ValueError(f"expected 3 channels, got shape {tuple(image.shape)}")
        return image.to(dev, non_blocking=True).contiguous()

    def _device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("GpuBitImageProcessor needs a CUDA device: dino_pose_b200 has no CPU execution path")
        return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

    def preprocess_batch(self, images_u8, out=None):
        """images_u8: device uint8 [B, H, W, 3] -> fp32 [B, 3, crop, crop] (one resize geometry for the batch)."""
        lib = _lib.lib()
        B, H, W, _ = images_u8.shape
        crop, se = self.crop_size["height"], self.size["shortest_edge"]
        need = lib.dp_preprocess_workspace_bytes(B, H, W, se, crop)
        if need < 0:
            raise ValueError(f"unsupported pre-processing geometry: image {H}x{W}, shortest edge {se}, crop {crop}")
        if self._ws is None or self._ws.numel() < need or self._ws.device != images_u8.device:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=images_u8.device)
        if out is None:
            out = torch.empty(B, 3, crop, crop, dtype=torch.float32, device=images_u8.device)
        rc = lib.dp_preprocess_u8(images_u8.data_ptr(), B, H, W, se, crop, self._mean255, self._std255, out.data_ptr(),
                                  self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream(images_u8.device).cuda_stream)
        _lib.check(rc, "dp_preprocess_u8")
        return out

    def __call__(self, images=None, return_tensors="pt", **kwargs):
        if images is None:
            images = kwargs.pop("image", None)
        if return_tensors not in ("pt", None):
            raise ValueError("GpuBitImageProcessor returns torch tensors (return_tensors='pt')")
        single = not isinstance(images, (list, tuple)) and not (isinstance(images, (np.ndarray, torch.Tensor)) and images.ndim == 4)
        items = [images] if single else list(images)
        dev = self._device()
        with torch.cuda.device(dev):
            u8 = [self._to_u8_hwc(im, dev) for im in items]
            crop = self.crop_size["height"]
            out = torch.empty(len(u8), 3, crop, crop, dtype=torch.float32, device=dev)
            # group_images_by_shape: images of one size share a launch
            groups = {}
            for idx, t in enumerate(u8):
                groups.setdefault(tuple(t.shape), []).append(idx)
            for shape, idxs in groups.items():
                stacked = u8[idxs[0]].unsqueeze(0) if len(idxs) == 1 else torch.stack([u8[i] for i in idxs])
                res = self.preprocess_batch(stacked)
                if len(idxs) == len(u8):
                    out = res
                else:
                    out[torch.as_tensor(idxs, device=dev)] = res
        return PixelBatch(pixel_values=out)
