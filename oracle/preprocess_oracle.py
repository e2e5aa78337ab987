"""numpy restatement of the reference's image pre-processing (TEST INFRASTRUCTURE).

The reference builds ``self.image_processor = AutoImageProcessor.from_pretrained(backbone)``
(model/dinov2_pose.py:15,182) and calls it on PIL images / frames (demo.py:80,171, benchmark_model.py:35,45,
data_loader/data_loader.py:52).  For ``facebook/dinov2-*`` that is HF ``BitImageProcessor`` with the hub's
preprocessor_config.json: resize shortest edge 256 (bicubic) -> center crop 224 -> rescale 1/255 -> ImageNet mean / std.
The arithmetic lives in two third-party packages that are NOT vendored in the reference:

* ``transformers`` 5.5.0 (``requirements.txt:6`` asks for >=4.30.0, un-pinned) -- HF: is
  ``transformers/image_processing_backends.py`` (``TorchvisionBackend``), B: ``models/bit/image_processing_bit.py``,
  T: ``image_transforms.py``:
    - ``process_image``  HF  PIL -> uint8 CHW tensor (``pil_to_tensor``), RGB conversion
    - ``resize``         HF  ``get_resize_output_image_size(size=256, default_to_square=False)`` (T): short side -> 256,
                             long side -> ``int(256 * long / short)``; then ``tvF.resize(uint8, bicubic, antialias=True)``
    - ``center_crop``    HF  ``top = int((h - 224) / 2.0)``, ``left = int((w - 224) / 2.0)``
    - ``rescale_and_normalize`` + ``_fuse_mean_std_and_rescale_factor``  HF  mean' = float32(mean) * float32(255),
                             std' likewise; ``(float32(u8) - mean') / std'`` in float32 (``tvF.normalize``)
* ``torch`` 2.11.0 -- ``tvF.resize`` on a uint8 CPU tensor ends in ATen's uint8 anti-aliased bicubic kernel
  (``aten/src/ATen/native/cpu/UpSampleKernel.cpp``: ``_compute_indices_min_size_weights_aa``,
  ``_compute_index_ranges_int16_weights``, ``HelperInterpCubic::aa_filter`` with a = -0.5), restated here from its
  published algorithm: per output coordinate i
      scale = in / out;  support = 2 * max(scale, 1);  ksize = ceil(support) * 2 + 1
      center = scale * (i + 0.5);  xmin = max(int(center - support + 0.5), 0)
      xsize = clamp(min(int(center + support + 0.5), in) - xmin, 0, ksize)
      w_j = cubic((j + xmin - center + 0.5) / max(scale, 1)),  normalised by their sum          (float64)
  the weights of one axis are quantised to int16 with the largest shift (< 22) that keeps round(max_w * 2^(p+1))
  below 2^15; a pass computes ``clamp((2^(p-1) + sum_j u8 * w_j) >> p, 0, 255)``; the horizontal pass runs first and
  its result is ROUNDED TO uint8 before the vertical pass.

Pinned two ways (tests/test_preprocess_cpu.py): against ``torch.nn.functional.interpolate`` itself on random sizes
(run live, torch is on the GPU box too), and against tests/golden/preprocess.npz = outputs of the real
``transformers.BitImageProcessor`` (``oracle/make_golden_preprocess.py``).  Bit-exact in both.
"""
from __future__ import annotations

import math

import numpy as np

SHORT_EDGE = 256
CROP = 224
IMAGE_MEAN = (0.485, 0.456, 0.406)       # hub preprocessor_config.json of facebook/dinov2-{small,base,large}
IMAGE_STD = (0.229, 0.224, 0.225)
RESCALE_FACTOR = 0.00392156862745098


def cubic_aa(x, a=-0.5):
    """HelperInterpCubic::aa_filter<double, use_keys_cubic=true> (cubic_convolution1 / cubic_convolution2)."""
    x = abs(x)
    if x < 1.0:
        return ((a + 2) * x - (a + 3)) * x * x + 1
    if x < 2.0:
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return 0.0


def axis_weights(in_size, out_size):
    """int16 weights of one axis: (weights [out, ksize] int64, xmin [out], xsize [out], precision)."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    ksize = int(math.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    w = np.zeros((out_size, ksize), dtype=np.float64)
    xmins = np.zeros(out_size, dtype=np.int64)
    xsizes = np.zeros(out_size, dtype=np.int64)
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), in_size) - xmin
        xsize = min(max(xsize, 0), ksize)
        total = 0.0
        row = []
        for j in range(xsize):
            v = cubic_aa((j + xmin - center + 0.5) * invscale)
            row.append(v)
            total += v
        if total != 0.0:
            row = [v / total for v in row]
        w[i, :xsize] = row
        xmins[i], xsizes[i] = xmin, xsize
    max_w = float(w.max())
    prec = 0
    while prec < 22:
        if int(0.5 + max_w * (1 << (prec + 1))) >= (1 << 15):
            break
        prec += 1
    v = w * float(1 << prec)
    wi = np.where(v < 0, np.trunc(-0.5 + v), np.trunc(0.5 + v)).astype(np.int64)
    return wi, xmins, xsizes, prec


def resample_axis(img, out_size, axis):
    """one separable pass on a uint8 array; ``axis`` is resampled, the result is uint8 again."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    wi, xmins, xsizes, prec = axis_weights(src.shape[0], out_size)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.int64)
    for i in range(out_size):
        n = int(xsizes[i])
        taps = src[xmins[i]:xmins[i] + n]
        acc = np.tensordot(wi[i, :n], taps, axes=(0, 0)) + (1 << (prec - 1))
        out[i] = np.clip(acc >> prec, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_u8(img, out_h, out_w):
    """uint8 HWC -> uint8 [out_h, out_w, C]: horizontal pass, uint8 rounding, vertical pass."""
    t = resample_axis(img, out_w, 1)
    return resample_axis(t, out_h, 0)


def resized_size(h, w, short_edge=SHORT_EDGE):
    """get_resize_output_image_size(size=short_edge, default_to_square=False) -> (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = short_edge, int(short_edge * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def crop_origin(h, w, crop=CROP):
    return int((h - crop) / 2.0), int((w - crop) / 2.0)


def fused_mean_std(mean=IMAGE_MEAN, std=IMAGE_STD, rescale_factor=RESCALE_FACTOR):
    """_fuse_mean_std_and_rescale_factor: float32 tensors times the python scalar 1 / rescale_factor (float32 product)."""
    k = np.float32(1.0 / rescale_factor)
    return np.asarray(mean, np.float32) * k, np.asarray(std, np.float32) * k


def preprocess(img, short_edge=SHORT_EDGE, crop=CROP, mean=IMAGE_MEAN, std=IMAGE_STD):
    """uint8 HWC RGB image -> float32 [3, crop, crop] pixel_values (one image)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    nh, nw = resized_size(h, w, short_edge)
    r = resize_u8(img, nh, nw)
    top, left = crop_origin(nh, nw, crop)
    c = r[top:top + crop, left:left + crop]
    m, s = fused_mean_std(mean, std)
    x = c.astype(np.float32)
    return ((x - m) / s).transpose(2, 0, 1).copy()


def synthetic_image(h, w, seed):
    """Deterministic uint8 test image: smooth gradients + texture + noise (so that both flat and busy regions occur)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3), dtype=np.float64)
    for c in range(3):
        img[..., c] = 127.5 + 90.0 * np.sin(xx * (0.013 + 0.004 * c) + seed) * np.cos(yy * (0.011 + 0.003 * c)) \
            + 30.0 * np.sin((xx + yy) * 0.21 * (c + 1))
    img += rng.normal(0.0, 12.0, img.shape)
    img[: h // 8, : w // 8] = 255.0      # saturated corners exercise the clamp after the negative cubic lobes
    img[-(h // 8):, -(w // 8):] = 0.0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
