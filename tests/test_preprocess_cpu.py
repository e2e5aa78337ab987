"""Image pre-processing oracle (oracle/preprocess_oracle.py) pinned on the CPU:
  * against outputs of the real ``transformers.BitImageProcessor`` frozen in tests/golden/preprocess.npz
    (oracle/make_golden_preprocess.py), bit for bit;
  * against ATen's own uint8 anti-aliased bicubic kernel (``F.interpolate`` on the CPU, run live) on more sizes."""
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import preprocess_oracle as po


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


def test_oracle_matches_bit_image_processor_golden(golden):
    cases = golden["cases"]
    for i, (h, w, seed) in enumerate(cases):
        pv = po.preprocess(po.synthetic_image(int(h), int(w), int(seed)))
        assert pv.dtype == np.float32 and pv.shape == (3, po.CROP, po.CROP)
        assert hashlib.sha256(np.ascontiguousarray(pv).tobytes()).hexdigest() == str(golden["sha256"][i]), f"case {i} ({h}x{w})"
        key = f"pixel_values_{i}"
        if key in golden.files:
            assert np.array_equal(pv.view(np.uint32), golden[key].view(np.uint32))


@pytest.mark.parametrize("h,w,oh,ow", [(97, 131, 256, 345), (301, 203, 379, 256), (512, 512, 256, 256), (600, 900, 256, 384),
                                       (64, 48, 341, 256), (1200, 800, 384, 256)])
def test_resize_restatement_matches_aten_uint8_kernel(h, w, oh, ow):
    img = np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1)[None].contiguous()
    ref = F.interpolate(t, size=(oh, ow), mode="bicubic", antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
    assert np.array_equal(po.resize_u8(img, oh, ow), ref)


def test_geometry():
    assert po.resized_size(480, 640) == (256, 341)
    assert po.resized_size(640, 427) == (383, 256)
    assert po.resized_size(256, 256) == (256, 256)
    assert po.crop_origin(256, 341) == (16, 58)
    m, s = po.fused_mean_std()
    assert m.dtype == np.float32 and abs(float(m[0]) - 123.675) < 1e-3 and abs(float(s[2]) - 57.375) < 1e-3
