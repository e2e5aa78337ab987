"""Freeze outputs of the REAL ``transformers.BitImageProcessor`` (DINOv2 preprocessor config) into
``tests/golden/preprocess.npz`` (TEST INFRASTRUCTURE; run in the build container):

    python -m oracle.make_golden_preprocess

The reference obtains this processor with ``AutoImageProcessor.from_pretrained("facebook/dinov2-*")``
(model/dinov2_pose.py:15,182); the hub is unreachable here, so the processor is constructed from the values of
that repository's preprocessor_config.json (crop 224, shortest edge 256, bicubic, 1/255, ImageNet mean / std).
Inputs are regenerated from their (h, w, seed) by ``preprocess_oracle.synthetic_image``; full fp32 outputs are
stored for three cases, a SHA-256 of the output bytes for all of them.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np
from PIL import Image

from . import preprocess_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "preprocess.npz")
# (h, w, seed): landscape / portrait / square / up-sampling / large down-sampling / odd sizes / already 256 / video frame
CASES = [(480, 640, 1), (640, 427, 2), (256, 256, 3), (100, 150, 4), (1080, 1920, 5), (333, 500, 6), (224, 224, 7),
         (720, 1280, 8), (37, 53, 9), (257, 511, 10)]
FULL = (0, 3, 5)


def main():
    from transformers import BitImageProcessor
    proc = BitImageProcessor(do_resize=True, size={"shortest_edge": po.SHORT_EDGE}, resample=3, do_center_crop=True,
                             crop_size={"height": po.CROP, "width": po.CROP}, do_rescale=True,
                             rescale_factor=po.RESCALE_FACTOR, do_normalize=True, image_mean=list(po.IMAGE_MEAN),
                             image_std=list(po.IMAGE_STD), do_convert_rgb=True)
    out = {"cases": np.asarray(CASES, dtype=np.int64)}
    digests = []
    for i, (h, w, seed) in enumerate(CASES):
        img = po.synthetic_image(h, w, seed)
        pv = proc(Image.fromarray(img), return_tensors="pt")["pixel_values"][0].numpy()   # demo.py:171 call form
        assert pv.shape == (3, po.CROP, po.CROP) and pv.dtype == np.float32
        mine = po.preprocess(img)
        assert np.array_equal(mine.view(np.uint32), pv.view(np.uint32)), f"oracle != BitImageProcessor for case {i}"
        digests.append(hashlib.sha256(np.ascontiguousarray(pv).tobytes()).hexdigest())
        if i in FULL:
            out[f"pixel_values_{i}"] = pv
        print(f"case {i} {h}x{w}: ok  sha256 {digests[-1][:16]}")
    out["sha256"] = np.asarray(digests)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
