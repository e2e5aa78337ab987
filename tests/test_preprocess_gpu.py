"""GPU pre-processing (dp_preprocess_u8 through the C ABI / GpuBitImageProcessor) against the oracle and the frozen
outputs of the real BitImageProcessor: integer + fp32 arithmetic, so the bar is BIT-EXACT."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc():
    from dino_pose_b200.preprocess import GpuBitImageProcessor
    return GpuBitImageProcessor()


def test_matches_bit_image_processor_golden(proc, golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for i, (h, w, seed) in enumerate(g["cases"]):
        img = po.synthetic_image(int(h), int(w), int(seed))
        pv = proc(img, return_tensors="pt").pixel_values
        assert pv.is_cuda and pv.shape == (1, 3, 224, 224) and pv.dtype == torch.float32
        out = pv[0].cpu().numpy()
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g["sha256"][i]), f"case {i} ({h}x{w})"
        key = f"pixel_values_{i}"
        if key in g.files:
            assert np.array_equal(out.view(np.uint32), g[key].view(np.uint32))


@pytest.mark.parametrize("h,w", [(231, 517), (719, 403), (256, 300), (3000, 2000), (90, 90), (225, 224)])
def test_matches_oracle_on_random_images(proc, h, w):
    img = np.random.default_rng(h * 7919 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = proc(img)["pixel_values"][0].cpu().numpy()
    ref = po.preprocess(img)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_batch_pil_and_mixed_sizes(proc):
    from PIL import Image
    a = po.synthetic_image(300, 400, 11)
    b = po.synthetic_image(300, 400, 12)
    c = po.synthetic_image(500, 333, 13)
    out = proc([Image.fromarray(a), b, torch.from_numpy(c).permute(2, 0, 1)], return_tensors="pt").to("cuda:0").pixel_values
    assert out.shape == (3, 3, 224, 224)
    for k, im in enumerate((a, b, c)):
        assert np.array_equal(out[k].cpu().numpy().view(np.uint32), po.preprocess(im).view(np.uint32)), k
    # a same-size stack goes through one launch (B = 2)
    st = proc(np.stack([a, b]))["pixel_values"]
    assert torch.equal(st, out[:2])


def test_model_exposes_the_gpu_processor_and_runs_end_to_end():
    """demo.py:166-176 call sequence on the drop-in model: processor -> pixel_values -> forward -> decode."""
    from dino_pose_b200.model import Dinov2PoseModel
    from dino_pose_b200.src.model_utils import get_keypoints_from_heatmaps
    torch.manual_seed(0)
    model = Dinov2PoseModel(num_keypoints=24, backbone="facebook/dinov2-small").to("cuda:0").eval()
    assert model.image_processor.crop_size["width"] == 224
    inputs = model.image_processor(po.synthetic_image(480, 640, 3), return_tensors="pt").to("cuda:0")
    with torch.no_grad():
        heatmaps, depths = model(inputs.pixel_values)
    assert heatmaps.shape == (1, 24, 48, 48) and depths.shape == (1, 24)
    kps = get_keypoints_from_heatmaps(heatmaps.squeeze().cpu().numpy())
    assert len(kps) == 24


def test_rejects_bad_input(proc):
    with pytest.raises(ValueError):
        proc(np.zeros((10, 10, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        proc(np.zeros((100, 100, 3), dtype=np.uint8)[:, :, :2])
