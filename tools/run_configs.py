#!/usr/bin/env python
"""Run every BASELINE.json configuration at its full size on one GPU and print images/s (CUDA events, synthetic data).
configs[1] is the benchmark (bench.py); the others are reported in profiles/ as supporting numbers."""
import json
import os

# random-init weights of the named architecture (BASELINE.json north_star: no network, no checkpoints): explicit opt-in
os.environ.setdefault("DINO_POSE_RANDOM_INIT", "1")
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dino_pose_b200.model import Dinov2PoseModel, Dinov2PoseModelLoRA  # noqa: E402
from dino_pose_b200.src.model_utils import decode_heatmaps               # noqa: E402
from dino_pose_b200.synthetic import make_inputs                        # noqa: E402
from dino_pose_b200.train import PoseTrainer                            # noqa: E402

dev = torch.device("cuda:0")
# name, arch, lora, batch, resolution, mode, algorithmic GFLOP per image
CONFIGS = [
    ("cfg0 S frozen infer b1 224 (+decode)", "facebook/dinov2-small", False, 1, 224, "infer", 16.04),
    ("cfg1 S LoRA train b64 224", "facebook/dinov2-small", True, 64, 224, "train", 24.3),
    ("cfg2 B LoRA train b128 224", "facebook/dinov2-base", True, 128, 224, "train", 62.9),
    ("cfg3 L frozen infer b256 224", "facebook/dinov2-large", False, 256, 224, "infer", 167.33),
    ("cfg4 S LoRA infer b64 448", "facebook/dinov2-small", True, 64, 448, "infer", 78.64),
    ("cfg4 S LoRA train b64 448", "facebook/dinov2-small", True, 64, 448, "train", 111.7),
    # Dinov2PoseModel(unfreeze_last_n_layers=4) (reference config.py:48): fwd 16.04 + heads bwd 7.6 + 4 layers x
    # (2 x 0.9095 GEMM + 2.5 x 0.1015 attention) GF
    ("f4 S unfreeze4 train b64 224", "facebook/dinov2-small", False, 64, 224, "train", 31.9, 4),
]


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    only = sys.argv[1:] or None
    for name, arch, lora, B, res, mode, gflop, *rest in CONFIGS:
        unfreeze = rest[0] if rest else 0
        if only and not any(o in name for o in only):
            continue
        torch.manual_seed(0)
        t0 = time.time()
        m = (Dinov2PoseModelLoRA(backbone=arch) if lora else Dinov2PoseModel(backbone=arch, unfreeze_last_n_layers=unfreeze)).to(dev)
        batch = {k: v.to(dev) for k, v in make_inputs(B, res, res, 0).items()}
        if mode == "train":
            tr = PoseTrainer(m)
            step = lambda: tr.step(batch["pixel_values"], batch["heatmaps"], batch["keypoints"], batch["z"])  # noqa: E731
        else:
            m.eval()

            def step():
                with torch.no_grad():
                    hm, _ = m(batch["pixel_values"])
                    decode_heatmaps(hm, (res, res))
        for _ in range(4):
            step()
        ms = timed(step, 10 if B > 1 else 50)
        print(json.dumps({"config": name, "batch": B, "ms": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1),
                          "algorithmic_tflops": round(gflop * B / ms, 1), "setup_s": round(time.time() - t0, 1),
                          "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
        del m, batch
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
