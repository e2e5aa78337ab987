#!/usr/bin/env python
"""Short profiling target: a few fine-tuning steps of BASELINE.json configs[1] (ViT-S/14 + LoRA, batch 64, 224x224)
with nothing else in the process -- the command ncu wraps (profiles/README.md).  Not a benchmark."""
import argparse
import os

# random-init weights of the named architecture (BASELINE.json north_star: no network, no checkpoints): explicit opt-in
os.environ.setdefault("DINO_POSE_RANDOM_INIT", "1")
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dino_pose_b200.model import Dinov2PoseModelLoRA   # noqa: E402
from dino_pose_b200.train import PoseTrainer           # noqa: E402
from dino_pose_b200.synthetic import make_inputs       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--arch", default="facebook/dinov2-small")
    ap.add_argument("--eval", action="store_true", help="inference forward + decode instead of the training step")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = Dinov2PoseModelLoRA(backbone=args.arch, lora_dropout=0.1).to(dev)
    batch = {k: v.to(dev) for k, v in make_inputs(args.batch, args.res, args.res, 0).items()}
    if args.eval:
        model.eval()
        with torch.no_grad():
            for _ in range(args.steps):
                model(batch["pixel_values"])
    else:
        tr = PoseTrainer(model)
        for _ in range(args.steps):
            tr.step(batch["pixel_values"], batch["heatmaps"], batch["keypoints"], batch["z"])
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
