"""Drop-in proof: the reference's OWN caller code, run against ``dino_pose_b200.model``.

CPU part (runs where the reference tree is mounted, i.e. in the build container): the source of ``DynamicLossWeighting``,
``keypoint_loss``, ``z_loss`` and ``train_one_epoch`` is taken verbatim from ``/root/reference/train.py`` (:17-202) with
``ast`` -- nothing of it is copied into this repo -- and executed with our ``Dinov2PoseModelLoRA`` (torch emulator of the
C-ABI ops, fp32 storage) and a stock ``torch.optim.AdamW``; the epoch losses must match the same function driving the
ORACLE model.  GPU part (no reference tree on the box): the same call sequence restated line by line
(train.py:134-188, benchmark_model.py:31-51), through the CUDA path, against the oracle on the host."""
import ast
import os
import time

import numpy as np
import pytest
import torch

from oracle import pose_oracle
from oracle.weights import make_inputs, make_state_dict

from dino_pose_b200.model import Dinov2PoseModel, Dinov2PoseModelLoRA

REF_TRAIN = "/root/reference/train.py"
ARCH_CPU = "test/dinov2-tiny"


def _reference_namespace():
    """exec the four definitions of the reference's train.py (no other part of the file: its imports need the data loaders)"""
    tree = ast.parse(open(REF_TRAIN).read())
    wanted = {"DynamicLossWeighting", "keypoint_loss", "z_loss", "train_one_epoch"}
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in wanted]
    assert {n.name for n in body} == wanted
    from tqdm import tqdm
    ns = {"torch": torch, "time": time, "tqdm": tqdm, "np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF_TRAIN, "exec"), ns)
    return ns


def _batches(n, B, seed0=0):
    out = []
    for s in range(n):
        b = make_inputs(B, 224, 224, seed0 + s)
        out.append({"image": b["pixel_values"], "2d_heatmaps": b["heatmaps"], "2d_keypoints": b["keypoints"], "z_coords": b["z"]})
    return out


class _OracleModel(torch.nn.Module):
    """the oracle's functional forward behind the nn.Module call form train_one_epoch uses"""

    def __init__(self, arch, sd, lora):
        super().__init__()
        self.arch, self.lora, self.sd = arch, lora, sd
        self.names = pose_oracle.trainable_names(sd, lora)
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(sd[n]) for n in self.names])
        for n, p in zip(self.names, self.ps):
            sd[n] = p

    def forward(self, px):
        return pose_oracle.model_forward(self.sd, px, self.arch, self.lora, training=self.training)


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="reference tree not mounted")
def test_reference_train_one_epoch_runs_on_our_model():
    from tests.emulator import TorchEmulator
    ns = _reference_namespace()
    lr, wd, eps = 1e-3, 1e-6, 1e-3          # eps: see tests/test_trainer_cpu.py (zero-gradient parameters under Adam)
    batches = _batches(3, 2)
    lora = {"rank": 8, "alpha": 16, "dropout": 0.0}
    # ours
    m = Dinov2PoseModelLoRA(backbone=ARCH_CPU, lora_rank=8, lora_alpha=16, lora_dropout=0.0)
    m.load_state_dict(make_state_dict(ARCH_CPU, 0, 8))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m._backend_factory, m._act_dtype = TorchEmulator, torch.float32
    opt = torch.optim.AdamW(filter(lambda p: p.requires_grad, m.parameters()), lr=lr, weight_decay=wd, eps=eps)
    w = ns["DynamicLossWeighting"](initial_weight=0.1)
    got = [ns["train_one_epoch"](m, batches, torch.device("cpu"), opt, w, e) for e in range(2)]
    val = ns["train_one_epoch"](m, batches[:1], torch.device("cpu"), opt, w, 0, is_validation=True)
    # the same reference function on the oracle model
    om = _OracleModel(ARCH_CPU, make_state_dict(ARCH_CPU, 0, 8), lora)
    oopt = torch.optim.AdamW(om.parameters(), lr=lr, weight_decay=wd, eps=eps)
    ow = ns["DynamicLossWeighting"](initial_weight=0.1)
    ref = [ns["train_one_epoch"](om, batches, torch.device("cpu"), oopt, ow, e) for e in range(2)]
    oval = ns["train_one_epoch"](om, batches[:1], torch.device("cpu"), oopt, ow, 0, is_validation=True)
    for a, b in zip(got + [val], ref + [oval]):
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-3 * abs(y) + 1e-7, (got, ref)
    assert got[1][0] != got[0][0]           # the optimizer really moved the parameters between the epochs
    assert abs(w.weight - ow.weight) < 1e-4


# ------------------------------------------------------------------------------------------------ GPU twins
def _train_one_epoch_restated(model, dataloader, device, optimizer, weighting, is_validation=False):
    """reference train.py:122-202, call for call (progress bar and printing dropped)"""
    model.train() if not is_validation else model.eval()
    run = [0.0, 0.0, 0.0]
    with torch.no_grad() if is_validation else torch.enable_grad():
        for batch in dataloader:
            pixel_values = batch["image"].to(device)
            heatmaps = batch["2d_heatmaps"].to(device)
            kps = batch["2d_keypoints"].to(device)
            z_coords = batch["z_coords"].to(device)
            if not is_validation:
                optimizer.zero_grad()
            pred_heatmaps, pred_z = model(pixel_values)
            conf = kps[..., 2]
            kp = pose_oracle.keypoint_loss(pred_heatmaps, heatmaps, conf)
            zl = pose_oracle.z_loss(pred_z, z_coords, conf)
            weight = weighting.weight if is_validation else weighting.update(kp.item(), zl.item())   # train.py:30-32
            loss = weighting.balanced(kp, zl) if not is_validation else kp + weight * zl
            if not is_validation:
                loss.backward()
                optimizer.step()
            for i, v in enumerate((loss, kp, zl)):
                run[i] += v.item()
    return tuple(v / len(dataloader) for v in run)


@pytest.mark.gpu
def test_train_loop_call_sequence_on_gpu():
    arch, dev = "facebook/dinov2-small", torch.device("cuda:0")
    lora = {"rank": 8, "alpha": 16, "dropout": 0.0}
    batches = _batches(2, 4)
    m = Dinov2PoseModelLoRA(backbone=arch, lora_rank=8, lora_alpha=16, lora_dropout=0.0)
    m.load_state_dict(make_state_dict(arch, 0, 8))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.to(dev)
    opt = torch.optim.AdamW(filter(lambda p: p.requires_grad, m.parameters()), lr=3e-5, weight_decay=1e-6)
    w = pose_oracle.DynamicLossWeighting()
    got = _train_one_epoch_restated(m, batches, dev, opt, w)
    val = _train_one_epoch_restated(m, batches[:1], dev, opt, w, is_validation=True)
    om = _OracleModel(arch, make_state_dict(arch, 0, 8), lora)
    oopt = torch.optim.AdamW(om.parameters(), lr=3e-5, weight_decay=1e-6)
    ow = pose_oracle.DynamicLossWeighting()
    ref = _train_one_epoch_restated(om, batches, torch.device("cpu"), oopt, ow)
    oval = _train_one_epoch_restated(om, batches[:1], torch.device("cpu"), oopt, ow, is_validation=True)
    for a, b in zip(got + val, ref + oval):
        assert abs(a - b) <= 2e-2 * abs(b) + 1e-6, (got, val, ref, oval)     # bf16 forward, stated tolerance 2e-2
    assert all(p.grad is not None for p in m.parameters() if p.requires_grad)


@pytest.mark.gpu
def test_benchmark_model_call_sequence_on_gpu():
    """reference benchmark_model.py:21-51: frozen ViT-S, eval, the model's own image processor on a PIL-style frame,
    three warm-up calls, timed calls, count_parameters()."""
    dev = torch.device("cuda")
    model = Dinov2PoseModel(num_keypoints=24, backbone="facebook/dinov2-small")
    model.to(dev)
    model.eval()
    processor = model.image_processor
    rng = np.random.RandomState(0)
    frame = rng.randint(0, 255, (224, 224, 3), dtype=np.uint8)
    for _ in range(3):
        inputs = processor(frame, return_tensors="pt")
        pixel_values = inputs.pixel_values.to(dev)
        with torch.no_grad():
            _ = model(pixel_values)
    times = []
    for _ in range(5):
        start = time.time()
        inputs = processor(frame, return_tensors="pt")
        pixel_values = inputs.pixel_values.to(dev)
        with torch.no_grad():
            heatmaps, depths = model(pixel_values)
        torch.cuda.synchronize()
        times.append(time.time() - start)
    assert heatmaps.shape == (1, 24, 48, 48) and depths.shape == (1, 24)
    assert torch.isfinite(heatmaps).all() and torch.isfinite(depths).all()
    assert model.count_parameters() > 0 and 1.0 / np.mean(times) > 30.0      # the reference's own real-time bar (:63-65)
