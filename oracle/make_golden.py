"""Freeze outputs of the REAL reference into ``tests/golden/*.npz`` (TEST INFRASTRUCTURE).

Run in the build container only (needs ``/root/reference``):

    python -m oracle.make_golden            # writes tests/golden/*.npz

Each case builds the unmodified reference model (``oracle/ref_harness.py`` documents the two
offline patches), loads the synthetic ``state_dict`` from ``oracle/weights.py`` with
``strict=True`` (which also proves the generated key names/shapes equal the reference's),
runs it in fp32 on CPU and stores the outputs.  Large tensors are stored as a strided
subsample (``subsample``) plus their L2 norm; the tests apply the same subsampling.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ref_harness
from .weights import ARCHS, make_inputs, make_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
MAX_KEEP = 4096

# name, arch, lora_rank, batch, res, mode
MODEL_CASES = [
    ("tiny_frozen_b2_224_eval", "test/dinov2-tiny", 0, 2, 224, "eval"),
    ("tiny_lora_b3_224_train", "test/dinov2-tiny", 8, 3, 224, "train"),
    ("small_frozen_b2_224_eval", "facebook/dinov2-small", 0, 2, 224, "eval"),
    ("small_lora_b2_224_eval", "facebook/dinov2-small", 8, 2, 224, "eval"),
    ("small_lora_b4_224_train", "facebook/dinov2-small", 8, 4, 224, "train"),
    ("small_frozen_b1_448_eval", "facebook/dinov2-small", 0, 1, 448, "eval"),
    ("small_lora_b2_448_train", "facebook/dinov2-small", 8, 2, 448, "train"),
    ("base_lora_b2_224_train", "facebook/dinov2-base", 8, 2, 224, "train"),
    ("large_frozen_b1_224_eval", "facebook/dinov2-large", 0, 1, 224, "eval"),
    # training step with the heads in eval mode (model.train(); model.pose_heads.eval()): BatchNorm on its running
    # statistics -- no batch-statistics cancellation, so the gradient fixtures are tight (mode "trainfz")
    ("tiny_lora_b3_224_trainfz", "test/dinov2-tiny", 8, 3, 224, "trainfz"),
    ("small_lora_b4_224_trainfz", "facebook/dinov2-small", 8, 4, 224, "trainfz"),
    ("base_lora_b2_224_trainfz", "facebook/dinov2-base", 8, 2, 224, "trainfz"),
]

# Dinov2PoseModel(unfreeze_last_n_layers=n) (reference model/dinov2_pose.py:25-39):
# name, arch, lora_rank (0), batch, res, mode, n
UNFREEZE_CASES = [
    ("tiny_unfreeze2_b3_224_train", "test/dinov2-tiny", 0, 3, 224, "train", 2),
    ("small_unfreeze2_b2_224_train", "facebook/dinov2-small", 0, 2, 224, "train", 2),
    ("small_unfreeze2_b2_224_trainfz", "facebook/dinov2-small", 0, 2, 224, "trainfz", 2),
]


def subsample(t):
    """Deterministic strided subsample of a tensor/array (<= MAX_KEEP values) as fp32 numpy."""
    a = t.detach().cpu().float().numpy() if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.float32)
    flat = a.reshape(-1)
    stride = max(1, flat.size // MAX_KEEP)
    return flat[::stride].copy()


def l2(t):
    return np.float64(torch.as_tensor(t).double().norm().item())


def build_reference_model(dp, arch, lora_rank, sd, unfreeze=0):
    if lora_rank:
        m = dp.Dinov2PoseModelLoRA(num_keypoints=24, backbone=arch, heatmap_size=48,
                                   lora_rank=lora_rank, lora_alpha=16, lora_dropout=0.0)
    else:
        m = dp.Dinov2PoseModel(num_keypoints=24, backbone=arch, heatmap_size=48, unfreeze_last_n_layers=unfreeze)
    ref_sd = m.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()) or set(ref_sd.keys()) == set(sd.keys()), \
        (set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), (k, ref_sd[k].shape, sd[k].shape)
    m.load_state_dict(sd, strict=True)
    # z-head dropout off for parity (cannot RNG-match torch dropout); LoRA dropout is 0 above
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def run_model_case(dp, losses, name, arch, lora_rank, batch, res, mode, unfreeze=0):
    torch.manual_seed(0)
    sd = make_state_dict(arch, seed=0, lora_rank=lora_rank)
    m = build_reference_model(dp, arch, lora_rank, sd, unfreeze)
    inp = make_inputs(batch, res, res, seed=0)
    out = {}
    hooks = {}

    def tap(key):
        def fn(_m, _i, o):
            hooks[key] = o if isinstance(o, torch.Tensor) else o[0]
        return fn

    L = ARCHS[arch][1]
    hs = [m.backbone.embeddings.register_forward_hook(tap("tokens_embed")),
          m.backbone.encoder.layer[0].register_forward_hook(tap("hidden0")),
          m.backbone.encoder.layer[L - 1].register_forward_hook(tap("hidden_last")),
          m.backbone.layernorm.register_forward_hook(tap("tokens_final")),
          m.pose_heads.heatmap_head.feature_refine[2].register_forward_hook(tap("fr0")),
          m.pose_heads.heatmap_head.feature_refine[3].register_forward_hook(tap("hg")),
          m.pose_heads.heatmap_head.feature_refine[6].register_forward_hook(tap("fr4")),
          m.pose_heads.heatmap_head.upsampling[0].register_forward_hook(tap("up0")),
          m.pose_heads.heatmap_head.upsampling[1].register_forward_hook(tap("up1")),
          m.pose_heads.heatmap_head.prediction[2].register_forward_hook(tap("pred0"))]
    if mode == "eval":
        m.eval()
        with torch.no_grad():
            hm, z = m(inp["pixel_values"])
    else:
        kp_loss_fn, z_loss_fn, DLW = losses
        m.train()
        if mode == "trainfz":
            m.pose_heads.eval()
        hm, z = m(inp["pixel_values"])
        conf = inp["keypoints"][..., 2]
        kp = kp_loss_fn(hm, inp["heatmaps"], conf)
        zl = z_loss_fn(z, inp["z"], conf)
        w = DLW(initial_weight=0.1, target_ratio=1.0, adjustment_rate=0.1)
        w.update(kp.item(), zl.item(), is_validation=False)    # train.py:154-158
        loss = w.get_balanced_loss(kp, zl)                     # train.py:163
        loss.backward()                                        # train.py:169
        out["kp_loss"], out["z_loss"], out["loss"] = kp.item(), zl.item(), loss.item()
        n_grad = 0
        for pname, p in m.named_parameters():
            if p.grad is None:
                assert not p.requires_grad or "backbone" in pname, pname
                continue
            n_grad += 1
            out["grad." + pname] = subsample(p.grad)
            out["gradnorm." + pname] = l2(p.grad)
        out["num_grad_tensors"] = n_grad
        for bname, b in m.named_buffers():
            if "running_" in bname:
                out["buf." + bname] = subsample(b)
    for h in hs:
        h.remove()
    out["heatmaps"] = hm.detach().numpy().astype(np.float32)
    out["z"] = z.detach().numpy().astype(np.float32)
    for k, v in hooks.items():
        out["sub." + k] = subsample(v)
        out["norm." + k] = l2(v.detach())
    out["meta"] = np.array([name, arch, str(lora_rank), str(batch), str(res), mode])
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:4]})


def decode_edge_maps():
    """Heat-maps exercising the behaviours listed in SURVEY.md section 8c."""
    rng = np.random.default_rng(7)
    maps = []
    a = rng.standard_normal((48, 48)).astype(np.float32) * 0.05
    a[5, 7] = a[5, 9] = a[3, 40] = 1.0          # ties -> first in row-major order
    maps.append(a)
    b = rng.standard_normal((48, 48)).astype(np.float32)
    b[10, 11] = np.nan
    b[30, 2] = np.nan                            # NaN counts as max -> first NaN
    maps.append(b)
    c = np.zeros((48, 48), np.float32)
    c[0, 47] = 1.0
    c[0, 46] = 0.5
    c[1, 47] = 0.25                              # border peak, clipped window
    maps.append(c)
    d = -np.abs(rng.standard_normal((48, 48)).astype(np.float32)) - 1.0   # all negative
    maps.append(d)
    e = np.zeros((48, 48), np.float32)           # all equal -> index 0, zero-sum window -> nan
    maps.append(e)
    f = np.zeros((48, 48), np.float32)
    f[47, 0] = 3.0                               # bottom-left corner
    maps.append(f)
    yy, xx = np.mgrid[0:48, 0:48]
    g = np.exp(-((yy - 20.3) ** 2 + (xx - 31.7) ** 2) / (2 * 1.5 ** 2)).astype(np.float32)  # Gaussian peak
    maps.append(g)
    h = rng.standard_normal((48, 48)).astype(np.float32)
    h[25, 25] = np.inf
    maps.append(h)
    while len(maps) % 24:
        maps.append(rng.standard_normal((48, 48)).astype(np.float32) * 0.1 + 0.06)
    return np.stack(maps).reshape(-1, 24, 48, 48)


def run_decode_cases():
    smu = ref_harness.import_reference_decode()
    rng = np.random.default_rng(0)
    cases = {
        "random": (rng.standard_normal((4, 24, 48, 48)).astype(np.float32) * 0.07 + 0.06, (224, 224)),
        "edges": (decode_edge_maps(), (224, 224)),
        "nonsquare": (rng.random((2, 24, 48, 48)).astype(np.float32), (640, 480)),
        "rect_map": (rng.random((1, 24, 32, 64)).astype(np.float32), (512, 256)),
    }
    out = {}
    with np.errstate(all="ignore"):
        for name, (hm, tgt) in cases.items():
            xy = smu.get_keypoints_from_heatmaps_batch(hm, tgt)           # [B,K,2] float64
            idx = np.zeros(hm.shape[:2] + (2,), np.int64)
            for b in range(hm.shape[0]):
                for k in range(hm.shape[1]):
                    r, c, _ = smu.argmax_ind(hm[b, k])
                    idx[b, k] = (r, c)
            out[name + ".heatmaps"] = hm
            out[name + ".target"] = np.array(tgt, np.int64)
            out[name + ".xy"] = np.asarray(xy, np.float64)
            out[name + ".idx"] = idx
    np.savez_compressed(os.path.join(GOLDEN_DIR, "decode.npz"), **out)
    print("wrote decode", {k: v.shape for k, v in out.items() if k.endswith(".xy")})


def run_loss_cases(losses):
    kp_loss_fn, z_loss_fn, DLW = losses
    inp = make_inputs(5, seed=3)
    g = torch.Generator().manual_seed(11)
    pred_hm = torch.randn(5, 24, 48, 48, generator=g) * 0.3
    pred_z = torch.randn(5, 24, generator=g)
    conf = inp["keypoints"][..., 2]
    out = {"pred_hm": pred_hm.numpy(), "pred_z": pred_z.numpy()}
    out["kp_loss"] = kp_loss_fn(pred_hm, inp["heatmaps"], conf).item()
    out["z_loss"] = z_loss_fn(pred_z, inp["z"], conf).item()
    w = DLW()
    seq = []
    for kp, zl in [(0.02, 0.8), (0.018, 0.7), (0.03, 0.2), (1e-9, 5.0), (4.0, 1e-9)]:
        wt = w.update(kp, zl)
        bal = w.get_balanced_loss(torch.tensor(kp), torch.tensor(zl)).item()
        seq.append((wt, w.kp_loss_avg, w.z_loss_avg, bal))
    out["dlw_seq"] = np.array(seq, np.float64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "losses.npz"), **out)
    print("wrote losses", out["kp_loss"], out["z_loss"])


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    dp, _lora, _ph = ref_harness.import_reference()
    losses = ref_harness.import_reference_losses()
    only = set(argv[1:])
    for case in MODEL_CASES + UNFREEZE_CASES:
        if only and case[0] not in only:
            continue
        run_model_case(dp, losses, *case)
    if not only or "decode" in only:
        run_decode_cases()
    if not only or "losses" in only:
        run_loss_cases(losses)


if __name__ == "__main__":
    main(sys.argv)
