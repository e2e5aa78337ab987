"""Pin the oracle (oracle/*.py) to the outputs of the REAL reference frozen in tests/golden
by oracle/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, pose_oracle
from oracle.make_golden import MODEL_CASES, UNFREEZE_CASES, subsample
from oracle.weights import make_inputs, make_state_dict

FAST = [c for c in MODEL_CASES if c[0].startswith(("tiny", "small"))]
SLOW = [c for c in MODEL_CASES if not c[0].startswith(("tiny", "small"))]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _relmax(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("case", FAST + SLOW + UNFREEZE_CASES, ids=lambda c: c[0])
def test_model_oracle_matches_reference(golden_dir, case):
    name, arch, lora_rank, batch, res, mode = case[:6]
    unfreeze = case[6] if len(case) > 6 else 0      # Dinov2PoseModel(unfreeze_last_n_layers=n), reference :25-39
    g = _load(golden_dir, name)
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state_dict(arch, seed=0, lora_rank=lora_rank)
    inp = make_inputs(batch, res, res, seed=0)
    lora = {"rank": lora_rank, "alpha": 16, "dropout": 0.0} if lora_rank else None
    tol = 2e-5
    if mode == "eval":
        aux = {}
        with torch.no_grad():
            hm, z = pose_oracle.model_forward(sd, inp["pixel_values"], arch, lora, False, aux=aux)
        for key, val in (("tokens_embed", aux["tokens_embed"]), ("hidden0", aux["hidden"][0]),
                         ("hidden_last", aux["hidden"][-1]), ("tokens_final", aux["tokens_final"]),
                         ("fr0", aux["fr0"]), ("hg", aux["hg"]), ("fr4", aux["fr4"]),
                         ("up0", aux["up0"]), ("up1", aux["up1"]), ("pred0", aux["pred0"])):
            assert _relmax(subsample(val), g["sub." + key]) < tol, key
    else:
        frozen = mode == "trainfz"      # model.train(); model.pose_heads.eval()
        out = pose_oracle.loss_and_grads(sd, inp, arch, lora, training=True, unfreeze=unfreeze,
                                         heads_training=False if frozen else None)
        hm, z = out["heatmaps"], out["z"]
        assert abs(out["kp_loss"].item() - g["kp_loss"]) < 1e-6
        assert abs(out["z_loss"].item() - g["z_loss"]) < 1e-6
        assert abs(out["loss"].item() - g["loss"]) < 1e-5
        n = 0
        for pname, grad in out["grads"].items():
            assert grad is not None, pname
            ref = g["grad." + pname]
            gn = float(g["gradnorm." + pname])
            if gn < 1e-6:
                # conv bias feeding a train-mode BatchNorm: analytically zero, pure rounding noise
                assert float(grad.norm()) < 1e-6, pname
            else:
                # gradients upstream of train-mode BN are cancelling sums (fp32 vs fp64 already
                # differs by ~5e-4 relative L2, measured) -> relative-L2 tolerance, not element-wise
                sub = subsample(grad)
                rel = np.linalg.norm(sub - ref) / (np.linalg.norm(ref) + 1e-30)
                # un-frozen backbone layers sit one more cancelling stage upstream: at tiny / batch 3 the fp32 oracle, the
                # fp32 reference and an fp64 run of the oracle differ pairwise by 3e-3 .. 7.5e-3 (measured); at ViT-S the
                # same comparison gives <= 5e-4
                # heads in eval mode: no batch-statistics cancellation; measured worst case 6.5e-4 (tiny, feature_refine.0.weight)
                assert rel < (2e-3 if frozen else 1.5e-2 if unfreeze else 5e-3), (pname, rel)
            n += 1
        assert n == int(g["num_grad_tensors"])
        for k in g.files:
            if k.startswith("buf."):
                assert _relmax(subsample(sd[k[4:]]), g[k]) < tol, k
    assert _relmax(hm.numpy(), g["heatmaps"]) < tol
    assert _relmax(z.numpy(), g["z"]) < tol


def test_decode_oracle_matches_reference(golden_dir):
    g = _load(golden_dir, "decode")
    for case in ("random", "edges", "nonsquare", "rect_map"):
        hm = g[case + ".heatmaps"]
        tgt = tuple(int(v) for v in g[case + ".target"])
        idx, xy = decode_oracle.decode_batch(hm, tgt)
        assert np.array_equal(idx, g[case + ".idx"]), case
        ref = g[case + ".xy"]
        # bit-exact including inf / nan placement
        assert np.array_equal(xy.view(np.uint64) == ref.view(np.uint64), np.ones(ref.shape, bool)) or \
            np.array_equal(np.nan_to_num(xy, nan=1234.5), np.nan_to_num(ref, nan=1234.5)), case


def test_first_argmax_forms_agree():
    rng = np.random.default_rng(5)
    for _ in range(50):
        a = rng.standard_normal(97).astype(np.float32)
        if rng.random() < 0.5:
            a[rng.integers(0, 97, 3)] = a.max()
        if rng.random() < 0.3:
            a[rng.integers(0, 97, 2)] = np.nan
        assert decode_oracle.first_argmax(a) == decode_oracle.first_argmax_loop(a)


def test_losses_oracle_matches_reference(golden_dir):
    g = _load(golden_dir, "losses")
    inp = make_inputs(5, seed=3)
    conf = inp["keypoints"][..., 2]
    kp = pose_oracle.keypoint_loss(torch.from_numpy(g["pred_hm"]), inp["heatmaps"], conf).item()
    zl = pose_oracle.z_loss(torch.from_numpy(g["pred_z"]), inp["z"], conf).item()
    assert abs(kp - float(g["kp_loss"])) < 1e-7
    assert abs(zl - float(g["z_loss"])) < 1e-7
    w = pose_oracle.DynamicLossWeighting()
    for (kpv, zv), row in zip([(0.02, 0.8), (0.018, 0.7), (0.03, 0.2), (1e-9, 5.0), (4.0, 1e-9)], g["dlw_seq"]):
        wt = w.update(kpv, zv)
        bal = w.balanced(torch.tensor(kpv), torch.tensor(zv)).item()
        assert np.allclose([wt, w.kp_avg, w.z_avg, bal], row, rtol=1e-6)
