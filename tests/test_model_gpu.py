"""End-to-end parity on a B200 through the drop-in module surface (Dinov2PoseModel / Dinov2PoseModelLoRA):
CUDA path vs the golden vectors frozen from the real reference, vs the oracle on fresh inputs, and -- for
gradients -- vs the same op graph emulated in torch (isolates kernel errors from bf16 rounding effects)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import decode_oracle, pose_oracle
from oracle.make_golden import MODEL_CASES, UNFREEZE_CASES, subsample
from oracle.weights import make_inputs, make_state_dict

TOL = 2e-2   # north_star: max|a-b| / max|b| <= 2e-2 for heat-maps and z vs the fp32 reference (eval mode)
# Train mode: the 14 batch-statistics BatchNorms re-normalise every head layer, and the fp32 ORACLE heads applied
# to backbone features perturbed at the bf16 level (4e-3 rel-L2) already move the heat-maps by 1.2-1.8e-2
# (measured, DESIGN.md "Tolerances"): the head's train-mode condition number is ~3.  Stated train-mode bound:
TOL_TRAIN_HM = 4e-2


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def build(arch, lora_rank, device="cuda", backend_factory=None, unfreeze=0):
    from dino_pose_b200.model import Dinov2PoseModel, Dinov2PoseModelLoRA
    if lora_rank:
        m = Dinov2PoseModelLoRA(backbone=arch, lora_rank=lora_rank, lora_alpha=16, lora_dropout=0.0)
    else:
        m = Dinov2PoseModel(backbone=arch, unfreeze_last_n_layers=unfreeze)
    m.load_state_dict(make_state_dict(arch, 0, lora_rank))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m._backend_factory = backend_factory
    return m.to(device)


EVAL_CASES = [c for c in MODEL_CASES if c[5] == "eval"]
TRAIN_CASES = [c for c in MODEL_CASES + UNFREEZE_CASES if c[5] == "train"]
# training step with the heads in eval mode (model.train(); model.pose_heads.eval()): BatchNorm on running statistics
FROZEN_CASES = [c for c in MODEL_CASES + UNFREEZE_CASES if c[5] == "trainfz"]


@pytest.mark.parametrize("case", EVAL_CASES, ids=lambda c: c[0])
def test_eval_forward_vs_reference_golden(golden_dir, case):
    name, arch, lora_rank, batch, res, _ = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m = build(arch, lora_rank).eval()
    inp = make_inputs(batch, res, res, 0)
    with torch.no_grad():
        hm, z = m(inp["pixel_values"].cuda())
    torch.cuda.synchronize()
    assert hm.dtype == torch.float32 and tuple(hm.shape) == g["heatmaps"].shape
    print(name, "eval hm max-rel", relmax(hm, g["heatmaps"]), "z", relmax(z, g["z"]))
    assert relmax(hm, g["heatmaps"]) < TOL, name
    assert relmax(z, g["z"]) < TOL, name


def _loss(hm, z, inp):
    conf = inp["keypoints"][..., 2]
    kp = pose_oracle.keypoint_loss(hm, inp["heatmaps"], conf)
    zl = pose_oracle.z_loss(z, inp["z"], conf)
    w = pose_oracle.DynamicLossWeighting()
    w.update(kp.item(), zl.item())
    return w.balanced(kp, zl), kp, zl


@pytest.mark.parametrize("case", TRAIN_CASES + FROZEN_CASES, ids=lambda c: c[0])
def test_train_step_vs_reference_golden(golden_dir, case):
    name, arch, lora_rank, batch, res, mode = case[:6]
    frozen = mode == "trainfz"
    unfreeze = case[6] if len(case) > 6 else 0     # Dinov2PoseModel(unfreeze_last_n_layers=n): full backward of n layers
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m = build(arch, lora_rank, unfreeze=unfreeze).train()
    if frozen:
        m.pose_heads.eval()
    inp = {k: v.cuda() for k, v in make_inputs(batch, res, res, 0).items()}
    hm, z = m(inp["pixel_values"])
    print(name, "train hm max-rel", relmax(hm.detach(), g["heatmaps"]), "z", relmax(z.detach(), g["z"]))
    assert relmax(hm.detach(), g["heatmaps"]) < (TOL if frozen else TOL_TRAIN_HM)
    assert relmax(z.detach(), g["z"]) < TOL
    loss, kp, zl = _loss(hm, z, inp)
    assert abs(kp.item() - float(g["kp_loss"])) / float(g["kp_loss"]) < 2e-2
    assert abs(zl.item() - float(g["z_loss"])) / float(g["z_loss"]) < 2e-2
    loss.backward()
    torch.cuda.synchronize()
    n = 0
    bad = {}
    allg, allr = [], []
    for pname, p in m.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        n += 1
        gn = float(g["gradnorm." + pname])
        if gn < 1e-6:
            # analytically zero.  The key bias (softmax is shift invariant) is the column sum of the bf16 dK tile: it
            # cancels to rounding noise, 2e-6 against 9e-4 for the query bias of the same layer
            assert float(p.grad.norm()) < (2e-5 if pname.endswith("attention.key.bias") else 1e-6), pname
            continue
        ref = g["grad." + pname]
        sub = subsample(p.grad)
        rel = np.linalg.norm(sub - ref) / (np.linalg.norm(ref) + 1e-30)
        cos = float(np.dot(sub, ref) / (np.linalg.norm(sub) * np.linalg.norm(ref) + 1e-30))
        ratio = float(np.linalg.norm(sub) / (np.linalg.norm(ref) + 1e-30))
        print(f"  {pname[-60:]:60s} relL2 {rel:.3e} cos {cos:.4f} |g|/|ref| {ratio:.4f}")
        allg.append(sub); allr.append(ref)
        # per-tensor bf16-vs-fp32 criterion (tests/test_engine_emulated.py explains why train-mode BN at batch 2-4
        # makes single small tensors move by tens of percent); the tight kernel-level check of the backward is
        # test_train_step_cuda_vs_emulated_op_graph below
        # Measured worst case over all golden cases: relL2 0.33, cos 0.946 (batch 2-4).  The norm ratio is the scale check:
        # bf16 noise is nearly orthogonal to the gradient, a wrong factor is not (a 1.3x error fails; the tight version of
        # this check, at batch 64 where the BatchNorm statistics are stable, is tests/test_parity_bench_shape_gpu.py).
        # Heads in eval mode (frozen statistics, "trainfz" goldens): no cancellation, the same program is held to
        # relL2 0.15 / cos 0.99 / ratio within 5 % per tensor (torch emulation of the op graph measures 0.12 / 0.993 at
        # tiny batch 3 and 0.095 / 0.9955 at ViT-S batch 4 from bf16 storage alone, tests/test_engine_emulated.py).
        lim = (0.15, 0.99, 0.95, 1.05) if frozen else (0.4, 0.93, 0.8, 1.2)
        if rel > lim[0] or cos < lim[1] or not (lim[2] < ratio < lim[3]):
            bad[pname] = (float(rel), cos, ratio)
    assert n == int(g["num_grad_tensors"])
    assert not bad, bad
    fa, fr = np.concatenate(allg), np.concatenate(allr)
    gcos = float(np.dot(fa, fr) / (np.linalg.norm(fa) * np.linalg.norm(fr)))
    gratio = float(np.linalg.norm(fa) / np.linalg.norm(fr))
    print(f"  all trainable gradients: cosine {gcos:.4f} norm ratio {gratio:.4f}")
    assert gcos > (0.995 if frozen else 0.97), gcos
    assert (0.98 if frozen else 0.95) < gratio < (1.02 if frozen else 1.05), gratio
    bufs = dict(m.named_buffers())
    for k in g.files:
        if k.startswith("buf."):
            assert relmax(subsample(bufs[k[4:]]), g[k]) < (1e-7 if frozen else TOL), k
        if k.startswith("buf.") and k.endswith("running_mean"):
            assert int(bufs[k[4:].replace("running_mean", "num_batches_tracked")].item()) == (0 if frozen else 1)


def test_heads_in_eval_mode_gradient_is_additive_over_the_batch():
    """CUDA twin of tests/test_engine_emulated.py's additivity check: with frozen BatchNorm statistics
    the gradient of the concatenated batch equals the sum of the shard gradients -- what an N-rank all-reduced step
    computes against a 1-rank step.  Tile shapes / split-K depend on the batch, so agreement is at the bf16 rounding
    level of the saved activations, not bit-exact."""
    arch = "facebook/dinov2-small"
    inp = {k: v.cuda() for k, v in make_inputs(8, 224, 224, 11).items()}
    gen = torch.Generator().manual_seed(3)
    w_hm, w_z = torch.randn(8, 24, 48, 48, generator=gen).cuda(), torch.randn(8, 24, generator=gen).cuda()

    def grads(sl):
        m = build(arch, 8).train()
        m.pose_heads.eval()
        hm, z = m(inp["pixel_values"][sl])
        ((hm * w_hm[sl]).sum() + (z * w_z[sl]).sum()).backward()
        torch.cuda.synchronize()
        return torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.requires_grad])
    full, a, b = (grads(sl) for sl in (slice(0, 8), slice(0, 4), slice(4, 8)))
    err = ((a + b - full).norm() / full.norm()).item()
    print("additivity over the batch, heads in eval mode: rel-L2", err)
    assert err < 2e-2, err


def test_train_step_cuda_vs_emulated_op_graph():
    """Same op graph, same bf16 rounding points, torch ops instead of our kernels (on the GPU): same
    bound as against the fp32 reference, but independent of it (no golden file involved); the tight per-kernel
    checks of the backward are in tests/test_kernels_gpu.py."""
    from tests.emulator import TorchEmulator
    arch = "facebook/dinov2-small"
    inp = {k: v.cuda() for k, v in make_inputs(4, 224, 224, 5).items()}
    grads = []
    outs = []
    for factory in (None, TorchEmulator):
        m = build(arch, 8, backend_factory=factory).train()
        hm, z = m(inp["pixel_values"])
        loss, _, _ = _loss(hm, z, inp)
        loss.backward()
        torch.cuda.synchronize()
        grads.append({n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad})
        outs.append((hm.detach(), z.detach()))
    print("cuda vs emulated: hm", relmax(outs[0][0], outs[1][0]), "z", relmax(outs[0][1], outs[1][1]))
    assert relmax(outs[0][0], outs[1][0]) < 2e-2
    assert relmax(outs[0][1], outs[1][1]) < 2e-2
    bad = {}
    for n in grads[0]:
        a, b = grads[0][n].double(), grads[1][n].double()
        if b.norm() < 1e-9:
            continue
        rel = ((a - b).norm() / b.norm()).item()
        print(f"  {n[-60:]:60s} relL2 {rel:.3e}")
        if rel > 0.35:   # train-mode BN at batch 4 amplifies 1-ulp bf16 differences (hm itself differs by ~1.6e-2)
            bad[n] = rel
    assert not bad, bad


def test_unfrozen_layers_cuda_vs_emulated_op_graph():
    """Backward through two un-frozen encoder layers: our kernels against the same op graph executed with
    torch ops at the same bf16 rounding points.  Tighter than the golden comparison for the backbone tensors."""
    from tests.emulator import TorchEmulator
    arch = "facebook/dinov2-small"
    inp = {k: v.cuda() for k, v in make_inputs(3, 224, 224, 6).items()}
    grads = []
    for factory in (None, TorchEmulator):
        m = build(arch, 0, backend_factory=factory, unfreeze=2).train()
        hm, z = m(inp["pixel_values"])
        loss, _, _ = _loss(hm, z, inp)
        loss.backward()
        torch.cuda.synchronize()
        grads.append({n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad})
    bad = {}
    nb = 0
    for n in grads[0]:
        if not n.startswith("backbone."):
            continue
        a, b = grads[0][n].double(), grads[1][n].double()
        if n.endswith("attention.key.bias"):
            assert a.norm() < 2e-5 and b.norm() < 2e-5, n      # analytically zero: bf16 rounding noise on both sides
            continue
        nb += 1
        rel = ((a - b).norm() / b.norm()).item()
        print(f"  {n[-60:]:60s} relL2 {rel:.3e}")
        if rel > 0.35:
            bad[n] = rel
    assert nb == 2 * 17
    assert not bad, bad


def test_unfrozen_layers_trainer_step_updates_backbone():
    """PoseTrainer over Dinov2PoseModel(unfreeze_last_n_layers=1): the flat layout covers the layer's 18 tensors, the
    CUDA-graph step changes them, and the next forward (eval) uses the updated weights."""
    from dino_pose_b200.train import PoseTrainer
    arch = "test/dinov2-tiny"
    m = build(arch, 0, unfreeze=1).train()
    tr = PoseTrainer(m, lr=1e-3)
    assert sum(1 for n in tr.layout["names"] if n.startswith("backbone.encoder.layer.1.")) == 18
    before = {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
    frozen_before = m.backbone.encoder.layer[0].mlp.fc1.weight.detach().clone()
    for s in range(2):
        b = {k: v.cuda() for k, v in make_inputs(4, 224, 224, s).items()}
        out = tr.step(b["pixel_values"], b["heatmaps"], b["keypoints"], b["z"])
    torch.cuda.synchronize()
    assert all(torch.isfinite(o).item() for o in out)
    changed = [n for n, p in m.named_parameters() if p.requires_grad and not torch.equal(p.detach(), before[n])]
    assert len([n for n in changed if n.startswith("backbone.")]) >= 17
    assert torch.equal(m.backbone.encoder.layer[0].mlp.fc1.weight.detach(), frozen_before)
    m.eval()
    with torch.no_grad():
        hm, z = m(b["pixel_values"])
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        rhm, rz = pose_oracle.model_forward(sd, b["pixel_values"].cpu(), arch, None, False)
    assert relmax(hm, rhm) < TOL and relmax(z, rz) < TOL


def test_decode_of_model_output_is_bit_exact():
    """north_star: decoded key-point indices bit-exact vs the reference decode applied to the SAME heat-maps."""
    from dino_pose_b200.src.model_utils import get_keypoints_from_heatmaps_batch, decode_heatmaps
    m = build("facebook/dinov2-small", 0).eval()
    inp = make_inputs(8, 224, 224, 3)
    with torch.no_grad():
        hm, _ = m(inp["pixel_values"].cuda())
    idx, xy, conf = decode_heatmaps(hm, (224, 224))
    ridx, rxy = decode_oracle.decode_batch(hm.cpu().numpy(), (224, 224))
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ridx)
    assert np.array_equal(xy.cpu().numpy().view(np.uint64), rxy.view(np.uint64))
    kps = get_keypoints_from_heatmaps_batch(hm, (224, 224))
    assert isinstance(kps, np.ndarray) and kps.dtype == np.float64 and kps.shape == (8, 24, 2)
    assert np.array_equal(kps.view(np.uint64), rxy.view(np.uint64))


def test_state_dict_roundtrip_and_module_surface():
    m = build("facebook/dinov2-small", 8)
    sd = m.state_dict()
    ref = make_state_dict("facebook/dinov2-small", 0, 8)
    assert list(sd.keys()) == list(ref.keys())
    assert m.count_parameters() == 7_839_344          # heads 7 833 200 + LoRA 6 144
    assert "LoRA" in type(m).__name__ and m.lora_config == {"rank": 8, "alpha": 16, "dropout": 0.0}
    assert m.backbone.config.hidden_size == 384 and m.feat_dim == 384
    assert m.heatmap_size == 48 and m.num_keypoints == 24
    m.apply_loading_fixes()
    assert not m.training


def test_cpu_tensor_raises_without_fallback():
    from dino_pose_b200.model import Dinov2PoseModel
    m = Dinov2PoseModel(backbone="test/dinov2-tiny").eval()
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        m(torch.zeros(1, 3, 224, 224))


def test_trainer_cuda_graph_equals_eager_and_reference_losses():
    """PoseTrainer on the GPU: (a) the CUDA-graph step and the eager step produce the same parameters;
    (b) the losses of three steps follow the reference loop (oracle forward, train.py losses, torch AdamW) within
    the bf16 tolerance; (c) warm-up / capture runs do not count as training steps."""
    from dino_pose_b200.train import PoseTrainer
    arch = "facebook/dinov2-small"
    steps, lr, wd, eps = 3, 1e-3, 1e-2, 1e-3
    batches = [{k: v.cuda() for k, v in make_inputs(4, 224, 224, s).items()} for s in range(steps)]
    results = []
    for use_graph in (False, True):
        m = build(arch, 8).train()
        tr = PoseTrainer(m, lr=lr, weight_decay=wd, eps=eps, use_graph=use_graph)
        losses = []
        for b in batches:
            out = tr.step(b["pixel_values"], b["heatmaps"], b["keypoints"], b["z"])
            losses.append([o.item() for o in out])
        torch.cuda.synchronize()
        assert int(tr.step_dev.item()) == steps
        nbt = [b for n, b in m.named_buffers() if n.endswith("num_batches_tracked")][0]
        assert int(nbt.item()) == steps
        results.append((losses, tr.flat_params.clone()))
    (l_e, p_e), (l_g, p_g) = results
    assert relmax(p_g, p_e) < 5e-3           # fp32 atomics order differs run to run (measured 1e-3 after 3 Adam steps); same kernels otherwise
    for a, b in zip(l_e, l_g):
        for x, y in zip(a, b):
            assert abs(x - y) < 2e-3 * abs(y)
    # reference loop on the CPU oracle
    sd = make_state_dict(arch, 0, 8)
    lora = {"rank": 8, "alpha": 16, "dropout": 0.0}
    names = pose_oracle.trainable_names(sd, lora)
    opt = torch.optim.AdamW([sd[n].requires_grad_(True) for n in names], lr=lr, weight_decay=wd, eps=eps)
    w = pose_oracle.DynamicLossWeighting()
    for s, b in enumerate(batches):
        b = {k: v.cpu() for k, v in b.items()}
        opt.zero_grad(set_to_none=True)
        hm, z = pose_oracle.model_forward(sd, b["pixel_values"], arch, lora, training=True)
        conf = b["keypoints"][..., 2]
        kp, zl = pose_oracle.keypoint_loss(hm, b["heatmaps"], conf), pose_oracle.z_loss(z, b["z"], conf)
        w.update(kp.item(), zl.item())
        loss = w.balanced(kp, zl)
        loss.backward()
        opt.step()
        print("step", s, "ours", l_g[s], "ref", (loss.item(), kp.item(), zl.item()))
        assert abs(l_g[s][1] - kp.item()) < 3e-2 * abs(kp.item())
        assert abs(l_g[s][2] - zl.item()) < 3e-2 * abs(zl.item())
