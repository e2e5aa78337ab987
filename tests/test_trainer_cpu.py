"""Host logic of the fine-tuning step (CPU, torch emulator of the C-ABI ops, fp32 storage): the fused loss /
AdamW / flat-parameter trainer against the reference loop (oracle forward + train.py losses + torch AdamW), and the
data-parallel path on world_size 2 over gloo (bucketed all-reduce of the flat gradient buffer)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pose_oracle
from oracle.weights import make_inputs, make_state_dict
from tests.emulator import TorchEmulator

from dino_pose_b200.model import Dinov2PoseModelLoRA
from dino_pose_b200.train import PoseTrainer

ARCH = "test/dinov2-tiny"


def build_model(seed=0):
    m = Dinov2PoseModelLoRA(backbone=ARCH, lora_rank=8, lora_alpha=16, lora_dropout=0.0)
    m.load_state_dict(make_state_dict(ARCH, seed, 8))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m._backend_factory = TorchEmulator
    m._act_dtype = torch.float32
    return m.train()


def reference_loop(batches, lr, wd, steps, eps):
    sd = make_state_dict(ARCH, 0, 8)
    lora = {"rank": 8, "alpha": 16, "dropout": 0.0}
    names = pose_oracle.trainable_names(sd, lora)
    params = [sd[n].requires_grad_(True) for n in names]
    opt = torch.optim.AdamW(params, lr=lr, weight_decay=wd, eps=eps)
    w = pose_oracle.DynamicLossWeighting()
    losses = []
    for s in range(steps):
        b = batches[s]
        opt.zero_grad(set_to_none=True)
        hm, z = pose_oracle.model_forward(sd, b["pixel_values"], ARCH, lora, training=True)
        conf = b["keypoints"][..., 2]
        kp, zl = pose_oracle.keypoint_loss(hm, b["heatmaps"], conf), pose_oracle.z_loss(z, b["z"], conf)
        w.update(kp.item(), zl.item())
        loss = w.balanced(kp, zl)
        loss.backward()
        opt.step()
        losses.append((loss.item(), kp.item(), zl.item()))
    return sd, names, losses, w


def test_trainer_matches_reference_loop():
    # eps is large on purpose: with the default 1e-8 Adam turns fp32 rounding noise on (mathematically) zero
    # gradients -- e.g. conv biases feeding a train-mode BatchNorm -- into +-lr steps, which no two correct
    # implementations reproduce; the AdamW arithmetic itself is exercised the same way
    lr, wd, steps, eps = 1e-3, 1e-2, 3, 1e-3
    batches = [make_inputs(3, 224, 224, s) for s in range(steps)]
    sd, names, ref_losses, w = reference_loop(batches, lr, wd, steps, eps)
    m = build_model()
    tr = PoseTrainer(m, lr=lr, weight_decay=wd, eps=eps, use_graph=False)
    for s in range(steps):
        b = batches[s]
        loss, kp, zl = tr.step(b["pixel_values"], b["heatmaps"], b["keypoints"], b["z"])
        for got, ref in zip((loss.item(), kp.item(), zl.item()), ref_losses[s]):
            assert abs(got - ref) <= 2e-4 * abs(ref) + 1e-7, (s, got, ref)
    params = dict(m.named_parameters())
    init = make_state_dict(ARCH, 0, 8)
    worst = 0.0
    for n in names:
        a, b = params[n].detach(), sd[n].detach()
        moved = (b - init[n]).norm().item()
        if moved < 1e-7:        # (mathematically) zero gradient: neither implementation moves it
            assert (a - init[n]).norm().item() < 1e-5, n
            continue
        worst = max(worst, ((a - b).norm() / moved).item())   # error relative to the 3-step UPDATE
    assert worst < 5e-2, worst   # fp32 noise of the BN-cancelling gradients through Adam (measured 2.8e-2)
    st = tr.weighting_state
    assert abs(st["kp_loss_avg"] - w.kp_avg) < 1e-5 * abs(w.kp_avg) + 1e-9
    assert abs(st["weight"] - w.weight) < 1e-5
    assert int(tr.step_dev.item()) == steps
    # parameters are views of the flat buffer, laid out in backward-completion order
    lay = tr.layout
    off, k = lay["offsets"][names[0]]
    assert params[names[0]].data_ptr() == tr.flat_params[off:].data_ptr()
    assert set(lay["names"]) == set(names)


def test_trainer_lr_schedule_and_state_dict():
    """the learning rate is a device scalar the recorded AdamW launch reads (a scheduler's `group['lr'] = x`
    takes effect on the next step), and state_dict / load_state_dict round-trip the moments, the step counter, the
    loss-weighting state and the rate (reference train.py:286-318 checkpoints and restores all of these)."""
    batches = [make_inputs(2, 224, 224, s) for s in range(4)]

    def run(tr, rng):
        for s in rng:
            b = batches[s]
            tr.step(b["pixel_values"], b["heatmaps"], b["keypoints"], b["z"])

    # lr = 0 freezes the parameters; raising it through the param group moves them without rebuilding anything
    ta = PoseTrainer(build_model(), lr=0.0, weight_decay=0.0, eps=1e-3, use_graph=False)
    p0 = ta.flat_params.clone()
    run(ta, [0])
    assert torch.equal(ta.flat_params, p0)
    ta.param_groups[0]["lr"] = 1e-3
    assert ta.lr == 1e-3 and abs(float(ta.hyper[0]) - 1e-3) < 1e-9
    run(ta, [1])
    assert (ta.flat_params - p0).abs().max().item() > 0
    # checkpoint after 2 steps, resume in a fresh trainer, compare with an uninterrupted run
    tb = PoseTrainer(build_model(), lr=1e-3, weight_decay=1e-2, eps=1e-3, use_graph=False)
    run(tb, [0, 1])
    ck_model = {k: v.clone() for k, v in tb.model.state_dict().items()}
    ck_opt = tb.state_dict()
    run(tb, [2, 3])
    m2 = build_model()
    m2.load_state_dict(ck_model)
    tc = PoseTrainer(m2.train(), lr=5.0, weight_decay=0.5, eps=1e-3, use_graph=False)   # wrong on purpose: must be restored
    tc.load_state_dict(ck_opt)
    assert tc.lr == 1e-3 and tc.wd == 1e-2 and int(tc.step_dev.item()) == 2
    run(tc, [2, 3])
    assert torch.allclose(tc.flat_params, tb.flat_params, rtol=0, atol=1e-6)
    assert torch.allclose(tc.exp_avg, tb.exp_avg, rtol=1e-5, atol=1e-9)
    assert abs(tc.weighting_state["weight"] - tb.weighting_state["weight"]) < 1e-7
    tc.load_state_dict({}, loss_weight=0.25)      # reference checkpoint['loss_weight']
    assert abs(tc.weighting_state["weight"] - 0.25) < 1e-7


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        full = make_inputs(4, 224, 224, 7)
        shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in full.items()}
        # local gradient of this shard (world-1 semantics), then the data-parallel step
        m1 = build_model()
        t1 = PoseTrainer.__new__(PoseTrainer)      # no process group: plain single-rank trainer
        PoseTrainer.__init__(t1, m1, lr=0.0, weight_decay=0.0, use_graph=False)
        t1.world = 1
        t1.step(shard["pixel_values"], shard["heatmaps"], shard["keypoints"], shard["z"])
        g_local = t1._steps[(2, 224, 224)]["plan"]["gflat"].clone()
        m2 = build_model()
        t2 = PoseTrainer(m2, lr=1e-3, weight_decay=0.0, use_graph=False, bucket_mb=0.5)
        assert t2.world == world
        t2.step(shard["pixel_values"], shard["heatmaps"], shard["keypoints"], shard["z"])
        g_sum = t2._steps[(2, 224, 224)]["plan"]["gflat"].clone()
        gathered = [torch.zeros_like(g_local) for _ in range(world)]
        dist.all_gather(gathered, g_local)
        expect = sum(gathered)
        err = ((g_sum - expect).norm() / expect.norm()).item()
        # buckets: contiguous cover of the flat buffer, more than one, in increasing order
        b = t2.buckets_sent
        ok_cover = b[0][0] == 0 and b[-1][1] == g_sum.numel() and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        # replicas stay identical after the update
        ps = [torch.zeros_like(t2.flat_params) for _ in range(world)]
        dist.all_gather(ps, t2.flat_params)
        same = bool(torch.equal(ps[0], ps[1]))
        moved = ((t2.flat_params - t1.flat_params).abs().max() > 0).item()
        q.put((rank, err, len(b), ok_cover, same, moved))
    finally:
        dist.destroy_process_group()


def test_data_parallel_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nb, ok_cover, same, moved in res:
        assert err < 1e-5, (rank, err)
        assert nb >= 3 and ok_cover, (rank, nb)
        assert same and moved


def test_trainer_unfrozen_layer_matches_reference_loop():
    """PoseTrainer over Dinov2PoseModel(unfreeze_last_n_layers=1) (reference model/dinov2_pose.py:25-39): two AdamW steps
    move the un-frozen encoder layer like the reference loop does (oracle forward, train.py losses, torch AdamW), the
    frozen layer stays put, and the flat layout puts the q / k / v gradient triples in contiguous slices."""
    from dino_pose_b200.model import Dinov2PoseModel
    lr, wd, steps, eps = 1e-3, 1e-2, 2, 1e-3
    batches = [make_inputs(3, 224, 224, s) for s in range(steps)]
    m = Dinov2PoseModel(backbone=ARCH, unfreeze_last_n_layers=1)
    m.load_state_dict(make_state_dict(ARCH, 0, 0))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m._backend_factory = TorchEmulator
    m._act_dtype = torch.float32
    tr = PoseTrainer(m.train(), lr=lr, weight_decay=wd, eps=eps, use_graph=False)
    lay = tr.layout
    ap = "backbone.encoder.layer.1.attention.attention."
    oq, ok_, ov = (lay["offsets"][ap + n + ".weight"][0] for n in ("query", "key", "value"))
    assert ok_ - oq == ov - ok_ == 128 * 128
    for b in batches:
        tr.step(b["pixel_values"], b["heatmaps"], b["keypoints"], b["z"])
    # reference loop
    sd = make_state_dict(ARCH, 0, 0)
    init = {k: v.clone() for k, v in sd.items()}
    names = pose_oracle.trainable_names(sd, None, unfreeze=1, arch=ARCH)
    opt = torch.optim.AdamW([sd[n].requires_grad_(True) for n in names], lr=lr, weight_decay=wd, eps=eps)
    w = pose_oracle.DynamicLossWeighting()
    for b in batches:
        opt.zero_grad(set_to_none=True)
        hm, z = pose_oracle.model_forward(sd, b["pixel_values"], ARCH, None, training=True)
        conf = b["keypoints"][..., 2]
        kp, zl = pose_oracle.keypoint_loss(hm, b["heatmaps"], conf), pose_oracle.z_loss(z, b["z"], conf)
        w.update(kp.item(), zl.item())
        w.balanced(kp, zl).backward()
        opt.step()
    params = dict(m.named_parameters())
    assert set(n for n, p in params.items() if p.requires_grad) == set(names)
    worst = 0.0
    for n in names:
        if not n.startswith("backbone."):
            continue
        a, b = params[n].detach(), sd[n].detach()
        moved = (b - init[n]).norm().item()
        if moved < 1e-7:
            continue
        worst = max(worst, ((a - b).norm() / moved).item())
    assert worst < 8e-2, worst      # error relative to the 2-step UPDATE (fp32 noise through Adam, as in the LoRA test)
    frozen = "backbone.encoder.layer.0.mlp.fc1.weight"
    assert torch.equal(params[frozen].detach(), init[frozen])
