"""GPU twin of tests/test_submodules_emulated.py: the stand-alone ``forward`` (and backward) of the head modules,
``Dinov2Model`` and the attention block through the C ABI, against the oracle's fp32 restatement on the host.
Tolerances are the bf16 ones of DESIGN.md (max-rel 2e-2 eval, 4e-2 train-mode heat-maps; gradients by rel-L2)."""
import pytest
import torch

from oracle import pose_oracle
from oracle.weights import make_inputs, make_state_dict

from dino_pose_b200.model import (Dinov2PoseModelLoRA, HourglassModule, SpatialAwareHeatmapHead, SpatialAwarePoseHeads,
                                  ZCoordinateHead)

pytestmark = pytest.mark.gpu
ARCH, D = "facebook/dinov2-small", 384


def dev():
    return torch.device("cuda:0")


def relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def sub_state(sd, prefix):
    return {k[len(prefix):]: v.clone() for k, v in sd.items() if k.startswith(prefix)}


def grad_copy(sd):
    return {k: v.clone().requires_grad_(v.is_floating_point() and "running_" not in k) for k, v in sd.items()}


def no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def test_pose_heads_module_eval_and_train():
    sd = make_state_dict(ARCH, 0, 0)
    heads = SpatialAwarePoseHeads(feat_channels=D, num_keypoints=24, heatmap_size=48, spatial_input_size=16,
                                  z_coord_config={"hidden_dims": (1024, 512, 256), "dropout_rate": 0.1})
    heads.load_state_dict(sub_state(sd, "pose_heads."))
    no_dropout(heads).to(dev())
    torch.manual_seed(3)
    x = torch.randn(8, D, 16, 16)
    heads.eval()
    with torch.no_grad():
        hm, z = heads(x.to(dev()))
        rhm, rz = pose_oracle.pose_heads(sd, x, False)
    assert hm.shape == (8, 24, 48, 48) and z.shape == (8, 24)
    assert relmax(hm, rhm) < 2e-2 and relmax(z, rz) < 2e-2
    heads.train()
    sdr = grad_copy(sd)
    xr = x.clone().requires_grad_(True)
    rhm, rz = pose_oracle.pose_heads(sdr, xr, True)
    xg = x.to(dev()).requires_grad_(True)
    hm, z = heads(xg)
    assert relmax(hm, rhm) < 4e-2 and relmax(z, rz) < 2e-2
    torch.manual_seed(4)
    w_hm, w_z = torch.randn_like(rhm), torch.randn_like(rz)
    ((hm * w_hm.to(dev())).sum() + (z * w_z.to(dev())).sum()).backward()
    ((rhm * w_hm).sum() + (rz * w_z).sum()).backward()
    # gradients through 14 train-mode BatchNorms at batch 8 in bf16: the documented bound of tests/test_model_gpu.py
    assert rel2(xg.grad, xr.grad) < 0.35
    g = {n: p.grad for n, p in heads.named_parameters()}
    for n in ("heatmap_head.prediction.3.weight", "heatmap_head.prediction.0.weight", "heatmap_head.upsampling.1.0.weight",
              "z_head.mlp.0.weight", "z_head.mlp.9.weight"):
        tol = 5e-2 if n.startswith("z_head") else 0.35   # z head: fp32 kernels on the mean of the bf16 feature map
        assert rel2(g[n], sdr["pose_heads." + n].grad) < tol, (n, rel2(g[n], sdr["pose_heads." + n].grad))


def test_heatmap_head_hourglass_z_head_modules():
    sd = make_state_dict(ARCH, 1, 0)
    torch.manual_seed(5)
    hmh = SpatialAwareHeatmapHead(feat_channels=D, num_keypoints=24, heatmap_size=48, spatial_input_size=16)
    hmh.load_state_dict(sub_state(sd, "pose_heads.heatmap_head."))
    hmh.to(dev()).eval()
    x = torch.randn(4, D, 16, 16)
    with torch.no_grad():
        got = hmh(x.to(dev()))
        ref = pose_oracle.heatmap_head(sd, x, False)
    assert got.shape == (4, 24, 48, 48) and relmax(got, ref) < 2e-2
    # 448 x 448 geometry: 32 x 32 feature map -> 96 x 96 -> 2x2 mean -> 48 x 48
    x32 = torch.randn(2, D, 32, 32)
    with torch.no_grad():
        got = hmh(x32.to(dev()))
        ref = pose_oracle.heatmap_head(sd, x32, False)
    assert got.shape == (2, 24, 48, 48) and relmax(got, ref) < 2e-2
    hg = HourglassModule(512, 512)
    pfx = "pose_heads.heatmap_head.feature_refine.3."
    hg.load_state_dict(sub_state(sd, pfx))
    hg.to(dev()).train()
    xh = torch.randn(8, 512, 16, 16)
    sdr = grad_copy(sd)
    xr = xh.clone().requires_grad_(True)
    ref = pose_oracle.hourglass(sdr, xr, pfx, True)
    xg = xh.to(dev()).requires_grad_(True)
    out = hg(xg)
    assert out.shape == ref.shape and relmax(out, ref) < 2e-2
    w = torch.randn_like(ref)
    (out * w.to(dev())).sum().backward()
    (ref * w).sum().backward()
    assert rel2(xg.grad, xr.grad) < 0.35
    for n in ("skip.0.weight", "up2.0.weight", "depthwise_conv.3.weight"):
        assert rel2(dict(hg.named_parameters())[n].grad, sdr[pfx + n].grad) < 0.35, n
    zh = ZCoordinateHead(D, 24, hidden_dims=(1024, 512, 256), dropout_rate=0.1)
    zh.load_state_dict(sub_state(sd, "pose_heads.z_head."))
    no_dropout(zh).to(dev()).train()
    f = torch.randn(16, D)
    fr = f.clone().requires_grad_(True)
    ref = pose_oracle.z_head(sdr, fr, True)
    fg = f.to(dev()).requires_grad_(True)
    out = zh(fg)
    assert relmax(out, ref) < 1e-3              # the z head is fp32 end to end (CUDA-core kernels)
    out.sum().backward()
    ref.sum().backward()
    assert rel2(fg.grad, fr.grad) < 1e-3
    assert rel2(zh.mlp[0].weight.grad, sdr["pose_heads.z_head.mlp.0.weight"].grad) < 1e-3


def test_backbone_and_attention_modules():
    sd = make_state_dict(ARCH, 0, 8)
    m = Dinov2PoseModelLoRA(backbone=ARCH, lora_rank=8, lora_alpha=16, lora_dropout=0.0)
    m.load_state_dict(sd)
    m.to(dev()).eval()
    inp = make_inputs(2, 224, 224, 0)
    with torch.no_grad():
        out = m.backbone(inp["pixel_values"].to(dev()))
        ref = pose_oracle.backbone(sd, inp["pixel_values"], ARCH, {"rank": 8, "alpha": 16}, False)
    assert out.last_hidden_state.shape == (2, 257, D)
    assert rel2(out.last_hidden_state, ref) < 1e-2 and relmax(out.last_hidden_state, ref) < 3e-2
    la = m.backbone.encoder.layer[-1].attention
    torch.manual_seed(0)
    xn = torch.randn(2, 257, D)
    ap = "backbone.encoder.layer.11.attention."
    with torch.no_grad():
        got = la.original_attention(xn.to(dev()))[0]
        ref_att = pose_oracle.attention_block(sd, xn, ap + "original_attention.", 6)
        full = la(xn.to(dev()))[0]
        ref_full = ref_att + pose_oracle.lora_delta(sd, ref_att, ap + "lora_output.", 16, 8)
    assert rel2(got, ref_att) < 1e-2
    assert rel2(full, ref_full) < 1e-2
