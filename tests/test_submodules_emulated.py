"""Drop-in surface below the pose models (CPU, torch emulator of the C-ABI op vocabulary, fp32 storage): the stand-alone
``forward`` of SpatialAwarePoseHeads / SpatialAwareHeatmapHead / HourglassModule / ZCoordinateHead (reference
model/pose_heads.py:395-400, :345-361, :268-285, :161-162), ``Dinov2Model`` and the attention block ``LoRAAttention``
wraps (model/lora.py:53-65), each against the oracle's restatement of the same module -- outputs, input gradients and
parameter gradients.  The GPU twin of this file is tests/test_submodules_gpu.py."""
import pytest
import torch

from oracle import pose_oracle
from oracle.weights import make_inputs, make_state_dict
from tests.emulator import TorchEmulator

from dino_pose_b200.model import (Dinov2PoseModelLoRA, HourglassModule, SpatialAwareHeatmapHead, SpatialAwarePoseHeads,
                                  ZCoordinateHead)

ARCH = "test/dinov2-tiny"
D = 128


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def emulated(m):
    m._backend_factory = TorchEmulator
    m._act_dtype = torch.float32
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def grad_copy(sd):
    return {k: v.clone().requires_grad_(v.is_floating_point() and "running_" not in k) for k, v in sd.items()}


def sub_state(sd, prefix):
    return {k[len(prefix):]: v.clone() for k, v in sd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("training", [False, True])
def test_pose_heads_module_forward_backward(training):
    sd = make_state_dict(ARCH, 0, 0)
    heads = SpatialAwarePoseHeads(feat_channels=D, num_keypoints=24, heatmap_size=48, spatial_input_size=16,
                                  z_coord_config={"hidden_dims": (1024, 512, 256), "dropout_rate": 0.1})
    heads.load_state_dict(sub_state(sd, "pose_heads."))
    emulated(heads).train(training)
    torch.manual_seed(3)
    x = torch.randn(2, D, 16, 16)
    sdr = grad_copy(sd)
    xr = x.clone().requires_grad_(True)
    rhm, rz = pose_oracle.pose_heads(sdr, xr, training)
    if not training:
        with torch.no_grad():
            hm, z = heads(x)
        assert rel(hm, rhm.detach()) < 1e-4 and rel(z, rz.detach()) < 1e-4
        return
    xg = x.clone().requires_grad_(True)
    hm, z = heads(xg)
    assert rel(hm.detach(), rhm.detach()) < 1e-4 and rel(z.detach(), rz.detach()) < 1e-4
    w_hm, w_z = torch.randn_like(rhm), torch.randn_like(rz)
    ((hm * w_hm).sum() + (z * w_z).sum()).backward()
    ((rhm * w_hm).sum() + (rz * w_z).sum()).backward()
    assert rel(xg.grad, xr.grad) < 2e-3
    checked = 0
    for n, p in heads.named_parameters():
        g_ref = sdr["pose_heads." + n].grad
        if g_ref is None or (n.endswith(".bias") and g_ref.abs().max() < 1e-3):
            continue        # conv biases in front of a train-mode BatchNorm: mathematically zero gradient (fp32 noise)
        assert p.grad is not None, n
        assert rel(p.grad, g_ref) < 5e-3, (n, rel(p.grad, g_ref))
        checked += 1
    assert checked > 40


def test_heatmap_head_hourglass_and_z_head_modules():
    sd = make_state_dict(ARCH, 1, 0)
    torch.manual_seed(5)
    # heat-map head alone, eval
    hmh = SpatialAwareHeatmapHead(feat_channels=D, num_keypoints=24, heatmap_size=48, spatial_input_size=16)
    hmh.load_state_dict(sub_state(sd, "pose_heads.heatmap_head."))
    emulated(hmh).eval()
    x = torch.randn(2, D, 16, 16)
    with torch.no_grad():
        got = hmh(x)
        ref = pose_oracle.heatmap_head(sd, x, False)
    assert got.shape == (2, 24, 48, 48) and rel(got, ref) < 1e-4
    # hourglass alone, train mode, with input and parameter gradients
    hg = HourglassModule(512, 512)
    pfx = "pose_heads.heatmap_head.feature_refine.3."
    hg.load_state_dict(sub_state(sd, pfx))
    emulated(hg).train()
    xh = torch.randn(2, 512, 16, 16)
    sdr = grad_copy(sd)
    xr = xh.clone().requires_grad_(True)
    ref = pose_oracle.hourglass(sdr, xr, pfx, True)
    xg = xh.clone().requires_grad_(True)
    out = hg(xg)
    assert out.shape == ref.shape and rel(out.detach(), ref.detach()) < 1e-4
    w = torch.randn_like(ref)
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert rel(xg.grad, xr.grad) < 2e-3
    for n in ("skip.0.weight", "down1.1.weight", "up2.0.weight", "depthwise_conv.0.weight", "bottleneck.3.weight"):
        assert rel(dict(hg.named_parameters())[n].grad, sdr[pfx + n].grad) < 5e-3, n
    # z head alone
    zh = ZCoordinateHead(D, 24, hidden_dims=(1024, 512, 256), dropout_rate=0.1)
    zh.load_state_dict(sub_state(sd, "pose_heads.z_head."))
    emulated(zh).train()
    f = torch.randn(3, D)
    fr = f.clone().requires_grad_(True)
    ref = pose_oracle.z_head(sdr, fr, True)
    fg = f.clone().requires_grad_(True)
    out = zh(fg)
    assert rel(out.detach(), ref.detach()) < 1e-5
    out.sum().backward()
    ref.sum().backward()
    assert rel(fg.grad, fr.grad) < 1e-4
    assert rel(zh.mlp[0].weight.grad, sdr["pose_heads.z_head.mlp.0.weight"].grad) < 1e-4


def test_backbone_and_attention_modules():
    sd = make_state_dict(ARCH, 0, 8)
    m = Dinov2PoseModelLoRA(backbone=ARCH, lora_rank=8, lora_alpha=16, lora_dropout=0.0)
    m.load_state_dict(sd)
    m.eval()
    inp = make_inputs(2, 224, 224, 0)
    bb = emulated(m.backbone)
    with torch.no_grad():
        out = bb(inp["pixel_values"])
        ref = pose_oracle.backbone(sd, inp["pixel_values"], ARCH, {"rank": 8, "alpha": 16}, False)
    assert out.last_hidden_state.shape == (2, 257, D) and out[0] is out.last_hidden_state
    assert rel(out.last_hidden_state, ref) < 1e-4
    assert torch.equal(out.pooler_output, out.last_hidden_state[:, 0])
    # the wrapped attention block of the last layer, then LoRAAttention on top of it (reference model/lora.py:53-65)
    la = m.backbone.encoder.layer[-1].attention
    emulated(la.original_attention)
    xn = torch.randn(2, 257, D)
    ap = "backbone.encoder.layer.1.attention.original_attention."
    with torch.no_grad():
        got = la.original_attention(xn)[0]
        ref_att = pose_oracle.attention_block(sd, xn, ap, 2)
    assert rel(got, ref_att) < 5e-3      # the stand-alone block packs bf16 weight copies (as the CUDA path always does)
