"""Stand-alone entry points for the sub-modules of the drop-in surface (called by the containers' ``forward``).

The reference's callers only ever run the whole pose model, but its sub-modules are ordinary ``nn.Module``s whose
``forward`` works on its own (reference model/pose_heads.py:161-162, :268-285, :345-361, :395-400; model/lora.py:26-28,
:53-65; HF ``Dinov2Model.forward``).  Here each of them runs the SAME recorded launches it contributes to the full
model, from its own input buffer: a ``PoseEngine`` over just that module's parameters, with a plan ``scope``
(engine.py ``build_scope_plan`` / ``build_plan(scope="backbone")``).  Layout changes at the boundary (NCHW fp32 <->
NHWC bf16) are torch copies; all arithmetic is in the sm_100a kernels.  There is no CPU path.
"""
from __future__ import annotations

import warnings

import torch

PATCH = 14


def _cuda_only(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: dino_pose_b200 has no CPU execution path; pass CUDA tensors")


def lora_delta(x, lora_A, lora_B, scaling, p_drop):
    """reference model/lora.py:26-28 on an arbitrary [..., in] tensor via ``dp_lora_fwd``
    (x_in = 0, lambda1 = 1, then subtract y): returns dropout(x @ A @ B) * scaling."""
    _cuda_only(x, "LoRALayer.forward")
    from .backend import CudaBackend
    be = CudaBackend()
    prog = be.begin()
    D = x.shape[-1]
    if lora_B.shape[1] != D:
        raise NotImplementedError("LoRALayer kernel path supports in_features == out_features (the reference's only use)")
    y = x.reshape(-1, D).float().contiguous()
    rows = y.shape[0]
    ones = torch.ones(D, device=x.device)
    zeros = torch.zeros_like(y)
    out = torch.empty_like(y)
    seed = torch.randint(0, 2 ** 62, (1,), device=x.device, dtype=torch.int64) if p_drop > 0 else None
    be.lora_fwd(y, lora_A.detach().float().contiguous(), lora_B.detach().float().contiguous(), ones, zeros, out, None,
                rows=rows, D=D, R=lora_A.shape[1], scaling=float(scaling), p_drop=float(p_drop), seed=seed)
    prog.run()
    return (out - y).reshape(x.shape).to(x.dtype)


# ------------------------------------------------------------------------------------------------ engine per module
def _backend_for(module, device):
    factory = getattr(module, "_backend_factory", None)
    if factory is not None:            # tests: torch emulator of the op vocabulary
        return factory()
    if device.type != "cuda":
        raise RuntimeError(f"{type(module).__name__}.forward: dino_pose_b200 runs on CUDA (sm_100a) only; there is no "
                           "CPU execution path")
    from .backend import CudaBackend
    return CudaBackend()


def _module_engine(module, prefix, cfg, device):
    """PoseEngine over ``module``'s own parameters, renamed to the names they have inside the full pose model.  Cached on
    the module; rebuilt when the parameter storage moved (``.to()`` / ``.cuda()`` / ``load_state_dict``)."""
    from .engine import PoseEngine
    params = {prefix + n: p for n, p in module.named_parameters()}
    bufs = {prefix + n: b for n, b in module.named_buffers()}
    key = (str(device), tuple(p.data_ptr() for p in params.values()), repr(sorted(cfg.items())))
    cached = module.__dict__.get("_dp_engine")
    if cached is not None and cached[0] == key:
        return cached[1]
    act = getattr(module, "_act_dtype", None)
    if act is not None:
        cfg = dict(cfg, act_dtype=act)
    eng = PoseEngine(params, bufs, cfg, _backend_for(module, device), device)
    module.__dict__["_dp_engine"] = (key, eng)
    return eng


_HEAD_SCOPES = {
    # kind: (parameter-name prefix inside the pose model, input is a feature map)
    "pose_heads": ("pose_heads.", True),
    "heatmap_head": ("pose_heads.heatmap_head.", True),
    "hourglass": ("pose_heads.heatmap_head.feature_refine.3.", True),
    "z_head": ("pose_heads.z_head.", False),
}


def _head_cfg(module, kind):
    """Engine configuration of a stand-alone head module (geometry checks mirror what the engine's layer table covers)."""
    if kind == "hourglass":
        if (module.in_channels, module.out_channels) != (512, 512):
            raise NotImplementedError("HourglassModule kernel path covers the 512 -> 512 instance the pose heads use "
                                      "(reference model/pose_heads.py:301)")
        return dict(D=512, L=0, heads=1, num_keypoints=24, heatmap_size=48, z_hidden=(), z_dropout=0.0)
    if kind == "z_head":
        return dict(D=module.feat_dim, L=0, heads=1, num_keypoints=module.num_keypoints, heatmap_size=48,
                    z_hidden=tuple(module.hidden_dims), z_dropout=float(module.mlp[2].p) if len(module.mlp) > 2 else 0.0)
    hm = module.heatmap_head if kind == "pose_heads" else module
    if hm.heatmap_size != 48 or hm.spatial_input_size != 16:
        raise NotImplementedError("SpatialAwareHeatmapHead kernel path covers heatmap_size=48, spatial_input_size=16 "
                                  "(reference model/dinov2_pose.py:45-54)")
    cfg = dict(D=hm.feat_channels, L=0, heads=1, num_keypoints=hm.num_keypoints, heatmap_size=hm.heatmap_size,
               z_hidden=(), z_dropout=0.0)
    if kind == "pose_heads":
        zh = module.z_head
        cfg.update(z_hidden=tuple(zh.hidden_dims), z_dropout=float(zh.mlp[2].p) if len(zh.mlp) > 2 else 0.0)
    return cfg


def _load_scope_input(plan, kind, x):
    t, B, g = plan["t"], plan["B"], plan["g"]
    if kind == "z_head":
        t["zin"].copy_(x)
    elif kind == "hourglass":
        t["hg_in"].view(B, g, g, 512).copy_(x.permute(0, 2, 3, 1))
    else:
        t["feat"].view(B, g, g, x.shape[1]).copy_(x.permute(0, 2, 3, 1))


def _scope_outputs(plan, kind):
    t, B, g = plan["t"], plan["B"], plan["g"]
    if kind == "pose_heads":
        return t["hm"].clone(), t["z"].clone()
    if kind == "heatmap_head":
        return (t["hm"].clone(),)
    if kind == "z_head":
        return (t["z"].clone(),)
    return (plan["hgout"].view(B, g, g, 512).permute(0, 3, 1, 2).float().contiguous(),)


def _run_scope_forward(eng, kind, x, training):
    if kind == "z_head":
        if x.dim() != 2:
            raise ValueError("ZCoordinateHead expects features of shape [B, feat_dim]")
        B, g = x.shape[0], 16
    else:
        if x.dim() != 4 or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected a square [B, C, g, g] feature map, got {tuple(x.shape)}")
        B, g = x.shape[0], x.shape[2]
        if g not in (16, 32):
            raise NotImplementedError(f"feature-map side {g}: the head kernels cover 16 (224x224) and 32 (448x448)")
    plan = eng.get_plan(B, g * PATCH, g * PATCH, training, scope=kind)
    _load_scope_input(plan, kind, x)
    if training:
        eng.seed.add_(1)
        plan["generation"] = plan.get("generation", 0) + 1
    plan["fwd"].run()
    return plan


class _ScopeFunction(torch.autograd.Function):
    """Autograd boundary of a stand-alone head module: one forward program, one backward program."""

    @staticmethod
    def forward(ctx, eng, kind, names, x, *params):
        plan = _run_scope_forward(eng, kind, x, True)
        ctx.eng, ctx.kind, ctx.names, ctx.plan, ctx.generation = eng, kind, names, plan, plan["generation"]
        ctx.x_shape = tuple(x.shape)
        return _scope_outputs(plan, kind)

    @staticmethod
    def backward(ctx, *grads):
        plan, kind, t = ctx.plan, ctx.kind, ctx.plan["t"]
        if plan["generation"] != ctx.generation or plan.get("consumed") == ctx.generation:
            raise RuntimeError("dino_pose_b200: backward() of a forward pass whose saved activations are gone (one set of "
                               "activation buffers per module and input shape; run forward / backward pairs in order)")
        plan["consumed"] = ctx.generation
        B, g = plan["B"], plan["g"]

        def seed(buf, grad, nhwc=False):
            if grad is None:
                buf.zero_()
            elif nhwc:
                buf.view(B, g, g, grad.shape[1]).copy_(grad.permute(0, 2, 3, 1))
            else:
                buf.copy_(grad)
        if kind == "pose_heads":
            seed(t["dhm"], grads[0]); seed(t["dz"], grads[1])
        elif kind == "heatmap_head":
            seed(t["dhm"], grads[0])
        elif kind == "z_head":
            seed(t["dz"], grads[0])
        else:
            seed(t["d_hg"], grads[0], nhwc=True)
        plan["bwd"].run()
        d_in = plan["d_in"]
        if kind == "z_head":
            dx = d_in.clone()
        else:
            dx = d_in.view(B, g, g, ctx.x_shape[1]).permute(0, 3, 1, 2).float().contiguous()
        return (None, None, None, dx) + tuple(plan["grads"][n].clone() for n in ctx.names)


def run_head_module(module, kind, x):
    """``forward`` of SpatialAwarePoseHeads / SpatialAwareHeatmapHead / HourglassModule / ZCoordinateHead."""
    prefix, _is_map = _HEAD_SCOPES[kind]
    if getattr(module, "_backend_factory", None) is None:
        _cuda_only(x, f"{type(module).__name__}.forward")
    eng = _module_engine(module, prefix, _head_cfg(module, kind), x.device)
    training = module.training
    named = [(prefix + n, p) for n, p in module.named_parameters() if p.requires_grad]
    wants_graph = torch.is_grad_enabled() and (bool(named) or x.requires_grad)
    if wants_graph and training:
        lay = eng.layout()
        names = [n for n, _ in named if n in lay["offsets"]]
        params = [p for n, p in named if n in lay["offsets"]]
        out = _ScopeFunction.apply(eng, kind, tuple(names), x.float(), *params)
    else:
        if wants_graph and not module.__dict__.get("_warned_eval_grad", False):
            warnings.warn(f"dino_pose_b200: {type(module).__name__}.forward in eval mode with autograd enabled returns outputs "
                          "WITHOUT a graph (the inference program is not differentiable)", RuntimeWarning, stacklevel=3)
            module.__dict__["_warned_eval_grad"] = True
        with torch.no_grad():
            out = _scope_outputs(_run_scope_forward(eng, kind, x.float(), training), kind)
    return out if kind == "pose_heads" else out[0]


# ------------------------------------------------------------------------------------------------ backbone
class BackboneOutput(tuple):
    """``transformers`` ``BaseModelOutputWithPooling`` look-alike: attribute and index access
    (reference model/dinov2_pose.py:145-147 reads ``outputs.last_hidden_state``)."""

    def __new__(cls, last_hidden_state, pooler_output):
        self = super().__new__(cls, (last_hidden_state, pooler_output))
        self.last_hidden_state, self.pooler_output = last_hidden_state, pooler_output
        self.hidden_states = self.attentions = None
        return self


def run_backbone(backbone, pixel_values):
    """Stand-alone ``Dinov2Model.forward`` (HF modeling_dinov2.py:455-500): embeddings, encoder, final LayerNorm over all
    tokens; ``pooler_output`` is the CLS row.  The backbone is frozen in both pose models, so no graph is recorded."""
    if getattr(backbone, "_backend_factory", None) is None:
        _cuda_only(pixel_values, "Dinov2Model.forward")
    from .model.lora import LoRAAttention
    cfg = backbone.config
    last = backbone.encoder.layer[-1].attention
    lora = None
    if isinstance(last, LoRAAttention):
        lora = {"rank": last.rank, "alpha": last.alpha, "dropout": float(last.lora_output.dropout.p)}
    ecfg = dict(D=cfg.hidden_size, L=cfg.num_hidden_layers, heads=cfg.num_attention_heads, num_keypoints=24,
                heatmap_size=48, lora=lora, z_hidden=(), z_dropout=0.0)
    eng = _module_engine(backbone, "backbone.", ecfg, pixel_values.device)
    if torch.is_grad_enabled() and any(p.requires_grad for p in backbone.parameters()) \
            and not backbone.__dict__.get("_warned_grad", False):
        warnings.warn("dino_pose_b200: stand-alone Dinov2Model.forward records no autograd graph; fine-tune through "
                      "Dinov2PoseModel / Dinov2PoseModelLoRA (fused forward + backward)", RuntimeWarning, stacklevel=3)
        backbone.__dict__["_warned_grad"] = True
    B, C, H, W = pixel_values.shape
    if C != 3:
        raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                         f"configuration. Expected 3 but got {C}.")
    with torch.no_grad():
        eng.check_frozen()
        training = bool(backbone.training and lora is not None and lora["dropout"] > 0)
        plan = eng.get_plan(B, H, W, training, scope="backbone")
        plan["t"]["px"].copy_(pixel_values)
        if training:
            eng.seed.add_(1)
        plan["fwd"].run()
        lhs = plan["t"]["lhs"].view(B, plan["T"], cfg.hidden_size).clone()
    return BackboneOutput(lhs, lhs[:, 0])


# ------------------------------------------------------------------------------------------------ attention block
def run_attention(attn, hidden_states):
    """HF ``Dinov2Attention.forward`` (modeling_dinov2.py:203-252): fused QKV projection, multi-head attention, output
    dense; returns ``(context,)`` like HF.  ``attn`` is the parameter container ``model.dinov2_backbone.Attention``."""
    if getattr(attn, "_backend_factory", None) is None:
        _cuda_only(hidden_states, "Attention.forward")
    x = hidden_states
    if x.dim() != 3:
        raise ValueError("hidden_states must be [batch, tokens, hidden]")
    B, T, D = x.shape
    heads = attn.attention.num_attention_heads
    be = _backend_for(attn, x.device)
    sa, so = attn.attention, attn.output
    key = (str(x.device), B, T, tuple(p._version for p in attn.parameters()), tuple(p.data_ptr() for p in attn.parameters()))
    cached = attn.__dict__.get("_dp_attn")
    if cached is None or cached[0] != key:
        bf = torch.bfloat16
        st = {"wqkv": torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0).detach().to(bf).contiguous(),
              "bqkv": torch.cat([sa.query.bias, sa.key.bias, sa.value.bias]).detach().float().contiguous(),
              "wo": so.dense.weight.detach().to(bf).contiguous(), "bo": so.dense.bias.detach().float().contiguous(),
              "xb": torch.empty(B * T, D, device=x.device, dtype=bf), "qkv": torch.empty(B * T, 3 * D, device=x.device, dtype=bf),
              "ctx": torch.empty(B * T, D, device=x.device, dtype=bf), "out": torch.empty(B * T, D, device=x.device)}
        prog = be.begin()
        M = B * T
        be.gemm(st["xb"], st["wqkv"], st["qkv"], M=M, N=3 * D, K=D, bias=st["bqkv"], name="attn.qkv")
        be.attention_fwd(st["qkv"], st["ctx"], B=B, T=T, heads=heads, scale=1.0 / (D // heads) ** 0.5)
        be.gemm(st["ctx"], st["wo"], st["out"], M=M, N=D, K=D, bias=st["bo"], out_dtype="f32", name="attn.dense")
        st["prog"] = prog
        attn.__dict__["_dp_attn"] = cached = (key, st)
    st = cached[1]
    with torch.no_grad():
        st["xb"].copy_(x.reshape(B * T, D))
        st["prog"].run()
        out = st["out"].view(B, T, D).to(x.dtype).clone()
    return (out,)
