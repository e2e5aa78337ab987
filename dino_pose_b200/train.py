"""Fine-tuning step of the reference's ``train.py`` hot loop (:134-188) on the sm_100a engine, data-parallel.

One ``PoseTrainer.step`` = ``optimizer.zero_grad`` -> ``model(pixel_values)`` -> reference losses
(``keypoint_loss`` train.py:89-102, ``z_loss`` :109-120, ``DynamicLossWeighting`` :17-69) -> backward ->
gradient all-reduce (world_size > 1) -> AdamW(lr 3e-5, wd 1e-6; train.py:280-284).

How it differs from the reference loop, all deliberate:
  * the whole step is ONE CUDA graph: forward program, ``dp_pose_loss`` (loss + backward seed, loss-weight EMA
    state on the device -- the reference's 7 ``.item()`` host syncs per step, train.py:155-156,173-178, would cap
    multi-GPU scaling), backward program, bucketed NCCL all-reduce, ``dp_adamw``.  The arithmetic is the reference's;
  * the trainable parameters (heads + LoRA) are re-homed as views into ONE flat fp32 buffer, laid out in the order
    the backward program finishes their gradients (``PoseEngine.layout``); gradients and AdamW moments use the same
    layout, so the optimizer is one element-wise kernel and a gradient bucket is a contiguous slice;
  * data parallelism: one process per GPU, the batch dimension is sharded, only these ~31 MB (ViT-S) of fp32
    gradients cross NVLink; every bucket's all-reduce is launched on a side stream as soon as the backward has
    finished that prefix of the flat buffer and overlaps the rest of the backward;
  * BatchNorm uses per-replica batch statistics (torch DDP default); running statistics are per rank.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def keypoint_loss(pred, target, conf):
    """reference train.py:89-102 (torch form, used by tests and by callers that own their loop)."""
    mask = (conf > 1).to(pred.dtype)[:, :, None, None]
    diff = (pred - target) ** 2
    return (torch.exp(-diff.detach()) * diff * mask).mean()


def z_loss(pred_z, target_z, conf):
    """reference train.py:109-120 (masked L1)."""
    mask = (conf > 1).to(pred_z.dtype)
    return (pred_z * mask - target_z * mask).abs().mean()


class _ParamGroup(dict):
    """``optimizer.param_groups[0]`` look-alike: assigning ``group['lr']`` (what torch's schedulers do) updates the device
    scalar the AdamW kernel reads."""

    def __init__(self, trainer, **kw):
        super().__init__(**kw)
        self._trainer = trainer

    def __setitem__(self, key, value):
        if key == "lr":
            self._trainer.set_lr(lr=value)
        elif key == "weight_decay":
            self._trainer.set_lr(weight_decay=value)
        else:
            super().__setitem__(key, value)


class PoseTrainer:
    """AdamW fine-tuning of a ``Dinov2PoseModelLoRA`` / ``Dinov2PoseModel`` on this rank's shard of the batch.

    ``step(pixel_values, target_heatmaps, keypoints, target_z)`` returns ``(loss, kp_loss, z_loss)`` as device
    scalars (views of a static 3-float buffer: read them before the next step, or ``.clone()``)."""

    def __init__(self, model, lr=3e-5, weight_decay=1e-6, betas=(0.9, 0.999), eps=1e-8, bucket_mb=8.0, use_graph=True,
                 process_group=None):
        self.model = model
        self.device = next(model.parameters()).device
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        # DP_BUCKET_MB / DP_NO_ALLREDUCE: A/B switches for the scaling analysis (profiles/r2_scaling.md); the second one skips
        # the gradient exchange altogether (replicas diverge: measurement only)
        bucket_mb = float(os.environ.get("DP_BUCKET_MB", bucket_mb))
        self.no_allreduce = bool(int(os.environ.get("DP_NO_ALLREDUCE", "0")))
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20) / 4))
        self.use_graph = use_graph and self.device.type == "cuda"
        # the exchange runs on its own HIGH-priority stream: NCCL's CTAs are scheduled ahead of the backward's whenever an SM
        # frees up (DP_COMM_PRIORITY=0: default priority, A/B)
        prio = -1 if int(os.environ.get("DP_COMM_PRIORITY", "-1")) < 0 else 0
        self.comm_stream = (torch.cuda.Stream(device=self.device, priority=prio)
                            if (self.world > 1 and self.device.type == "cuda") else None)
        # SMs the persistent GEMM grids of the backward leave to NCCL once the first bucket is in flight (0 = none; A/B in
        # profiles/r2_scaling.md)
        self.reserve_sms = int(os.environ.get("DP_COMM_RESERVE_SMS", "0")) if self.comm_stream is not None else 0
        self.buckets_sent = []          # [(lo, hi)] of the last step, for tests / introspection
        self._flatten_parameters()
        dev = self.device
        n = self.layout["total"]
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.step_dev = torch.zeros((), dtype=torch.int64, device=dev)
        # learning rate / weight decay live on the device: the recorded (and graph-captured) AdamW launch reads them there,
        # so a scheduler (reference train.py:286-293, ReduceLROnPlateau(factor=0.7)) can change them between steps
        self.hyper = torch.tensor([lr, weight_decay], dtype=torch.float32, device=dev)
        self.param_groups = [_ParamGroup(self, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)]
        self.loss_sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self.loss_state = torch.tensor([0.0, 0.0, 0.0, 0.1], device=dev)   # kp_avg, z_avg, started, weight (train.py:18)
        self.loss_out = torch.zeros(3, device=dev)
        self.loss_scales = torch.zeros(2, device=dev)
        self._steps = {}     # (B, H, W) [+ (mode,) for heads-in-eval steps] -> dict(plan, programs, static inputs, graph)

    # ------------------------------------------------------------------ parameters
    def _flatten_parameters(self):
        model = self.model
        model.train()
        eng = model._get_engine(self.device)
        self.layout = lay = eng.layout()
        flat = torch.zeros(lay["total"], device=self.device)
        params = dict(model.named_parameters())
        with torch.no_grad():
            for name in lay["names"]:
                off, k = lay["offsets"][name]
                p = params[name]
                flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + k].view(p.shape)
        self.flat_params = flat
        model._engine = None        # plans recorded on the old parameter storage are stale
        self.engine = model._get_engine(self.device)

    # ------------------------------------------------------------------ optimizer surface (torch.optim.AdamW look-alike)
    def set_lr(self, lr=None, weight_decay=None):
        """Change the learning rate / weight decay of every following step (also of an already captured graph)."""
        if lr is not None:
            self.lr = float(lr)
            self.hyper[0:1].fill_(self.lr)
            dict.__setitem__(self.param_groups[0], "lr", self.lr)
        if weight_decay is not None:
            self.wd = float(weight_decay)
            self.hyper[1:2].fill_(self.wd)
            dict.__setitem__(self.param_groups[0], "weight_decay", self.wd)

    def state_dict(self):
        """Everything `train.py:304-318` checkpoints about the optimizer side: AdamW moments and step counter (per
        parameter NAME, so a checkpoint survives a change of the flat layout), hyper-parameters, the loss-weighting
        state (`checkpoint['loss_weight']`, DynamicLossWeighting :17-69) and the dropout seed counter."""
        lay = self.layout
        per = {}
        for name in lay["names"]:
            off, k = lay["offsets"][name]
            shape = tuple(dict(self.model.named_parameters())[name].shape)
            per[name] = {"exp_avg": self.exp_avg[off:off + k].detach().clone().view(shape),
                         "exp_avg_sq": self.exp_avg_sq[off:off + k].detach().clone().view(shape)}
        ls = self.loss_state.tolist()
        return {"state": per, "step": int(self.step_dev.item()),
                "param_groups": [{"lr": self.lr, "weight_decay": self.wd, "betas": tuple(self.betas), "eps": self.eps}],
                "loss_weighting": {"kp_loss_avg": ls[0], "z_loss_avg": ls[1], "started": ls[2], "weight": ls[3]},
                "dropout_seed": int(self.engine.seed.item()) if self.engine.seed is not None else 0}

    def load_state_dict(self, sd, loss_weight=None):
        """Inverse of ``state_dict``; ``loss_weight`` alone restores the reference checkpoint's `loss_weight` entry."""
        lay = self.layout
        with torch.no_grad():
            for name, st in sd.get("state", {}).items():
                if name not in lay["offsets"]:
                    raise KeyError(f"optimizer state for unknown / non-trainable parameter {name!r}")
                off, k = lay["offsets"][name]
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            if "step" in sd:
                self.step_dev.fill_(int(sd["step"]))
            for g in sd.get("param_groups", [])[:1]:
                if tuple(g.get("betas", self.betas)) != tuple(self.betas) or g.get("eps", self.eps) != self.eps:
                    # betas / eps are immediates of the recorded launch: re-record the optimizer programs
                    self.betas, self.eps = tuple(g.get("betas", self.betas)), g.get("eps", self.eps)
                    self._steps = {}
                self.set_lr(g.get("lr"), g.get("weight_decay"))
            lw = sd.get("loss_weighting")
            if lw is not None:
                self.loss_state.copy_(torch.tensor([lw["kp_loss_avg"], lw["z_loss_avg"], lw["started"], lw["weight"]]))
            if loss_weight is not None:
                self.loss_state[3:4].fill_(float(loss_weight))
            if "dropout_seed" in sd and self.engine.seed is not None:
                self.engine.seed.fill_(int(sd["dropout_seed"]))

    @property
    def weighting_state(self):
        """(kp_loss_avg, z_loss_avg, weight) of the reference's DynamicLossWeighting, read back from the device."""
        s = self.loss_state.tolist()
        return {"kp_loss_avg": s[0] if s[2] else None, "z_loss_avg": s[1] if s[2] else None, "weight": s[3]}

    # ------------------------------------------------------------------ step construction
    def _build(self, B, H, W, mode=True):
        eng, be, dev = self.engine, self.engine.be, self.device
        plan = eng.get_plan(B, H, W, mode)     # mode 2: heads in eval mode (frozen BatchNorm statistics, see engine.get_plan)
        K, hm = eng.K, plan["t"]["hm"]
        st = {"plan": plan}
        st["thm"] = torch.zeros_like(hm)
        st["kps"] = torch.zeros(B, K, 3, device=dev)
        st["tz"] = torch.zeros(B, K, device=dev)
        st["loss"] = be.begin()
        be.pose_loss(hm, st["thm"], st["kps"], plan["t"]["z"], st["tz"], self.loss_sums, self.loss_state, self.loss_out,
                     self.loss_scales, plan["t"]["dhm"], plan["t"]["dz"], B=B, K=K, HW=hm.shape[2] * hm.shape[3])
        # AdamW in two launches over disjoint slices of the flat buffers: [0, split) = the head parameters, whose gradients are
        # final (and, data-parallel, all-reduced) while the backbone part of the backward (final LayerNorm, last block's MLP,
        # LoRA / un-frozen layers) is still running -- that slice is updated on the side stream underneath it; [split, total)
        # follows the last gradient.  DP_SPLIT_ADAMW=0: one launch at the end (A/B).
        total, split = self.layout["total"], self.layout["group_end"].get("fr0", 0)
        if not (0 < split < total) or not bool(int(os.environ.get("DP_SPLIT_ADAMW", "1"))):
            split = 0
        st["opt_split"] = split
        kw = dict(lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.wd,
                  grad_scale=1.0 / self.world, hyper=self.hyper)
        g = plan["gflat"]
        if split:
            st["opt_head"] = be.begin()
            be.adamw(self.flat_params[:split], g[:split], self.exp_avg[:split], self.exp_avg_sq[:split], self.step_dev,
                     n=split, bump=False, **kw)
        st["opt"] = be.begin()
        be.adamw(self.flat_params[split:], g[split:], self.exp_avg[split:], self.exp_avg_sq[split:], self.step_dev,
                 n=total - split, **kw)
        st["graph"] = None
        return st

    def _on_mark(self, flat, st=None):
        """Bucketing: called by the backward program when flat[:upto] is final.  At the head / backbone boundary
        (``st['opt_split']``) the bucket is flushed whatever its size and the head slice's AdamW is issued right behind
        its all-reduce (data-parallel: on the exchange stream; single GPU: on a side stream)."""
        sent = [0]
        total = flat.numel()
        self.buckets_sent = []
        split = st.get("opt_split", 0) if st is not None else 0
        head_done = [False]
        exchange = self.world > 1 and not self.no_allreduce

        def side_stream():
            if self.device.type != "cuda":
                return None
            if self.comm_stream is not None:
                return self.comm_stream
            if getattr(self, "opt_stream", None) is None:
                self.opt_stream = torch.cuda.Stream(device=self.device)
            return self.opt_stream

        def on_mark(tag):
            kind, upto = tag
            if kind != "grads_final":
                return
            at_split = split and upto >= split and not head_done[0]
            if exchange and (upto - sent[0] >= self.bucket_elems or upto >= total or at_split) and upto > sent[0]:
                lo, hi = sent[0], upto
                sent[0] = hi
                self.buckets_sent.append((lo, hi))
                if self.comm_stream is not None:
                    cur = torch.cuda.current_stream()
                    self.comm_stream.wait_stream(cur)
                    with torch.cuda.stream(self.comm_stream):
                        dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                    if self.reserve_sms and not st.get("reserved"):
                        # from the first exchange on, the persistent GEMM grids of the rest of the backward leave SMs to
                        # NCCL's CTAs (see dp_set_reserved_sms); _run() gives them back after the backward
                        st["reserved"] = True
                        self.engine.be.lib.dp_set_reserved_sms(self.reserve_sms)
                else:   # CPU (gloo) test path
                    dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            if at_split:
                head_done[0] = True
                ss = side_stream()
                if ss is None:
                    st["opt_head"].run()
                else:
                    if ss is not self.comm_stream or not exchange:
                        ss.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(ss):
                        st["opt_head"].run()
                    st["side_used"] = ss
        return on_mark

    @torch.no_grad()
    def _run(self, st):
        """forward -> loss + seeds -> backward (+ overlapped all-reduce) -> AdamW, on the current stream."""
        eng, plan = self.engine, st["plan"]
        eng.seed.add_(1)
        plan["fwd"].run()
        st["loss"].run()
        st["side_used"] = None
        st["reserved"] = False
        try:
            eng.backward(plan, "static", "static", on_mark=self._on_mark(plan["gflat"], st))
        finally:
            if st.get("reserved"):
                self.engine.be.lib.dp_set_reserved_sms(0)
        if self.comm_stream is not None and not self.no_allreduce:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        elif st.get("side_used") is not None:
            torch.cuda.current_stream().wait_stream(st["side_used"])
        st["opt"].run()

    def _capture(self, st):
        self._run(st)                     # warm-up outside capture (lazy kernel attributes, NCCL communicators)
        self._run(st)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                self._run(st)
        torch.cuda.current_stream().wait_stream(side)
        st["graph"] = g

    def _snapshot(self):
        return [t.clone() for t in (self.flat_params, self.exp_avg, self.exp_avg_sq, self.step_dev, self.loss_state,
                                    self.engine.seed)] + [b.clone() for b in self.model.buffers()]

    def _restore(self, snap):
        own = [self.flat_params, self.exp_avg, self.exp_avg_sq, self.step_dev, self.loss_state, self.engine.seed]
        for t, s in zip(own + list(self.model.buffers()), snap):
            t.copy_(s)

    def _load_inputs(self, st, tensors):
        """Copy this step's inputs into the static buffers the recorded programs read.  Host tensors go through a
        double-buffered staging area on a copy stream, so the H2D transfer of step k overlaps the still-running
        graph of step k-1 (``step`` returns as soon as the replay is enqueued); the compute stream then only does a
        device-to-device copy."""
        static = (st["plan"]["t"]["px"], st["thm"], st["kps"], st["tz"])
        if self.device.type != "cuda" or all(t.is_cuda for t in tensors):
            for d, t in zip(static, tensors):
                d.copy_(t, non_blocking=True)
            return
        if "staging" not in st:
            st["staging"] = [[torch.empty_like(d) for d in static] for _ in range(2)]
            st["stage_free"] = [torch.cuda.Event(), torch.cuda.Event()]
            st["stage_idx"] = 0
            self.copy_stream = getattr(self, "copy_stream", None) or torch.cuda.Stream(device=self.device)
            for ev in st["stage_free"]:
                ev.record()
        i = st["stage_idx"] = 1 - st["stage_idx"]
        cur, cs = torch.cuda.current_stream(), self.copy_stream
        cs.wait_event(st["stage_free"][i])          # the D2D copy that last read this staging slot has finished
        with torch.cuda.stream(cs):
            for d, t in zip(st["staging"][i], tensors):
                d.copy_(t, non_blocking=True)
        cur.wait_stream(cs)
        for d, sbuf in zip(static, st["staging"][i]):
            d.copy_(sbuf, non_blocking=True)
        st["stage_free"][i].record(cur)

    def step(self, pixel_values, target_heatmaps, keypoints, target_z):
        """One fine-tuning step on this rank's shard.  Inputs may live on the host (pinned) or on the device."""
        model = self.model
        if not model.training:
            model.train()
        self.engine.check_frozen()
        B, _, H, W = pixel_values.shape
        mode = model.engine_mode() if hasattr(model, "engine_mode") else True
        key = (B, H, W) if mode is True else (B, H, W, mode)
        st = self._steps.get(key)
        if st is None or st["plan"] is not self.engine.plans.get((B, H, W, True)):
            st = self._steps[key] = self._build(B, H, W, mode)
        self._load_inputs(st, (pixel_values, target_heatmaps, keypoints, target_z))
        if not self.use_graph:
            self._run(st)
        else:
            if st["graph"] is None:
                snap = self._snapshot()     # the two warm-up runs and the capture must not count as training steps
                self._capture(st)
                self._restore(snap)
            st["graph"].replay()
        return self.loss_out[0], self.loss_out[1], self.loss_out[2]
