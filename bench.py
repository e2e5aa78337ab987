#!/usr/bin/env python
"""Benchmark of the dino_pose hot path on B200 (contract: see the task description / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

N > 1 is launched by the driver with torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.

Default workload = BASELINE.json configs[1], the configuration the metric is quoted on:
  DINOv2-S/14 + LoRA fine-tuning step -- forward, reference losses (train.py:89-120), backward of the heads +
  final LayerNorm + last block's MLP branch + LoRA adapter, AdamW(lr 3e-5, wd 1e-6) -- 24 key-points, 224x224,
  batch 64 per GPU, synthetic images, random-init weights of the named architecture (no network).
--config selects the other BASELINE configurations (supporting numbers, same JSON shape):
  infer_s_b1 (configs[0]), train_s (configs[1], default), train_b (configs[2]), infer_l (configs[3]),
  infer_s448 / train_s448 (configs[4]).  Inference configurations include the key-point decode.
"""
from __future__ import annotations

import argparse
import json
import os

# random-init weights of the named architecture (BASELINE.json north_star: no network, no checkpoints): explicit opt-in
os.environ.setdefault("DINO_POSE_RANDOM_INIT", "1")
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "images/s"
S, B_, L_ = "facebook/dinov2-small", "facebook/dinov2-base", "facebook/dinov2-large"
# algorithmic GFLOP per image: BASELINE.md section 3
CONFIGS = {
    "infer_s_b1": dict(idx=0, arch=S, lora=False, batch=1, res=224, mode="infer", gflop=16.04, cpu_batch=1,
                       text="dinov2-small (ViT-S/14) frozen backbone + heads, 24 keypoints, 224x224, batch 1 inference + decode"),
    "train_s": dict(idx=1, arch=S, lora=True, batch=64, res=224, mode="train", gflop=24.3, cpu_batch=64,
                    text="dinov2-small (ViT-S/14) + LoRA fine-tuning step, 24 keypoints, 224x224, batch 64 per GPU"),
    "train_b": dict(idx=2, arch=B_, lora=True, batch=128, res=224, mode="train", gflop=62.9, cpu_batch=16,
                    text="dinov2-base (ViT-B/14) + LoRA fine-tuning step, 24 keypoints, 224x224, batch 128 per GPU"),
    "infer_l": dict(idx=3, arch=L_, lora=False, batch=256, res=224, mode="infer", gflop=167.33, cpu_batch=8,
                    text="dinov2-large (ViT-L/14) frozen-backbone inference + heads + decode, 224x224, batch 256 per GPU"),
    "infer_s448": dict(idx=4, arch=S, lora=True, batch=64, res=448, mode="infer", gflop=78.64, cpu_batch=8,
                       text="dinov2-small 448x448 (1025 tokens) pose inference + decode, batch 64 per GPU"),
    "train_s448": dict(idx=4, arch=S, lora=True, batch=64, res=448, mode="train", gflop=111.7, cpu_batch=8,
                       text="dinov2-small 448x448 (1025 tokens) + LoRA fine-tuning step, batch 64 per GPU"),
}


def metric_name(cfg):
    if cfg["mode"] == "train":
        return (f"images/sec, DINOv2-{cfg['arch'].split('-')[-1][0].upper()}/14 pose LoRA fine-tuning step (fwd + loss + LoRA/heads bwd "
                f"+ AdamW), {cfg['res']}x{cfg['res']}")
    return f"images/sec, DINOv2-{cfg['arch'].split('-')[-1][0].upper()}/14 pose inference (backbone + heads + decode), {cfg['res']}x{cfg['res']}"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        pw = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ oracle-port steps
def oracle_step_factory(cfg, batch, device, autocast=None, threads=None):
    """The reference's arithmetic for this workload as the ORACLE PORT (oracle/pose_oracle.py: the reference modules
    restated on ATen ops, pinned to the reference by tests/golden): torch eager, fp32 parameters.  device = cpu is the
    reference's own CPU path (cpu_baseline / --impl reference); device = cuda is the "torch on this B200" comparator."""
    import torch
    from oracle import pose_oracle
    from oracle.weights import make_inputs, make_state_dict
    if threads:
        torch.set_num_threads(threads)
    rank = 8 if cfg["lora"] else 0
    sd = {k: v.to(device) for k, v in make_state_dict(cfg["arch"], 0, rank).items()}
    lora = {"rank": 8, "alpha": 16, "dropout": 0.0} if cfg["lora"] else None
    inp = {k: v.to(device) for k, v in make_inputs(batch, cfg["res"], cfg["res"], 0).items()}
    ctx = (lambda: torch.autocast(device_type="cuda", dtype=torch.bfloat16)) if autocast else (lambda: torch.autocast("cpu", enabled=False))
    if cfg["mode"] == "infer":
        def step():
            with torch.no_grad(), ctx():
                hm, z = pose_oracle.model_forward(sd, inp["pixel_values"], cfg["arch"], lora, training=False)
            return hm
        return step
    names = pose_oracle.trainable_names(sd, lora)
    params = [sd[n].requires_grad_(True) for n in names]
    opt = torch.optim.AdamW(params, lr=3e-5, weight_decay=1e-6)
    w = pose_oracle.DynamicLossWeighting()

    def step():
        opt.zero_grad(set_to_none=True)
        with ctx():
            hm, z = pose_oracle.model_forward(sd, inp["pixel_values"], cfg["arch"], lora, training=True, z_dropout=0.1)
        hm, z = hm.float(), z.float()
        conf = inp["keypoints"][..., 2]
        kp = pose_oracle.keypoint_loss(hm, inp["heatmaps"], conf)
        zl = pose_oracle.z_loss(z, inp["z"], conf)
        w.update(kp.item(), zl.item())          # the reference's host syncs (train.py:155-156) are part of its step
        loss = w.balanced(kp, zl)
        loss.backward()
        opt.step()
        return loss
    return step


def time_cpu(cfg, batch, steps, warmup, threads):
    step = oracle_step_factory(cfg, batch, "cpu", threads=threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = cfg["cpu_batch"]
    ips, dt = time_cpu(cfg, batch, max(1, args.steps), max(1, args.warmup), threads)
    bounded = "" if batch == cfg["batch"] else f" (bounded sample: batch {batch} per step instead of {cfg['batch']})"
    sample = f"oracle port (torch eager fp32, reference arithmetic), batch {batch} per step, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(cfg), "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['text']} (BASELINE.json configs[{cfg['idx']}])", "name": args.config,
                   "global_batch": batch, "arm": f"reference arithmetic on the host CPU (oracle port){bounded}"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ our arm
def family_graph_ms(progs, kernel, reps=5):
    """GPU time of ONE kernel family: a CUDA graph holding only that family's launches of one step (step order, buffers of
    the last real step), replayed back to back, timed with events on the launching stream.  Returns (ms per replay, launches)."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for p in progs:
            p.run_family(kernel)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            n = sum(p.run_family(kernel) for p in progs)
        g.replay()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(side)
        for _ in range(reps):
            g.replay()
        f1.record(side)
        torch.cuda.synchronize()
    torch.cuda.current_stream().wait_stream(side)
    return f0.elapsed_time(f1) / reps, n


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warnings.filterwarnings("ignore", message=".*RANDOMLY initialised.*")

    from dino_pose_b200.backend import CudaBackend
    from dino_pose_b200.model import Dinov2PoseModel, Dinov2PoseModelLoRA
    from dino_pose_b200.src.model_utils import decode_heatmaps
    from dino_pose_b200.synthetic import make_inputs
    from dino_pose_b200.train import PoseTrainer

    B = args.batch or cfg["batch"]
    res, train = cfg["res"], cfg["mode"] == "train"
    torch.manual_seed(0)                      # identical random-init replica on every rank
    if cfg["lora"]:
        model = Dinov2PoseModelLoRA(num_keypoints=24, backbone=cfg["arch"], heatmap_size=48, lora_rank=8, lora_alpha=16,
                                    lora_dropout=0.1).to(dev)
    else:
        model = Dinov2PoseModel(num_keypoints=24, backbone=cfg["arch"], heatmap_size=48).to(dev)
    host = {k: v.pin_memory() for k, v in make_inputs(B, res, res, seed=rank).items()}
    devb = {k: v.to(dev) for k, v in host.items()}
    trainer = None
    if train:
        trainer = PoseTrainer(model)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = 4

        def step_resident():
            return trainer.step(devb["pixel_values"], devb["heatmaps"], devb["keypoints"], devb["z"])

        def run_e2e():
            loss, _, _ = trainer.step(host["pixel_values"], host["heatmaps"], host["keypoints"], host["z"])
            return loss.reshape(1)
        result_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    else:
        model.eval()
        h2d = host["pixel_values"].numel() * 4
        d2h = B * 24 * 4 * 8                    # decoded key-points read back: (x, y) float64 + (row, col) per map, 4 x 8 bytes
        px_stage = torch.empty_like(devb["pixel_values"])

        def step_resident():
            with torch.no_grad():
                hm, _z = model(devb["pixel_values"])
                return decode_heatmaps(hm, (res, res))

        def run_e2e():
            with torch.no_grad():
                px_stage.copy_(host["pixel_values"], non_blocking=True)
                hm, _z = model(px_stage)
                idx, xy, _c = decode_heatmaps(hm, (res, res))
            return torch.cat([xy.reshape(-1), idx.reshape(-1).double()])
        result_host = [torch.zeros(B * 24 * 4, dtype=torch.float64).pin_memory() for _ in range(2)]

    ready = [None, None]
    e2e_state = {"k": 0, "last": None}

    def step_e2e():
        """One end-to-end step through the public API: pinned HOST inputs (H2D inside the step) and a D2H read of the step's
        result (training: the loss; inference: the decoded key-points).  The result of step k is read back while step k+1 is
        already enqueued (one-step delayed), so the host never stalls the device; every step's result is read inside the
        timed region (the last one by ``drain_e2e``)."""
        k = e2e_state["k"]
        out = run_e2e()
        slot = k & 1
        result_host[slot][:out.numel()].copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ready[slot] = ev
        prev = ready[slot ^ 1]
        if prev is not None:
            prev.synchronize()
            e2e_state["last"] = float(result_host[slot ^ 1][0])
        e2e_state["k"] = k + 1

    def drain_e2e():
        slot = (e2e_state["k"] - 1) & 1
        if ready[slot] is not None:
            ready[slot].synchronize()
            e2e_state["last"] = float(result_host[slot][0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    warm = max(3, args.warmup)
    for _ in range(warm):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_resident, args.steps)                 # THE number: exactly K steps, max over ranks
    # stability evidence (not the reported value): the same K-step measurement repeated
    repeats = [timed(step_resident, args.steps) for _ in range(4)]
    for _ in range(2):
        step_e2e()
    drain_e2e()
    ms_e2e = timed(step_e2e, args.steps, after=drain_e2e)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=3)

    # data-parallel sanity on the hardware: every rank must hold bit-identical parameters after the timed steps
    replica_diff = None
    if world > 1 and trainer is not None:
        ref = trainer.flat_params.clone()
        dist.broadcast(ref, 0)
        d = (trainer.flat_params - ref).abs().max().reshape(1)
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        replica_diff = float(d.item())

    # ---- per-kernel accounting: one instrumented replay of the recorded programs (CUDA events around every launch, on the
    # launching stream), rank 0, after the timed region
    roof, extras, launches, per_kernel = None, [], 0, {}
    peaks = load_peaks()
    if rank == 0:
        eng = model._get_engine(dev)
        if train:
            st = trainer._steps[(B, res, res)]
            plan = st["plan"]
            # the head slice's AdamW is its own program (issued from the backward's bucket mark, dino_pose_b200/train.py)
            progs = [plan["fwd"], st["loss"], plan["bwd"]] + ([st["opt_head"]] if st.get("opt_head") is not None else []) \
                + [st["opt"]]
        else:
            plan = eng.plans[(B, res, res, False)]
            progs = [plan["fwd"]]
        # decode as a recorded program on the step's heat-maps (inference configurations run it every step; for training
        # configurations it is measured here only for the HBM roofline north_star asks for)
        be = CudaBackend()
        hm = plan["t"]["hm"]
        maps = B * 24
        dbuf = (torch.empty((maps, 2), dtype=torch.int32, device=dev), torch.empty((maps, 2), dtype=torch.float64, device=dev),
                torch.empty((maps,), dtype=torch.float32, device=dev))
        dec = be.begin()
        be.decode(hm, *dbuf, maps=maps, H=48, W=48, target_w=res, target_h=res)
        launches = sum(len(p) for p in progs) + (1 if not train else 0)
        agg = {}
        reps = 3
        per_launch = {}
        for _ in range(reps):
            recs = [r for p in progs for r in p.run_timed()]
            for i, rec in enumerate(recs):
                pl = per_launch.setdefault(i, dict(rec, ms=0.0))
                pl["ms"] += rec["ms"] / reps
            for rec in recs:
                a = agg.setdefault(rec["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
                a["ms"] += rec["ms"]; a["flops"] += rec["flops"]; a["bytes"] += rec["bytes"]; a["n"] += 1
        dump = os.environ.get("DP_BENCH_DUMP")
        if dump:
            with open(dump, "w") as f:
                # stream: 0 = the step's main stream, 1 / 2 = the side streams of the recorded programs (engine.py fork / join)
                f.write("idx,name,kernel,ms,gflop,tflops,mbytes,gbs,stream\n")
                for i in sorted(per_launch):
                    r = per_launch[i]
                    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
                    gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
                    f.write(f'{i},{r["name"]},{r["kernel"]},{r["ms"]:.5f},{r["flops"] / 1e9:.3f},{tf:.1f},'
                            f'{r["bytes"] / 1e6:.2f},{gb:.0f},{int(r.get("side") or 0)}\n')
        per_kernel = {k: round(v["ms"] / reps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:14]}
        name, a = max(agg.items(), key=lambda kv: kv[1]["ms"])
        # The event pair around a single eager launch also contains the launch hand-over (~5 us per pair), so `achieved` is
        # taken from a CUDA graph that holds ONLY this family's launches of one step, replayed back to back.  That replay is
        # a ~10 ms burst of one kernel family at full clocks, hence the BURST peak is the denominator of `frac`; the
        # sustained figure is given beside it.  (Done last: the stray BatchNorm statistics it accumulates are re-zeroed.)
        fam_ms, fam_launches = family_graph_ms(progs, name)
        fam_flops = a["flops"] / reps
        achieved = fam_flops / (fam_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")   # ncu --set full of one step, tools/ncu_summary.py traffic
        if os.path.exists(tpath) and args.config == "train_s":
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(name, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        step_tf = cfg["gflop"] * 1e9 * B / (ms_step * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_burst"], "peak_sustained": peaks["tf_sustained"],
                "frac_sustained": achieved / peaks["tf_sustained"], "traffic": traffic,
                "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over "
                                  "the launches of this kernel family in one step)" if traffic is not None else None,
                "algorithmic_flops_per_launch": a["flops"] / a["n"],
                "peak_source": peaks["source"] + ": burst figure for `frac` (family timed alone in a ~10 ms graph replay), "
                                                 "sustained figure beside it",
                "timing": "CUDA graph of this family's launches of one step (step order, buffers of the last step), "
                          "replayed 5x, events on the launching stream",
                "share_of_step": fam_ms / ms_step, "launches_per_step": fam_launches,
                "avg_launch_ms": fam_ms / max(1, fam_launches),
                "achieved_single_launch_events": a["flops"] / (a["ms"] * 1e-3) / 1e12 if a["ms"] > 0 else 0.0,
                "avg_launch_ms_events": a["ms"] / a["n"],
                "per_kernel_ms_per_step": per_kernel,
                "step_tflops": step_tf, "step_frac_of_peak": step_tf / peaks["tf_burst"],
                "step_frac_of_peak_sustained": step_tf / peaks["tf_sustained"]}
        # the other kernels north_star names: LayerNorm and decode against the HBM roofline, attention against the tensor peak
        for kern, bound in (("layernorm_fwd", "hbm"), ("attention_fwd", "tensor"), ("decode", "hbm"), ("adamw", "hbm")):
            pl = [dec] if kern == "decode" else progs
            src = agg.get(kern)
            if kern != "decode" and src is None:
                continue
            ms_k, n_k = family_graph_ms(pl, kern, reps=20 if kern == "decode" else 5)
            if n_k == 0:
                continue
            if kern == "decode":
                work = maps * (48 * 48 * 4.0 + 28.0)
            else:
                work = (src["bytes"] if bound == "hbm" else src["flops"]) / reps
            if bound == "hbm":
                ach, peak, unit = work / (ms_k * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s"
            else:
                ach, peak, unit = work / (ms_k * 1e-3) / 1e12, peaks["tf_burst"], "TFLOP/s"
            extras.append({"kernel": kern, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                           "launches_per_step": n_k, "avg_launch_ms": ms_k / n_k,
                           "algorithmic_per_launch": work / n_k})
        for Lr in plan.get("layers", {}).values():
            if "sums" in Lr.t:
                Lr.t["sums"].zero_()

    # ---- comparators, rank 0 at N = 1: the reference's CPU path (oracle port) and the same port run by torch on this GPU
    cpu = gpu_cmp = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = cfg["cpu_batch"]
        ips, dt = time_cpu(cfg, cb, 2, 1, threads)
        cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle port (torch eager fp32) of the same workload at batch {cb}"
                         + ("" if cb == B else f" (bounded: the GPU arm runs batch {B})") + ", 1 warm-up + 2 timed steps"}
    if rank == 0 and world == 1 and not args.no_gpu_comparator:
        gpu_cmp = {"what": "oracle port (reference arithmetic on stock ATen / cuBLASLt / cuDNN / SDPA-free eager kernels) on "
                           "this B200, same workload and batch, 2 warm-up + 3 timed steps, CUDA events", "unit": UNIT}
        for label, ac in (("torch_eager_fp32", False), ("torch_autocast_bf16", True)):
            try:
                torch.cuda.empty_cache()
                stp = oracle_step_factory(cfg, B, dev, autocast=ac)
                for _ in range(2):
                    stp()
                torch.cuda.synchronize()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                for _ in range(3):
                    stp()
                c1.record()
                torch.cuda.synchronize()
                ms_c = c0.elapsed_time(c1) / 3
                gpu_cmp[label] = {"value": B / (ms_c * 1e-3), "ms_per_step": ms_c}
                del stp
            except Exception as ex:   # e.g. out of memory at the largest configurations
                gpu_cmp[label] = {"error": f"{type(ex).__name__}: {str(ex)[:120]}"}
        torch.cuda.empty_cache()

    if rank == 0:
        total_b = B * world
        srt = sorted(repeats + [ms_step])
        out = {
            "metric": metric_name(cfg), "value": total_b / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg['text']} (BASELINE.json configs[{cfg['idx']}])", "name": args.config,
                       "global_batch": total_b, "parallelism": f"dp{world}" if train else f"replicas x{world} (no collective)",
                       "l2": "per-step working set (activations + saved tensors) exceeds the 126 MB L2; no explicit flush"
                             if B > 1 else "batch-1 latency run: working set fits L2 by construction (the configuration is the "
                                           "reference's single-image demo path)",
                       "step": ("fwd + reference losses + bwd (heads, final LN, last-block MLP, LoRA) + AdamW, replayed as one "
                                "CUDA graph") if train else "forward program (one CUDA graph) + key-point decode"},
            "per_gpu": B / (ms_step * 1e-3),
            "ms_per_step_repeats": {"all": [round(x, 4) for x in [ms_step] + repeats], "median": srt[len(srt) // 2],
                                    "note": "the reported value is the FIRST K-step measurement; the others repeat it"},
            "e2e": {"value": total_b / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": sampler.summary(),
            "roofline": roof,
            "extra_rooflines": extras,
            "cpu_baseline": cpu,
            "gpu_comparator": gpu_cmp,
        }
        if replica_diff is not None:
            out["replica_param_max_abs_diff"] = replica_diff
        print(json.dumps(out), flush=True)
    if world > 1:
        # leave together, then exit without tearing NCCL down: destroy_process_group() blocks when communicators are
        # still referenced by captured CUDA graphs (observed: rank 0 printed its line, then both ranks hung)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="train_s", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the configuration's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-comparator", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
