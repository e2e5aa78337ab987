#!/usr/bin/env python
"""Benchmark of the dino_pose hot path on B200 (contract: see the task description / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config train_s|infer_s|...]

N > 1 is launched by the driver with torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
  DINOv2-S/14 + LoRA fine-tuning step -- forward, reference losses (train.py:89-120), backward of the heads +
  final LayerNorm + last block's MLP branch + LoRA adapter, AdamW(lr 3e-5, wd 1e-6) -- 24 key-points, 224x224,
  batch 64 per GPU, synthetic images, random-init weights of the named architecture (no network).
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

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec, DINOv2-S/14 pose LoRA fine-tuning step (fwd + loss + LoRA/heads bwd + AdamW), 224x224"
UNIT = "images/s"
ARCH = "facebook/dinov2-small"
FLOPS_PER_IMAGE_STEP = 24.3e9     # BASELINE.md section 3 (ViT-S/14 224^2 LoRA train step, algorithmic)


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
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_step_factory(batch, threads):
    """The reference's own CPU path for this workload = reference modules on torch eager fp32.  The reference
    tree does not travel to the GPU box, so the timed implementation is the oracle PORT (oracle/pose_oracle.py,
    pinned to the reference by tests/golden) -- same ATen ops, same fp32 arithmetic."""
    import torch
    from oracle import pose_oracle
    from oracle.weights import make_inputs, make_state_dict
    torch.set_num_threads(threads)
    sd = make_state_dict(ARCH, 0, 8)
    lora = {"rank": 8, "alpha": 16, "dropout": 0.0}
    names = pose_oracle.trainable_names(sd, lora)
    params = [sd[n].requires_grad_(True) for n in names]
    opt = torch.optim.AdamW(params, lr=3e-5, weight_decay=1e-6)
    inp = make_inputs(batch, 224, 224, 0)
    w = pose_oracle.DynamicLossWeighting()

    def step():
        opt.zero_grad(set_to_none=True)
        hm, z = pose_oracle.model_forward(sd, inp["pixel_values"], ARCH, lora, training=True, z_dropout=0.1)
        conf = inp["keypoints"][..., 2]
        kp = pose_oracle.keypoint_loss(hm, inp["heatmaps"], conf)
        zl = pose_oracle.z_loss(z, inp["z"], conf)
        w.update(kp.item(), zl.item())
        loss = w.balanced(kp, zl)
        loss.backward()
        opt.step()
        return loss.item()
    return step


def time_cpu(batch, steps, warmup, threads):
    step = cpu_reference_step_factory(batch, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 16
    ips, dt = time_cpu(batch, max(1, args.steps), max(1, args.warmup), threads)
    sample = f"oracle port (torch eager fp32, reference arithmetic), batch {batch} per step, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "dinov2-small + LoRA fine-tuning step, 24 keypoints, 224x224 (CPU, bounded sample: "
                               f"batch {batch} per step instead of 64)"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
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

    from dino_pose_b200.model import Dinov2PoseModelLoRA
    from dino_pose_b200.train import PoseTrainer
    from dino_pose_b200.synthetic import make_inputs

    B = args.batch
    torch.manual_seed(0)                      # identical random-init replica on every rank
    model = Dinov2PoseModelLoRA(num_keypoints=24, backbone=ARCH, heatmap_size=48, lora_rank=8, lora_alpha=16,
                                lora_dropout=0.1).to(dev)
    trainer = PoseTrainer(model)
    host = {k: v.pin_memory() for k, v in make_inputs(B, 224, 224, seed=rank).items()}
    devb = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def step_resident():
        return trainer.step(devb["pixel_values"], devb["heatmaps"], devb["keypoints"], devb["z"])

    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ready = [None, None]
    e2e_state = {"k": 0, "last": None}

    def step_e2e():
        """One end-to-end step through the public API: pinned HOST inputs (H2D inside PoseTrainer.step) and a D2H
        read of the step's loss.  The loss of step k is read back while step k+1 is already enqueued (one-step
        delayed logging), so the host never stalls the device; every step's loss is read inside the timed region
        (the last one by ``drain_e2e``)."""
        b = host
        k = e2e_state["k"]
        loss, _, _ = trainer.step(b["pixel_values"], b["heatmaps"], b["keypoints"], b["z"])
        slot = k & 1
        loss_host[slot].copy_(loss.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        loss_ready[slot] = ev
        prev = loss_ready[slot ^ 1]
        if prev is not None:
            prev.synchronize()
            e2e_state["last"] = float(loss_host[slot ^ 1][0])
        e2e_state["k"] = k + 1

    def drain_e2e():
        slot = (e2e_state["k"] - 1) & 1
        if loss_ready[slot] is not None:
            loss_ready[slot].synchronize()
            e2e_state["last"] = float(loss_host[slot][0])

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

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_resident, args.steps)
    for _ in range(2):
        step_e2e()
    drain_e2e()
    ms_e2e = timed(step_e2e, args.steps, after=drain_e2e)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=3)

    # per-kernel roofline: one instrumented replay of the forward and backward programs (CUDA events around
    # every launch, on the launching stream), rank 0, after the timed region
    roof = None
    launches = 0
    if rank == 0:
        st = trainer._steps[(B, 224, 224)]
        plan = st["plan"]
        progs = [plan["fwd"], st["loss"], plan["bwd"], st["opt"]]
        launches = sum(len(p) for p in progs)
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
                f.write("idx,name,kernel,ms,gflop,tflops,mbytes,gbs\n")
                for i in sorted(per_launch):
                    r = per_launch[i]
                    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
                    gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
                    f.write(f'{i},{r["name"]},{r["kernel"]},{r["ms"]:.5f},{r["flops"] / 1e9:.3f},{tf:.1f},'
                            f'{r["bytes"] / 1e6:.2f},{gb:.0f}\n')
        peaks = load_peaks()
        top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        name, a = top
        total_ms = sum(v["ms"] for v in agg.values())
        achieved_events = a["flops"] / (a["ms"] * 1e-3) / 1e12 if a["ms"] > 0 else 0.0
        # The event pair around a single eager launch also contains the launch hand-over (measured ~5 us per pair:
        # the event table sums to ~1.3 ms more than the graph-replayed step).  The figure reported as `achieved` is
        # therefore taken from a CUDA graph that holds ONLY this family's launches of one step, in step order, on the
        # buffers of the last real step, replayed back to back: GPU time / launches, timed with events on the
        # launching stream.  (Done last: the stray BatchNorm statistics it accumulates are re-zeroed below.)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        fam_graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for p in progs:
                p.run_family(name)
            torch.cuda.synchronize()
            with torch.cuda.graph(fam_graph, stream=side):
                fam_launches = sum(p.run_family(name) for p in progs)
            fam_graph.replay()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fam_reps = 5
            f0.record(side)
            for _ in range(fam_reps):
                fam_graph.replay()
            f1.record(side)
            torch.cuda.synchronize()
        fam_ms = f0.elapsed_time(f1) / fam_reps
        torch.cuda.current_stream().wait_stream(side)
        for L in plan["layers"].values():
            if "sums" in L.t:
                L.t["sums"].zero_()
        fam_flops = a["flops"] / reps
        achieved = fam_flops / (fam_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")   # ncu --set full of one step, tools/ncu_summary.py traffic
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(name, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peaks["tf_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"], "traffic": traffic,
                "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over "
                                  "the launches of this kernel family in one step)" if traffic is not None else None,
                "algorithmic_flops_per_launch": a["flops"] / a["n"],
                "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "timing": "CUDA graph of this family's launches of one step (step order, buffers of the last step), "
                          "replayed 5x, events on the launching stream",
                "share_of_step": fam_ms / ms_step, "launches_per_step": fam_launches,
                "avg_launch_ms": fam_ms / max(1, fam_launches),
                "achieved_single_launch_events": achieved_events,
                "single_launch_events_note": "event pair around each eager launch; includes ~5 us of launch hand-over per pair",
                "avg_launch_ms_events": a["ms"] / a["n"],
                "per_kernel_ms_per_step": {k: round(v["ms"] / reps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:12]},
                "step_tflops": FLOPS_PER_IMAGE_STEP * B / (ms_step * 1e-3) / 1e12,
                "step_frac_of_peak": FLOPS_PER_IMAGE_STEP * B / (ms_step * 1e-3) / 1e12 / peaks["tf_sustained"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = 16
        ips, dt = time_cpu(cb, 2, 1, threads)
        cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle port (torch eager fp32), same step at batch {cb}, 1 warm-up + 2 timed steps"}

    if rank == 0:
        total_b = B * world
        out = {
            "metric": METRIC, "value": total_b / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "dinov2-small (ViT-S/14) + LoRA fine-tuning step, 24 keypoints, 224x224, batch "
                                   f"{B} per GPU (BASELINE.json configs[1])",
                       "global_batch": total_b, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations + saved tensors > 1 GB) exceeds the 126 MB L2; "
                             "no explicit flush",
                       "step": "fwd + reference losses + bwd (heads, final LN, last-block MLP, LoRA) + AdamW, replayed as "
                               "one CUDA graph" if trainer.use_graph else "fwd + losses + bwd + AdamW (eager launches)"},
            "per_gpu": B / (ms_step * 1e-3),
            "e2e": {"value": total_b / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
        }
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
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
