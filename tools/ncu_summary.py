#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run in the CPU container on files brought back in gpurun_out/).

    python tools/ncu_summary.py full   gpurun_out/X.ncu-rep          > profiles/X_full.md
    python tools/ncu_summary.py launch gpurun_out/X_ncu_launches.csv > profiles/X_launches.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.sum", "lts__t_sector_hit_rate.pct"]


def short(name):
    name = re.sub(r"void |dp::|<unnamed>::|at::native::|at::", "", name)
    return re.sub(r"\(.*", "", name)[:80]


def _raw(path):
    """raw-page CSV of a report; `path` may already be that CSV (exported on the GPU box to stay under the size limit)"""
    if path.endswith(".csv"):
        return open(path).read()
    return subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full(path):
    raw = _raw(path)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{path}`\n")
    print("| # | kernel | grid | " + " | ".join(k for k in KEYS if k in idx) + " |")
    print("|---|---|---|" + "---|" * sum(k in idx for k in KEYS))
    for n, r in enumerate(rows[2:]):
        vals = [f"{r[idx[k]]} {units[idx[k]]}" for k in KEYS if k in idx]
        print(f"| {n} | `{short(r[idx['Kernel Name']])}` | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |")


def launch(path):
    txt = open(path).read().splitlines()
    i = [k for k, l in enumerate(txt) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list `{path}`: {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised)\n")
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")


def traffic(path):
    """JSON for bench.py's roofline.traffic: DRAM bytes (read + write) per launch, averaged per kernel family."""
    import json
    raw = _raw(path)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    fam = collections.OrderedDict()
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        key = "gemm_kmajor_tcgen05" if "gemm_fwd" in name else ("gemm_wgrad_tcgen05" if "gemm_wgrad" in name else name.split("<")[0])
        b = sum(float(r[idx[k]]) * mult.get(units[idx[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        us = float(r[idx["gpu__time_duration.sum"]]) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)
        tens = float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in idx else 0.0
        f = fam.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "tensor_pct_time_weighted": 0.0})
        f["launches"] += 1
        f["dram_bytes"] += b
        f["us"] += us
        f["tensor_pct_time_weighted"] += tens * us
    out = {}
    for k, f in fam.items():
        out[k] = {"launches": f["launches"], "dram_bytes_per_launch": f["dram_bytes"] / f["launches"],
                  "ncu_us_per_launch": f["us"] / f["launches"],
                  "tensor_pipe_active_pct": f["tensor_pct_time_weighted"] / f["us"] if f["us"] else 0.0, "source": path}
    print(json.dumps(out, indent=1))


def _family(name):
    name = short(name)
    if "gemm_fwd" in name:
        return "gemm_kmajor_tcgen05"
    if "gemm_wgrad" in name:
        return "gemm_wgrad_tcgen05"
    base = name.split("<")[0] if not name.startswith("void ") else name
    base = re.sub(r"^.*::", "", name.split("(")[0].split("<")[0])
    return re.sub(r"_kernel$", "", base)


def steptraffic(path, last=261, end=None):
    """Long-format CSV (`ncu --metrics ... --csv`, one row per launch and metric) of the whole program: keeps the LAST
    `last` launches (one steady-state fine-tuning step), prints a per-family markdown table on stderr-free stdout and
    writes the JSON bench.py reads (profiles/traffic.json) when a second argument names it."""
    import json
    txt = open(path).read()
    rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
    idx = {h: i for i, h in enumerate(rows[0])}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]]})
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", "")) * mult.get(r[idx["Metric Unit"]], 1.0)
    import os
    last = int(os.environ.get("STEP_LAST", last))
    end = int(os.environ["STEP_END"]) if "STEP_END" in os.environ else end   # capture cut short: pick a complete step
    step = list(launches.values())[:end][-int(last):]
    fam = collections.OrderedDict()
    for d in step:
        f = fam.setdefault(_family(d["name"]), {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "tens": 0.0})
        f["launches"] += 1
        f["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        f["us"] += d.get("gpu__time_duration.sum", 0.0)
        f["tens"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    tot = sum(f["us"] for f in fam.values())
    print(f"# one fine-tuning step under ncu (`{path}`, last {len(step)} launches): {tot:.0f} us serialised, cold caches\n")
    print("| kernel family | launches | us | share | DRAM MB / launch | DRAM GB/s | tensor pipe active % |")
    print("|---|---|---|---|---|---|---|")
    out = {}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| `{k}` | {f['launches']} | {f['us']:.1f} | {100 * f['us'] / tot:.1f}% | {f['dram_bytes'] / f['launches'] / 1e6:.2f} | "
              f"{f['dram_bytes'] / f['us'] / 1e3:.0f} | {f['tens'] / f['us'] if f['us'] else 0:.1f} |")
        out[k] = {"launches": f["launches"], "dram_bytes_per_launch": f["dram_bytes"] / f["launches"],
                  "ncu_us_per_launch": f["us"] / f["launches"], "tensor_pipe_active_pct": f["tens"] / f["us"] if f["us"] else 0.0,
                  "source": path}
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    {"full": full, "launch": launch, "traffic": traffic, "steptraffic": steptraffic}[sys.argv[1]](sys.argv[2])
