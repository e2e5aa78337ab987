#!/bin/bash
# every BASELINE configuration at full size on one GPU through bench.py (JSON lines -> profiles/r2_configs_1gpu.jsonl)
O=gpurun_out; mkdir -p $O; : > $O/r2_configs_1gpu.jsonl
for c in infer_s_b1 train_s train_b infer_l infer_s448 train_s448; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > $O/cfg_$c.log 2> $O/cfg_$c.err; tail -n 1 $O/cfg_$c.log >> $O/r2_configs_1gpu.jsonl
  python - <<PY
import json
try:
    d=json.loads(open("$O/cfg_$c.log").read().strip().splitlines()[-1])
    g=d.get("gpu_comparator") or {}
    print("$c", round(d["ms_per_step"],3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "step TF/s", round(d["roofline"]["step_tflops"]), "cpu", round((d.get("cpu_baseline") or {}).get("value",0),1), "torch fp32", round((g.get("torch_eager_fp32") or {}).get("value",0)), "bf16", round((g.get("torch_autocast_bf16") or {}).get("value",0)))
except Exception as e:
    print("$c failed", e); print(open("$O/cfg_$c.err").read()[-600:])
PY
done
