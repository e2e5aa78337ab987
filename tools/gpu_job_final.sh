#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke, the reference arm, the default benchmark line
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/final_tests.log 2>&1; tail -n 2 $O/final_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; tail -n 1 $O/final_smoke.log | cut -c1-250
( time timeout 900 python bench.py --impl reference ) > $O/final_ref.log 2> $O/final_ref.err; tail -n 1 $O/final_ref.log | cut -c1-400; grep real $O/final_ref.err
( time timeout 900 python bench.py ) > $O/final_bench.log 2> $O/final_bench.err; grep real $O/final_bench.err
python - <<PY
import json
d=json.loads(open("$O/final_bench.log").read().strip().splitlines()[-1])
keys=("value","unit","ms_per_step","steps","warmup","dtype","vs_baseline","gpu_launches","clocks")
print({k:d.get(k) for k in keys}); print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"]); print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","frac","traffic")}); print("gpu_comparator", d.get("gpu_comparator")); print([ (e["kernel"], round(e["frac"],3)) for e in d.get("extra_rooflines",[])])
PY
