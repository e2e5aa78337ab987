#!/bin/bash
# full GPU suite + smoke + default bench (+ event dump)
O=gpurun_out; mkdir -p $O; TAG=${TAG:-f}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider > $O/${TAG}_tests.log 2>&1; tail -n 4 $O/${TAG}_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -n 2 $O/${TAG}_smoke.log | cut -c1-300
DP_BENCH_DUMP=$O/${TAG}_launches_events.csv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/${TAG}_bench.log 2>$O/${TAG}_bench.err; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench.log").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), d["roofline"]["per_kernel_ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("$O/${TAG}_bench.err").read()[-1500:])
PY
