#!/bin/bash
# usage: N=4 bash tools/gpu_job_reserve.sh -- A/B of DP_COMM_RESERVE_SMS (SMs left to NCCL by the backward's persistent grids)
O=gpurun_out; mkdir -p $O; N=${N:-4}
run() { tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/rsv_n${N}_$tag.log 2> $O/rsv_n${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/rsv_n${N}_$tag.log") if l.startswith("{")][-1])
    print("N=$N $tag", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "replica_diff", d.get("replica_param_max_abs_diff"))
except Exception as e:
    print("N=$N $tag failed", e); print(open("$O/rsv_n${N}_$tag.err").read()[-600:])
PY
}
for v in ${VARIANTS:-0 8 16 24}; do run r$v DP_COMM_RESERVE_SMS=$v; done
if [ -n "$CTAS" ]; then run r${CTAS}c DP_COMM_RESERVE_SMS=$CTAS NCCL_MAX_CTAS=$CTAS; fi
