#!/bin/bash
# usage: N=8 bash tools/gpu_job_scaling.sh   -- bench.py under torchrun at N GPUs: default, then the A/B switches
O=gpurun_out; mkdir -p $O; N=${N:-2}
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator > $O/scale_n${N}_$tag.log 2> $O/scale_n${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/scale_n${N}_$tag.log") if l.startswith("{")][-1])
    print("N=$N $tag", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), "replica_diff", d.get("replica_param_max_abs_diff"))
except Exception as e:
    print("N=$N $tag failed", e); print(open("$O/scale_n${N}_$tag.err").read()[-800:])
PY
}
run default A=1
if [ "${AB:-0}" = "2" ]; then
  run prio0 DP_COMM_PRIORITY=0
  run reserve4 DP_RESERVE_SMS=4 NCCL_MAX_CTAS=4
  run reserve8 DP_RESERVE_SMS=8 NCCL_MAX_CTAS=8
  run reserve16 DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16
fi
if [ "${AB:-0}" = "1" ]; then
  run noallreduce DP_NO_ALLREDUCE=1
  run maxctas4 NCCL_MAX_CTAS=4
  run maxctas8 NCCL_MAX_CTAS=8
  run bucket4 DP_BUCKET_MB=4
  run bucket16 DP_BUCKET_MB=16
fi
for c in ${CFGS}; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > $O/scale_n${N}_$c.log 2> $O/scale_n${N}_$c.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/scale_n${N}_$c.log") if l.startswith("{")][-1])
    print("N=$N $c", round(d["ms_per_step"],4), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("N=$N $c failed", e); print(open("$O/scale_n${N}_$c.err").read()[-800:])
PY
done
