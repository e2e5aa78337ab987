#!/bin/bash
# bench.py on the default configuration (+ event dump) and the supporting configurations named in $CFGS
O=gpurun_out; mkdir -p $O; TAG=${TAG:-b}
DP_BENCH_DUMP=$O/${TAG}_launches_events.csv timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "exit=$?"; tail -c 6000 $O/${TAG}_bench.log; tail -n 5 $O/${TAG}_bench.err
for c in ${CFGS}; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/${TAG}_bench_$c.log 2> $O/${TAG}_bench_$c.err; echo "== $c exit=$?"; tail -c 3000 $O/${TAG}_bench_$c.log; tail -n 3 $O/${TAG}_bench_$c.err; done
