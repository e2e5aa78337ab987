#!/bin/bash
# Final round-2 evidence for the code as committed: launch list of the benchmark command and a --set full capture of decode.
O=gpurun_out; mkdir -p $O; T=r2z
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-comparator"
timeout 300 $BENCH > $O/${T}_plain.log 2>&1 || { tail -n 5 $O/${T}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/${T}_ncu_launches_bench.csv $BENCH > $O/${T}_ncu0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_kernel' -s 2 -c 1 -o /tmp/${T}_dec $BENCH > $O/${T}_ncu1.log 2>&1
ncu -i /tmp/${T}_dec.ncu-rep --page raw --csv > $O/${T}_full_decode_raw.csv 2>/dev/null
ls -la $O/${T}_* | head; tail -n 2 $O/${T}_plain.log $O/${T}_ncu0.log $O/${T}_ncu1.log | cut -c1-300
