#!/bin/bash
# Round-2 ncu evidence (reports are converted to CSV ON THE BOX: gpurun_out/ may carry at most 64 MiB back).
O=gpurun_out; mkdir -p $O; T=r2
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-comparator"
timeout 300 $BENCH > $O/${T}_plain.log 2>&1 || { tail -n 5 $O/${T}_plain.log; exit 1; }
# 1. launch list of the benchmark command itself
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/${T}_ncu_launches_bench.csv $BENCH > $O/${T}_ncu0.log 2>&1
# 2. --set full of one backbone layer of a steady-state step (LN, qkv, attention, proj, LN, fc1, fc2, ...)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_fwd_kernel|layernorm_fwd|attention_tc257' -s 260 -c 9 -o /tmp/${T}_fulla $BENCH > $O/${T}_ncu1.log 2>&1
ncu -i /tmp/${T}_fulla.ncu-rep --page raw --csv > $O/${T}_full_backbone_raw.csv 2>/dev/null
# 3. the LoRA-fused projection, the decode kernel, one BatchNorm-backward pair
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_rowln|decode_kernel' -s 3 -c 2 -o /tmp/${T}_fullb $BENCH > $O/${T}_ncu2.log 2>&1
ncu -i /tmp/${T}_fullb.ncu-rep --page raw --csv > $O/${T}_full_lora_decode_raw.csv 2>/dev/null
# 4. flash attention forward (T = 1025) and the attention backward kernels (T = 257)
B=16 T=1025 timeout 120 python tools/attn_tune.py > $O/${T}_attn_plain.log 2>&1 && \
B=16 T=1025 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_flash' -s 1 -c 1 -o /tmp/${T}_fullc python tools/attn_tune.py > $O/${T}_ncu3.log 2>&1
ncu -i /tmp/${T}_fullc.ncu-rep --page raw --csv > $O/${T}_full_attn_flash_raw.csv 2>/dev/null
timeout 120 python tools/attn_bwd_tune.py > $O/${T}_attnbwd_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_bwd' -s 3 -c 3 -o /tmp/${T}_fulld python tools/attn_bwd_tune.py > $O/${T}_ncu4.log 2>&1
ncu -i /tmp/${T}_fulld.ncu-rep --page raw --csv > $O/${T}_full_attnbwd_raw.csv 2>/dev/null
ls -la $O/${T}_* | head -20
tail -n 2 $O/${T}_ncu0.log $O/${T}_ncu1.log $O/${T}_ncu2.log $O/${T}_ncu3.log $O/${T}_ncu4.log | cut -c1-200
