#!/bin/bash
# Round-1 closing evidence (reports converted to CSV on the box; gpurun_out/ carries at most 64 MiB back):
#   1. bench.py exits 0 without ncu, then the ncu launch list of the SAME command (gpu__time_duration only)
#   2. DRAM traffic / tensor-pipe activity of every launch of two steps (metric subset) -> profiles/traffic.json
#      (EXPENSIVE: ~15 GPU-minutes for ~1100 launches x 8 metrics; the 900 s limit cut the r1h capture short)
#   3. --set full of: the T=257 attention kernel, six consecutive backbone GEMMs, the gelu'-epilogue GEMM (fc2 input
#      gradient), the LoRA accumulate kernel, and the attention-backward kernels of the un-frozen-layer path (f4)
TAG=${1:-r1h}; O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 python bench.py --steps 2 --warmup 3 > $O/${TAG}_bench_plain.log 2>&1 || { tail -n 5 $O/${TAG}_bench_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/${TAG}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu0.log 2>&1
timeout 300 python tools/prof_step.py --steps 2 > $O/${TAG}_plain.log 2>&1 || { tail -n 5 $O/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_step_metrics.csv python tools/prof_step.py --steps 2 > $O/${TAG}_ncu1.log 2>&1
full() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 sk=$3 cn=$4; shift 4
  timeout 600 ncu --set full --clock-control none $NCU_BASE -k regex:"$rx" -s $sk -c $cn -o /tmp/${TAG}_$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i /tmp/${TAG}_$name.ncu-rep --page raw --csv > $O/${TAG}_full_$name.csv 2>/dev/null
}
full attn 'attention_tc257' 30 1 python tools/prof_step.py --steps 2
full gemms 'gemm_fwd_kernel' 200 6 python tools/prof_step.py --steps 2
# (template arguments are only part of the matched name with --kernel-name-base demangled; the plain regex above matched nothing in r1h)
NCU_BASE="--kernel-name-base demangled" full fc2dg 'gemm_fwd_kernel<128, 0, 0, 0, 16>' 2 1 python tools/prof_step.py --steps 2
full lora 'lora_bwd' 4 2 python tools/prof_step.py --steps 2
full attnbwd 'attention_bwd' 8 2 python tools/run_configs.py f4
ls -la $O | grep ${TAG}
tail -n 2 $O/${TAG}_ncu0.log $O/${TAG}_ncu1.log | cut -c1-200
