#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python tools/gemm_tune.py proj_ln,fc2_ln,proj,fc2 128 2 > $O/lab4_tune.log 2>&1; cat $O/lab4_tune.log
