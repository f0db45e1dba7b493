#!/bin/bash
mkdir -p gpurun_out
PREV=$PWD/valley_b200/lib/libvalley_b200_prev.so
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre34_7b_new.log 2>&1
VLY_LIB_PATH=$PREV timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre34_7b_prev.log 2>&1
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre34_7b_new2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|prefill_attention|gemv|splice|pool" -c 400 --csv --log-file gpurun_out/launches_r02_prefill7b.csv python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/ncu34.log 2>&1
echo done
