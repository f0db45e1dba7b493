#!/bin/bash
mkdir -p gpurun_out
PREV=$PWD/valley_b200/lib/libvalley_b200_prev.so
for i in 1 2; do
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre30_13b_new_$i.log 2>&1
VLY_LIB_PATH=$PREV timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre30_13b_prev_$i.log 2>&1
done
VLY_GEMM_CG2=0 timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre30_13b_new_cg0.log 2>&1
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre30_7b_new.log 2>&1
VLY_LIB_PATH=$PREV timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre30_7b_prev.log 2>&1
echo done
