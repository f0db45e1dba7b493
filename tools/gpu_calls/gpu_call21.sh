#!/bin/bash
mkdir -p gpurun_out
PREV=$PWD/valley_b200/lib/libvalley_b200_prev.so
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_gputests_21.log
VLY_MEGA_DBG=1 timeout 150 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab21_13b_b4_new.log 2>&1
VLY_LIB_PATH=$PREV VLY_MEGA_DBG=1 timeout 150 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab21_13b_b4_prev.log 2>&1
VLY_MEGA_DBG=1 timeout 150 python tools/bench_decode.py --model valley-13b --batch 2 --steps 120 > gpurun_out/ab21_13b_b2_new.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre21_13b.log 2>&1
VLY_GEMM_CG2=1 timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre21_13b_cg2.log 2>&1
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre21_7b.log 2>&1
VLY_GEMM_CG2=1 timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre21_7b_cg2.log 2>&1
echo done
