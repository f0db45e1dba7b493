#!/bin/bash
mkdir -p gpurun_out
PREV=$PWD/valley_b200/lib/libvalley_b200_prev.so
timeout 200 python tools/bench_vit.py --frames 8 16 32 64 128 256 > gpurun_out/vit29_new.log 2>&1
VLY_LIB_PATH=$PREV timeout 200 python tools/bench_vit.py --frames 8 16 32 64 128 256 > gpurun_out/vit29_prev.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre29_13b_new.log 2>&1
VLY_LIB_PATH=$PREV timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre29_13b_prev.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_gputests_29.log
echo done
