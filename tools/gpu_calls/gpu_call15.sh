#!/bin/bash
mkdir -p gpurun_out
HEAD=$PWD/valley_b200/lib/libvalley_b200_head.so
timeout 200 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit15_new.log 2>&1
VLY_LIB_PATH=$HEAD timeout 200 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit15_head.log 2>&1
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre15_7b_new.log 2>&1
VLY_LIB_PATH=$HEAD timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre15_7b_head.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre15_13b_new.log 2>&1
VLY_LIB_PATH=$HEAD timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre15_13b_head.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_gputests_15.log
echo done
