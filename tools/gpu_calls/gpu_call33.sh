#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_gputests_33.log
timeout 200 python tools/bench_vit.py --frames 8 16 32 64 128 > gpurun_out/vit33.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre33_13b.log 2>&1
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre33_7b.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 1 > gpurun_out/pre33_13b_b1.log 2>&1
echo done
