#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit16_tepi.log 2>&1
VLY_GEMM_TEPI=0 timeout 300 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit16_off.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "not fulldepth" 2>&1 | tail -12 > gpurun_out/r02_gputests_16.log
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre16_7b_tepi.log 2>&1
VLY_GEMM_TEPI=0 timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre16_7b_off.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre16_13b_tepi.log 2>&1
VLY_GEMM_TEPI=0 timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre16_13b_off.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fulldepth.py -q -x 2>&1 | tail -12 > gpurun_out/r02_gputests_16b.log
echo done
