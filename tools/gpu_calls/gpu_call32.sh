#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bench_vit.py --frames 8 16 32 > gpurun_out/vit32_def.log 2>&1
VLY_GEMM_BN=256 timeout 200 python tools/bench_vit.py --frames 8 16 32 > gpurun_out/vit32_bn256.log 2>&1
VLY_GEMM_BN=256 VLY_GEMM_CG2=1 timeout 200 python tools/bench_vit.py --frames 8 16 32 > gpurun_out/vit32_bn256_cg2.log 2>&1
VLY_GEMM_BN=128 timeout 200 python tools/bench_vit.py --frames 8 16 32 > gpurun_out/vit32_bn128.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre32_13b_def.log 2>&1
VLY_GEMM_BN=256 timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre32_13b_bn256.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre32_13b_def2.log 2>&1
VLY_GEMM_BN=256 timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre32_13b_bn256_2.log 2>&1
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre32_7b_def.log 2>&1
VLY_GEMM_BN=256 timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre32_7b_bn256.log 2>&1
VLY_GEMM_BN=256 VLY_GEMM_CG2=1 timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre32_7b_bn256_cg2.log 2>&1
echo done
