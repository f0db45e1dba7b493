#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bench_vit.py --frames 8 16 32 64 > gpurun_out/vit28_def.log 2>&1
VLY_GEMM_CG2=1 timeout 200 python tools/bench_vit.py --frames 8 16 32 64 > gpurun_out/vit28_cg2.log 2>&1
VLY_GEMM_TEPI=2 timeout 200 python tools/bench_vit.py --frames 8 16 32 64 > gpurun_out/vit28_tepi2.log 2>&1
VLY_GEMM_CG2=0 timeout 200 python tools/bench_vit.py --frames 8 16 32 64 > gpurun_out/vit28_cg0.log 2>&1
VLY_NO_PDL=1 timeout 200 python tools/bench_vit.py --frames 8 16 > gpurun_out/vit28_nopdl.log 2>&1
echo done
