#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/check_decode_determinism.py > gpurun_out/det36_13b_b4.log 2>&1
VLY_ATTN_IKEYS=32 timeout 400 python tools/check_decode_determinism.py > gpurun_out/det36_13b_b4_ik32.log 2>&1
echo done
