#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/umma_probe > gpurun_out/umma_probe2.log 2>&1
L=$PWD/valley_b200/lib
for v in r1 NO_PHASE_PREFETCH PRODUCER_BYREF ATTN_KEYS_FIXED32; do
  VLY_LIB_PATH=$L/libvalley_b200_$v.so VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab7_7b_b1_$v.log 2>&1
done
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab7_7b_b1_cur.log 2>&1
echo done
