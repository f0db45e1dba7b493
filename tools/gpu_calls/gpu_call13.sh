#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_gputests_13.log
timeout 600 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab13_7b_b1.log 2>&1
VLY_MEGA_DBG=1 VLY_ATTN_IKEYS=32 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab13_7b_b1_ik32.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_INFLIGHT_KB=116 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab13_7b_b1_if116.log 2>&1
VLY_LIB_PATH=$PWD/valley_b200/lib/libvalley_b200_r1.so VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab13_7b_b1_old.log 2>&1
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab13_13b_b4.log 2>&1
VLY_MEGA_DBG=1 VLY_ATTN_IKEYS=16 timeout 120 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab13_13b_b4_ik16.log 2>&1
echo done
