#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or greedy_decode or left_padded or eos or fused_sampling or production" 2>&1 | tail -25 > gpurun_out/r02_gputests_10.log
R1=$PWD/valley_b200/lib/libvalley_b200_r1.so
VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 timeout 300 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab10_13b_b4_old.log 2>&1
VLY_MEGA_DBG=1 timeout 300 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab10_13b_b4_umma.log 2>&1
VLY_MEGA_DBG=1 VLY_UMMA_XOOB=0 timeout 300 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab10_13b_b4_umma_xfull.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_INFLIGHT_KB=60 timeout 300 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab10_13b_b4_umma_if2.log 2>&1
echo done
