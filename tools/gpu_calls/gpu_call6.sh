#!/bin/bash
mkdir -p gpurun_out
tools/umma_probe > gpurun_out/umma_probe.log 2>&1
R1=$PWD/valley_b200/lib/libvalley_b200_r1.so
VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab6_7b_b1_old.log 2>&1
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab6_7b_b1_new.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=3 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab6_7b_b1_new_st3.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_INFLIGHT_KB=80 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab6_7b_b1_new_if80.log 2>&1
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/ab6_13b_b1_new.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_INFLIGHT_KB=80 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/ab6_13b_b1_new_if80.log 2>&1
VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab6_13b_b4_old.log 2>&1
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab6_13b_b4_new.log 2>&1
echo done
