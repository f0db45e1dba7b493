#!/bin/bash
mkdir -p gpurun_out
tools/ringbw quick > gpurun_out/ringbw3.log 2>&1
R1=$PWD/valley_b200/lib/libvalley_b200_r1.so
for rep in 1 2; do
  VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab4_7b_b1_old_$rep.log 2>&1
  VLY_MEGA_DBG=1 VLY_MEGA_STAGES=3 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab4_7b_b1_new_st3_$rep.log 2>&1
  VLY_MEGA_DBG=1 VLY_MEGA_STAGES=4 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab4_7b_b1_new_st4_$rep.log 2>&1
done
VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab4_13b_b4_old.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_ROWS=8 VLY_MEGA_STAGE_KB=34 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab4_13b_b4_new_r8kb34.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_ROWS=8 VLY_MEGA_STAGE_KB=42 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab4_13b_b4_new_r8kb42.log 2>&1
VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/ab4_13b_b1_old.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=4 VLY_MEGA_STAGE_KB=41 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/ab4_13b_b1_new.log 2>&1
echo done
