#!/bin/bash
mkdir -p gpurun_out
R1=$PWD/valley_b200/lib/libvalley_b200_r1.so
run() { name=$1; shift; env VLY_MEGA_DBG=1 "$@" timeout 150 python tools/bench_decode.py --model $M --batch $B --steps 120 > gpurun_out/ab22_$name.log 2>&1; }
M=valley2-7b; B=1
run 7b_b1_thr92
run 7b_b1_thr97 VLY_MEGA_ROWS_THR=97
run 7b_b1_rows8 VLY_MEGA_ROWS=8
run 7b_b1_r1 VLY_LIB_PATH=$R1
M=valley-13b
run 13b_b1_thr92
run 13b_b1_thr97 VLY_MEGA_ROWS_THR=97
run 13b_b1_r1 VLY_LIB_PATH=$R1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "greedy_decode or left_padded or eos or fused_sampling or production or long_prompt or decode" 2>&1 | tail -5 > gpurun_out/r02_gputests_22.log
echo done
