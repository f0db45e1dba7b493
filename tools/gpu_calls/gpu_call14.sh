#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/umma_rate > gpurun_out/umma_rate_r02.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "tcgen05 or production or greedy_decode or left_padded or eos or long_prompt" 2>&1 | tail -8 > gpurun_out/r02_gputests_14.log
run() { name=$1; shift; env VLY_MEGA_DBG=1 "$@" timeout 150 python tools/bench_decode.py --model $M --batch $B --steps 120 > gpurun_out/ab14_$name.log 2>&1; }
M=valley2-7b; B=1
run 7b_b1_def
run 7b_b1_rows8 VLY_MEGA_ROWS=8
run 7b_b1_rows8_if128 VLY_MEGA_ROWS=8 VLY_MEGA_INFLIGHT_KB=128
run 7b_b1_if128 VLY_MEGA_INFLIGHT_KB=128
run 7b_b1_old VLY_LIB_PATH=$PWD/valley_b200/lib/libvalley_b200_r1.so
B=4
run 7b_b4_umma
run 7b_b4_hmma VLY_DECODE_UMMA=0
run 7b_b4_old VLY_LIB_PATH=$PWD/valley_b200/lib/libvalley_b200_r1.so
echo done
