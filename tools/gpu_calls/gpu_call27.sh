#!/bin/bash
mkdir -p gpurun_out
PREV=$PWD/valley_b200/lib/libvalley_b200_prev.so
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "decode or left_padded or eos or fused_sampling or production or long_prompt or tcgen05 or mask" 2>&1 | tail -5 > gpurun_out/r02_gputests_27.log
run() { name=$1; shift; env VLY_MEGA_DBG=1 "$@" timeout 200 python tools/bench_decode.py --model $M --batch $B --steps $ST > gpurun_out/ab27_$name.log 2>&1; }
M=valley-13b; B=4; ST=250
run 13b_b4_s250_new
run 13b_b4_s250_prev VLY_LIB_PATH=$PREV
B=2
run 13b_b2_s250_new
run 13b_b2_s250_prev VLY_LIB_PATH=$PREV
M=valley2-7b; B=1; ST=250
run 7b_b1_s250_new
run 7b_b1_s250_prev VLY_LIB_PATH=$PREV
echo done
