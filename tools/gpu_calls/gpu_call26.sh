#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env VLY_MEGA_DBG=1 "$@" timeout 200 python tools/bench_decode.py --model $M --batch $B --steps $ST > gpurun_out/ab26_$name.log 2>&1; }
M=valley-13b; B=4; ST=120
run 13b_b4_s120_def
run 13b_b4_s120_ik48 VLY_ATTN_IKEYS=48
run 13b_b4_s120_ik32 VLY_ATTN_IKEYS=32
ST=40
run 13b_b4_s40_def
run 13b_b4_s40_ik48 VLY_ATTN_IKEYS=48
ST=250
run 13b_b4_s250_def
run 13b_b4_s250_ik48 VLY_ATTN_IKEYS=48
run 13b_b4_s250_ik64 VLY_ATTN_IKEYS=64
echo done
