#!/bin/bash
mkdir -p gpurun_out
tools/ringbw > gpurun_out/ringbw2.log 2>&1
for v in "8 33" "8 42" "8 28" "4 21"; do
  set -- $v
  VLY_MEGA_DBG=1 VLY_MEGA_ROWS=$1 VLY_MEGA_STAGE_KB=$2 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/dbg3_13b_b4_r$1_kb$2.log 2>&1
done
VLY_MEGA_DBG=1 VLY_MEGA_ROWS=8 VLY_MEGA_STAGE_KB=33 python tools/bench_decode.py --model valley2-7b --batch 4 --steps 120 > gpurun_out/dbg3_7b_b4_r8_kb33.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=3 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/dbg3_7b_b1_st3.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=3 VLY_MEGA_STAGE_KB=41 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/dbg3_13b_b1_st3_kb41.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=4 VLY_MEGA_STAGE_KB=41 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/dbg3_13b_b1_st4_kb41.log 2>&1
python -m pytest tests/test_gpu_parity.py -q -x -k "decode or left_padded or production or eos or sampling or long_prompt" 2>&1 | tail -3 > gpurun_out/r02_gputests_3.log
echo done
