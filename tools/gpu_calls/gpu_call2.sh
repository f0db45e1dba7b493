#!/bin/bash
# new decode kernel: correctness first, then breakdowns
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -15 > gpurun_out/r02_gputests_2.log
python -m pytest tests/test_gpu_fulldepth.py -q -s 2>&1 | grep -E "rel-Fro|passed|failed|Error|error" | cut -c1-700 > gpurun_out/r02_fulldepth_2.log
for cfg in "valley-13b 4" "valley2-7b 1" "valley-13b 1" "valley2-7b 4" "valley-13b 2"; do
  set -- $cfg
  VLY_MEGA_DBG=1 python tools/bench_decode.py --model $1 --batch $2 --steps 120 > gpurun_out/dbg2_$1_b$2.log 2>&1
done
VLY_MEGA_DBG=1 VLY_MEGA_STAGE_KB=33 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/dbg2_13b_b4_kb33.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_ROWS=8 VLY_MEGA_STAGE_KB=42 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/dbg2_13b_b4_r8kb42.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_INFLIGHT=3 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/dbg2_7b_b1_if3.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=4 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/dbg2_7b_b1_st4.log 2>&1
echo done
