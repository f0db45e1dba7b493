#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or greedy_decode or left_padded or eos or fused_sampling or production" 2>&1 | tail -25 > gpurun_out/r02_gputests_11.log
for i in 1 2 3; do
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab11_13b_b4_umma_$i.log 2>&1; echo "rc=$?" >> gpurun_out/ab11_13b_b4_umma_$i.log
done
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley-13b --batch 2 --steps 120 > gpurun_out/ab11_13b_b2_umma.log 2>&1; echo "rc=$?" >> gpurun_out/ab11_13b_b2_umma.log
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley-13b --batch 3 --steps 120 > gpurun_out/ab11_13b_b3_umma.log 2>&1; echo "rc=$?" >> gpurun_out/ab11_13b_b3_umma.log
timeout 600 python -m pytest tests/test_gpu_fulldepth.py -q -s -k 13b 2>&1 | grep -E "rel-Fro|KV cache|passed|failed|Error" | cut -c1-600 > gpurun_out/r02_fulldepth_11.log
echo done
