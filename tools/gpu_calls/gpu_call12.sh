#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or greedy_decode or left_padded or eos or fused_sampling or production or long_prompt" 2>&1 | tail -25 > gpurun_out/r02_gputests_12.log
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab12_13b_b4_umma.log 2>&1; echo "rc=$?" >> gpurun_out/ab12_13b_b4_umma.log
VLY_MEGA_DBG=1 VLY_ATTN_CTA=0 timeout 120 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab12_13b_b4_umma_dist.log 2>&1; echo "rc=$?" >> gpurun_out/ab12_13b_b4_umma_dist.log
VLY_MEGA_DBG=1 VLY_ATTN_CTA=1 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab12_7b_b1_cta.log 2>&1; echo "rc=$?" >> gpurun_out/ab12_7b_b1_cta.log
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/ab12_7b_b1.log 2>&1; echo "rc=$?" >> gpurun_out/ab12_7b_b1.log
VLY_MEGA_DBG=1 timeout 120 python tools/bench_decode.py --model valley2-7b --batch 4 --steps 120 > gpurun_out/ab12_7b_b4.log 2>&1; echo "rc=$?" >> gpurun_out/ab12_7b_b4.log
echo done
