#!/bin/bash
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed"
# 1) launch list of exactly one timed bench step (13B, B=4, 8 frames, 256 tokens)
VLY_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_bench_step.csv python bench.py --no-cpu-baseline --no-7b --steps 1 --warmup 3 > gpurun_out/ncu20_a.log 2>&1
# 2) full captures of the decode-step kernels
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_step --launch-skip 2 -c 1 -o gpurun_out/prof_mega_r02_13b_b4 -f python tools/profile_target.py --model valley-13b --batch 4 --skip-vit --decode 4 > gpurun_out/ncu20_b.log 2>&1
ncu -i gpurun_out/prof_mega_r02_13b_b4.ncu-rep --page raw --csv --metrics $M > gpurun_out/prof_mega_r02_13b_b4_summary.csv 2>> gpurun_out/ncu20_b.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:decode_step --launch-skip 2 -c 1 -o gpurun_out/prof_mega_r02_7b_b1 -f python tools/profile_target.py --model valley2-7b --batch 1 --skip-vit --decode 4 > gpurun_out/ncu20_c.log 2>&1
ncu -i gpurun_out/prof_mega_r02_7b_b1.ncu-rep --page raw --csv --metrics $M > gpurun_out/prof_mega_r02_7b_b1_summary.csv 2>> gpurun_out/ncu20_c.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:decode_step --launch-skip 2 -c 1 -o gpurun_out/prof_mega_r02_13b_b1 -f python tools/profile_target.py --model valley-13b --batch 1 --skip-vit --decode 4 > gpurun_out/ncu20_d.log 2>&1
ncu -i gpurun_out/prof_mega_r02_13b_b1.ncu-rep --page raw --csv --metrics $M > gpurun_out/prof_mega_r02_13b_b1_summary.csv 2>> gpurun_out/ncu20_d.log
# 3) ViT kernels, 64 frames: one mid-layer instance of each
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_tc|vit_attention_pp" --launch-skip 52 -c 5 -o gpurun_out/prof_vit_r02 -f python tools/profile_target.py --frames 64 --skip-llm > gpurun_out/ncu20_e.log 2>&1
ncu -i gpurun_out/prof_vit_r02.ncu-rep --page raw --csv --metrics $M > gpurun_out/prof_vit_r02_summary.csv 2>> gpurun_out/ncu20_e.log
ls -la gpurun_out/*.ncu-rep >> gpurun_out/ncu20_e.log
echo done
