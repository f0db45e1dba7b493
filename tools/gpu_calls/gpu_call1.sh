#!/bin/bash
# first GPU call of round 2: full GPU suite (incl. full-depth parity), ring calibration, decode breakdowns, default bench + CPU arm
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; free -g | head -2; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/memory.max; grep -c amx_bf16 /proc/cpuinfo; } > gpurun_out/box.log 2>&1
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02_gputests_1.log
tools/ringbw > gpurun_out/ringbw.log 2>&1
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/dbg_13b_b4.log 2>&1
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/dbg_7b_b1.log 2>&1
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=6 VLY_MEGA_INFLIGHT=3 python tools/bench_decode.py --model valley2-7b --batch 1 --steps 120 > gpurun_out/dbg_7b_b1_s6.log 2>&1
VLY_MEGA_DBG=1 python tools/bench_decode.py --model valley-13b --batch 1 --steps 120 > gpurun_out/dbg_13b_b1.log 2>&1
python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_a_ref.json 2> gpurun_out/bench_r02_a_ref.err
echo done
