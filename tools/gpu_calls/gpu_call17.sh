#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit17_new.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "vit or fulldepth or prefill or pool or projector or forward" 2>&1 | tail -12 > gpurun_out/r02_gputests_17.log
timeout 200 python tools/bench_prefill.py --model valley2-7b --batch 1 > gpurun_out/pre17_7b.log 2>&1
timeout 300 python tools/bench_prefill.py --model valley-13b --batch 4 > gpurun_out/pre17_13b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r02_vit64.csv python tools/bench_vit.py --frames 64 > gpurun_out/ncu17.log 2>&1
echo done
