#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit18_new.log 2>&1
VLY_LIB_PATH=$PWD/valley_b200/lib/libvalley_b200_prev.so timeout 300 python tools/bench_vit.py --frames 8 32 64 256 > gpurun_out/vit18_prev.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "vit or fulldepth or pool or forward or gather" 2>&1 | tail -12 > gpurun_out/r02_gputests_18.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|vit_attention|vit_embed|im2col" -c 600 --csv --log-file gpurun_out/launches_r02_vit64.csv python tools/bench_vit.py --frames 64 > gpurun_out/ncu18.log 2>&1
echo done
