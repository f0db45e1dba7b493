#!/bin/bash
mkdir -p gpurun_out
VLY_ATTN_DBG=1 timeout 200 python tools/bench_vit.py --frames 64 > gpurun_out/vit31_dbg.log 2>&1
echo done
