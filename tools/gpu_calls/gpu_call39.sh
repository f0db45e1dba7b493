#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "multi_pass or production_shapes" 2>&1 | tail -8 > gpurun_out/r02_gputests_39.log
echo done
