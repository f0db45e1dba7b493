#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n8_final.json 2> gpurun_out/bench_r02_n8_final.err
echo "rc=$?" >> gpurun_out/bench_r02_n8_final.err
echo done
