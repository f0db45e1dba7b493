#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "rc=$?" >> gpurun_out/bench_r02_n2.err
echo done
