#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/n8_gpus.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err
echo "rc=$?" >> gpurun_out/bench_r02_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 2 --warmup 3 --model valley2-7b --batch 1 --new-tokens 32 --no-cpu-baseline --no-7b --vit-sweep > gpurun_out/bench_r02_n8_sweep.json 2> gpurun_out/bench_r02_n8_sweep.err
echo "rc=$?" >> gpurun_out/bench_r02_n8_sweep.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/test_fused_gather.py > gpurun_out/n8_fused_gather.log 2>&1
echo "rc=$?" >> gpurun_out/n8_fused_gather.log
echo done
