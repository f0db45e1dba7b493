#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/test_fused_gather.py > gpurun_out/n2_fused_gather.log 2>&1
echo "rc=$?" >> gpurun_out/n2_fused_gather.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "rc=$?" >> gpurun_out/bench_r02_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --model valley2-7b --batch 1 --new-tokens 32 --no-cpu-baseline --no-7b --vit-sweep > gpurun_out/bench_r02_n2_sweep.json 2> gpurun_out/bench_r02_n2_sweep.err
echo "rc=$?" >> gpurun_out/bench_r02_n2_sweep.err
echo done
