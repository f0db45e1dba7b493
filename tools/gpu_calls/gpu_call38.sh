#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
timeout 600 python bench.py --no-cpu-baseline --no-7b --vit-sweep --steps 2 > gpurun_out/bench_r02_n1_sweep.json 2> gpurun_out/bench_r02_n1_sweep.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1
echo done
