#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
echo done
