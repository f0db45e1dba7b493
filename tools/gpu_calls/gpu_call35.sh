#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/check_batch_invariance.py /tmp/f_def.pt > gpurun_out/inv35_def.log 2>&1
VLY_GEMM_TEPI=0 timeout 200 python tools/check_batch_invariance.py /tmp/f_t0.pt > gpurun_out/inv35_tepi0.log 2>&1
VLY_GEMM_TEPI=1 VLY_GEMM_CG2=0 timeout 200 python tools/check_batch_invariance.py /tmp/f_c0.pt > gpurun_out/inv35_cg0.log 2>&1
python - > gpurun_out/inv35_cmp.log 2>&1 <<'PY'
import torch
a, b, c = torch.load("/tmp/f_def.pt"), torch.load("/tmp/f_t0.pt"), torch.load("/tmp/f_c0.pt")
print("default vs TEPI=0:", torch.equal(a, b), (a.float() - b.float()).abs().max().item())
print("default vs CG2=0 :", torch.equal(a, c), (a.float() - c.float()).abs().max().item())
PY
echo done
