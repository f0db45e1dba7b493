#!/bin/bash
mkdir -p gpurun_out
R1=$PWD/valley_b200/lib/libvalley_b200_r1.so
run() { name=$1; shift; "$@" > gpurun_out/ab5_$name.log 2>&1; }
for cfg in "valley2-7b 1" "valley-13b 4" "valley-13b 1" "valley2-7b 4"; do
  set -- $cfg
  VLY_LIB_PATH=$R1 VLY_MEGA_DBG=1 python tools/bench_decode.py --model $1 --batch $2 --steps 120 > gpurun_out/ab5_$1_b$2_old.log 2>&1
  VLY_MEGA_DBG=1 python tools/bench_decode.py --model $1 --batch $2 --steps 120 > gpurun_out/ab5_$1_b$2_new.log 2>&1
done
VLY_MEGA_DBG=1 VLY_MEGA_STAGES=3 python tools/bench_decode.py --model valley-13b --batch 4 --steps 120 > gpurun_out/ab5_valley-13b_b4_new_st3.log 2>&1
python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3 > gpurun_out/r02_gputests_5.log
# ViT with / without programmatic dependent launch
python - > gpurun_out/vit_pdl.log 2>&1 <<'PY'
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from valley_b200 import synthetic as syn
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM
spec = syn.VALLEY2_7B
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), 0)
m.load_state_dict(syn.iter_state_dict(spec, 0, device="cuda:0", llm=False))
for F in (8, 32, 64, 256):
    px = syn.make_pixels(1, F, 1, dtype=torch.float16)[0].cuda()
    for _ in range(3): m.encode_frames(px)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m.encode_frames(px)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"F={F}: {ms:.3f} ms  {F / ms * 1e3:.0f} frames/s  NO_PDL={os.environ.get('VLY_NO_PDL')}")
PY
VLY_NO_PDL=1 python - > gpurun_out/vit_nopdl.log 2>&1 <<'PY'
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from valley_b200 import synthetic as syn
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM
spec = syn.VALLEY2_7B
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), 0)
m.load_state_dict(syn.iter_state_dict(spec, 0, device="cuda:0", llm=False))
for F in (8, 32, 64, 256):
    px = syn.make_pixels(1, F, 1, dtype=torch.float16)[0].cuda()
    for _ in range(3): m.encode_frames(px)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m.encode_frames(px)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"F={F}: {ms:.3f} ms  {F / ms * 1e3:.0f} frames/s  NO_PDL={os.environ.get('VLY_NO_PDL')}")
PY
echo done
