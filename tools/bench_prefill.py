"""Prefill timing (ViT + pool/projector + splice + LLaMA prefill of the whole prompt): python tools/bench_prefill.py --model valley-13b --batch 4"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from valley_b200 import synthetic as syn
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="valley2-7b")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--frames", type=int, default=8)
a = ap.parse_args()
spec = syn.SPECS[a.model]
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), 0)
m.load_state_dict(syn.iter_state_dict(spec, 0, device="cuda:0"))
m.logits_all_positions = False
ids = syn.make_prompt_ids(spec, a.batch, a.frames, 0).cuda()
px = syn.make_pixels(a.batch, a.frames, 0, dtype=torch.float16).cuda()
for _ in range(3):
    out = m(input_ids=ids, images=px)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = m(input_ids=ids, images=px); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"{a.model} B={a.batch} S={ids.shape[1]}: ViT + projector + prefill {best:.3f} ms   argmax[0]={int(out.logits[0, -1].argmax())} lib={os.environ.get('VLY_LIB_PATH', 'default')}")
