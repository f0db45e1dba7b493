"""Same request twice (and through two caches): greedy ids must be identical.  python tools/check_decode_determinism.py [--model valley-13b --batch 4 --new 256]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from valley_b200 import synthetic as syn
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="valley-13b"); ap.add_argument("--batch", type=int, default=4); ap.add_argument("--new", type=int, default=256); ap.add_argument("--frames", type=int, default=16)
a = ap.parse_args()
spec = syn.SPECS[a.model]
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), 0)
m.load_state_dict(syn.iter_state_dict(spec, 0, device="cuda:0"))
ids = syn.make_prompt_ids(spec, a.batch, a.frames, 0).cuda()
px = syn.make_pixels(a.batch, a.frames, 0, dtype=torch.float16).cuda()
S = ids.shape[1]
outs = [m.generate(input_ids=ids, images=px, max_new_tokens=a.new)[:, S:].clone() for _ in range(4)]
for i in range(1, 4):
    eq = torch.equal(outs[0], outs[i])
    first = None
    if not eq:
        d = (outs[0] != outs[i]).nonzero()
        first = d[0].tolist()
    print(f"run 0 vs run {i}: equal={eq} first difference (row, step)={first}")
