"""Short, deterministic target for ncu: one ViT encode (F frames), one prefill, a few decode steps on Valley2-7b.
   python tools/profile_target.py [--frames 64] [--decode 3] [--model valley2-7b] [--skip-vit] [--skip-llm]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from valley_b200 import synthetic as syn  # noqa: E402
from valley_b200._lib import check  # noqa: E402
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--decode", type=int, default=3)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--model", default="valley2-7b")
ap.add_argument("--skip-vit", action="store_true")
ap.add_argument("--skip-llm", action="store_true")
a = ap.parse_args()
spec = syn.SPECS[a.model]
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), 0)
m.load_state_dict(syn.iter_state_dict(spec, 0, device="cuda:0", vision=not a.skip_vit, llm=not a.skip_llm))
for k, v in syn.sentinel_ids(spec).items():
    setattr(m.get_model().vision_tower.config, k, v)
torch.cuda.synchronize()
torch.cuda.profiler.start()      # ncu --profile-from-start off: skip weight generation / packing
if not a.skip_vit:
    px = syn.make_pixels(1, a.frames, 1, dtype=torch.float16)[0].cuda()
    torch.cuda.nvtx.range_push("vit")
    m.encode_frames(px)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
if not a.skip_llm:
    ids = syn.make_prompt_ids(spec, a.batch, 8, 0).cuda()
    cache = m.new_cache(a.batch)
    _, _, _, emb, _ = m.prepare_inputs_labels_for_multimodal(ids, None, None, None, None)
    _, nxt = m._prefill(cache, emb, 0)
    torch.cuda.synchronize()
    out = torch.empty(a.batch, a.decode, dtype=torch.int64, device="cuda")
    check(m._lib.vly_generate_greedy(m._ctx, cache._h, nxt.data_ptr(), a.decode, out.data_ptr(), 0))
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", m.launches())
