"""ViT features must not depend on which other frames share the launch: encode 64 frames at once vs 4 x 16 and 8 x 8, bitwise.
   python tools/check_batch_invariance.py [out.pt]   (saves the 64-frame result for cross-process comparisons)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from valley_b200 import synthetic as syn
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM
spec = syn.VALLEY2_7B
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), 0)
m.load_state_dict(syn.iter_state_dict(spec, 0, device="cuda:0", llm=False))
px = syn.make_pixels(1, 64, 7, dtype=torch.float16)[0].cuda()
full = m.encode_frames(px).clone()
for chunk in (16, 8, 32):
    parts = torch.cat([m.encode_frames(px[i:i + chunk]).clone() for i in range(0, 64, chunk)])
    d = (parts.float() - full.float()).abs()
    print(f"64 at once vs {64 // chunk} x {chunk}: equal={torch.equal(parts, full)}  max|d|={d.max().item():.3e}  differing elements={int((d > 0).sum())}")
again = m.encode_frames(px)
print("run-to-run equal:", torch.equal(again, full), " env TEPI=", os.environ.get("VLY_GEMM_TEPI"), " CG2=", os.environ.get("VLY_GEMM_CG2"))
if len(sys.argv) > 1:
    torch.save(full.cpu(), sys.argv[1])
