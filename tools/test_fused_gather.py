"""2+ GPU check (torchrun): fused ViT+gather (peer stores from the GEMM epilogue) == ViT + NCCL all_gather, bit for bit,
for contiguous and round-robin ("interleaved") frame dealing, even and ragged shards; and the timeout flag stays clear."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from valley_b200 import dist as vdist, synthetic as syn
from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
spec = syn.TINY
m = ValleyLlamaForCausalLM(ValleyConfig.from_spec(spec), local)
m.load_state_dict(syn.iter_state_dict(spec, 0, device=f"cuda:{local}", llm=False))
ok = True
fused = vdist.FusedFrameGather(m, 2 * world + 1)          # one buffer, large enough for the biggest case; rows beyond n_frames are stale
for n_frames in (2 * world, 2 * world + 1, 5):
    px = syn.make_pixels(1, n_frames, 3)[0]
    for inter in (False, True):
        if inter:
            mine = px[list(vdist.dealt_frames(n_frames, world, rank))]
        else:
            lo, hi = vdist.shard_bounds(n_frames, world, rank)
            mine = px[lo:hi]
        ref = vdist.encode_frames_sharded(m.encode_frames, mine.cuda(), n_frames, interleaved=inter)
        fused.n_frames_total = n_frames
        got = fused.encode(mine.cuda(), inter)[:n_frames]
        torch.cuda.synchronize()
        same = torch.equal(got, ref)
        fused.release()
        fused.check()
        ok &= same
        print(f"rank {rank}: n_frames={n_frames} interleaved={inter} fused == nccl: {same}  max|d|={(got.float() - ref.float()).abs().max().item():.3e}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("FUSED GATHER", "OK" if int(t.item()) == 1 else "MISMATCH")
dist.destroy_process_group()
