"""2+ GPU check (torchrun): fused ViT+gather (peer stores from the GEMM epilogue) == ViT + NCCL all_gather, bit for bit."""
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
for n_frames in (2 * world, 2 * world + 1, 5):
    px = syn.make_pixels(1, n_frames, 3)[0]
    lo, hi = vdist.shard_bounds(n_frames, world, rank)
    ref = vdist.encode_frames_sharded(m.encode_frames, px[lo:hi].cuda(), n_frames)
    fg = vdist.FusedFrameGather.__new__(vdist.FusedFrameGather) if False else None
    if n_frames == 2 * world:
        fused = vdist.FusedFrameGather(m, 2 * world + 1)      # buffer large enough for the biggest case
    # reuse the same buffer for every case (rows beyond n_frames are stale and ignored)
    fused.n_frames_total = n_frames
    got = fused.encode(px[lo:hi].cuda())[:n_frames]
    torch.cuda.synchronize()
    same = torch.equal(got, ref)
    fused.release()
    ok &= same
    print(f"rank {rank}: n_frames={n_frames} fused == nccl: {same}  max|d|={(got.float() - ref.float()).abs().max().item():.3e}", flush=True)
dist.barrier()
if rank == 0:
    print("FUSED GATHER", "OK" if ok else "MISMATCH")
dist.destroy_process_group()
