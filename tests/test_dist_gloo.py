"""CPU, world_size 2, gloo: the N>1 host path -- frame sharding, padded all-gather, per-rank video ownership."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from valley_b200 import dist as vdist


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 64, 513):
        for w in (1, 2, 3, 4, 8):
            b = [vdist.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == vdist.all_shard_sizes(n, w)
            dealt = [list(vdist.dealt_frames(n, w, r)) for r in range(w)]
            assert sorted(sum(dealt, [])) == list(range(n)) and [len(d) for d in dealt] == vdist.dealt_sizes(n, w)
            assert max(map(len, dealt)) - min(map(len, dealt)) <= 1


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_frames * 3 * 2 * 2, dtype=torch.float32).view(n_frames, 3, 2, 2)
        lo, hi = vdist.shard_bounds(n_frames, world, rank)
        fake_encode = lambda px: px.flatten(1)[:, :6].reshape(-1, 2, 3) * 2.0       # stands in for the ViT ([F,tokens,D])
        got = vdist.encode_frames_sharded(fake_encode, full[lo:hi], n_frames)
        want = fake_encode(full)
        ok = torch.equal(got, want)
        mine = list(vdist.dealt_frames(n_frames, world, rank))              # round-robin dealing: every rank needs remote frames
        got_i = vdist.encode_frames_sharded(fake_encode, full[mine], n_frames, interleaved=True)
        ok = ok and torch.equal(got_i, want)
        vl, vh = vdist.my_videos(5)
        q.put((rank, ok, (vl, vh)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_encode_allgather_world2():
    for n_frames in (8, 5):           # even and ragged shards
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=120) for _ in ps)
        [p.join(timeout=60) for p in ps]
        assert all(r[1] for r in res), res
        assert res[0][2] == (0, 3) and res[1][2] == (3, 5)
