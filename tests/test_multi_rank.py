"""world_size-2 gloo test (CPU) of the N>1 path: clip sharding + the single id-map all_gather.
The kernels themselves are exercised by the -m gpu tests; here only the host-side logic runs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slotvps_b200.parallel import WIRE_DTYPE, gather_id_maps, shard_clips


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_clips(n_clips, rank, world)
    # stand-in for the per-clip hot path: an id map that encodes the clip index
    local = torch.stack([torch.full((4, 6), c, dtype=torch.int64) for c in mine]) if mine else torch.zeros((0, 4, 6), dtype=torch.int64)
    allmaps = gather_id_maps(local, n_clips, dist)
    wire = gather_id_maps(local, n_clips, dist, wire_dtype=WIRE_DTYPE)          # int16 on the wire, same ids
    assert wire.dtype == WIRE_DTYPE and torch.equal(wire.to(torch.int64), allmaps)
    q.put((rank, mine, allmaps[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_shard_clips_partition():
    for n, w in [(300, 8), (300, 4), (7, 2), (1, 2), (5, 8)]:
        seen = []
        for r in range(w):
            seen += shard_clips(n, r, w)
        assert seen == list(range(n))
        sizes = [len(shard_clips(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_gloo():
    world, n_clips = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mine, order in res:
        assert order == list(range(n_clips)), (rank, order)          # every rank sees every clip, in clip order
    assert sorted(sum((m for _, m, _ in res), [])) == list(range(n_clips))
