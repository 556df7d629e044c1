"""world_size-2 CPU test (gloo) of the host-side multi-rank logic bench.py uses: per-rank stream sharding
is disjoint and complete, and the fixed-size record slabs gather to rank 0 in rank order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    rng = np.random.default_rng(0)
    caps = [rng.integers(0, 256, 4 * 20000, dtype=np.uint8) for _ in range(3)]
    S = 4
    views, n = bench.stream_views(caps, S, rank)
    # identity of every stream: (capture index, byte offset)
    ids = []
    for s, v in enumerate(views):
        g = rank * S + s
        c = caps[g % 3]
        off = v.__array_interface__["data"][0] - c.__array_interface__["data"][0]
        assert np.array_equal(v, c[off: off + n])
        ids.append((g % 3, off))
    # the record slabs: rank r fills its slab with a rank/stream tag, rank 0 gathers them
    slab = torch.full((S, 64), rank, dtype=torch.uint8)
    for s in range(S):
        slab[s, 0] = rank * S + s
    lst = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
    dist.gather(slab, lst, dst=0)
    if rank == 0:
        tags = [int(t[s, 0]) for t in lst for s in range(S)]
        q.put(("tags", tags))
    allids = [None] * world
    dist.all_gather_object(allids, ids)
    if rank == 0:
        q.put(("ids", allids))
    dist.destroy_process_group()


def test_sharding_and_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["tags"] == list(range(8))                       # gathered in rank order, stream order
    flat = [tuple(x) for r in out["ids"] for x in r]
    assert len(set(flat)) == len(flat) == 8                    # global streams are distinct across ranks
