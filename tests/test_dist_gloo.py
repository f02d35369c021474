"""world_size-2 gloo test of the prototype candidate exchange plumbing (wseg_b200/dist.py) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wseg_b200 import dist as wdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c, k, d = 3, 4, 6
        g = torch.Generator().manual_seed(100 + rank)
        val = torch.rand(c, k, generator=g)
        idx = (torch.arange(c * k, dtype=torch.int32).reshape(c, k) + 1000 * rank)
        fea = torch.randn(c, k, d, generator=g)
        assert wdist.rank_world() == (rank, world)
        gv, gi, gf = wdist.gather_candidates(val, idx, fea)
        assert gv.shape == (c, world * k) and gi.dtype == torch.int32 and gf.shape == (c, world * k, d)
        # my own block sits at [rank*k, (rank+1)*k) and is bit-identical
        assert torch.equal(gv[:, rank * k:(rank + 1) * k], val)
        assert torch.equal(gi[:, rank * k:(rank + 1) * k], idx)
        assert torch.equal(gf[:, rank * k:(rank + 1) * k], fea)
        # every rank holds the same bytes
        flat = torch.cat([gv.reshape(-1), gi.view(torch.float32).reshape(-1), gf.reshape(-1)])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, flat)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_gather_candidates_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
