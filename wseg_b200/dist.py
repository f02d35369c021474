"""Data-parallel plumbing for the one collective on the path: the prototype candidate exchange
(SURVEY.md §8e).  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def rank_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_candidates(top_val: torch.Tensor, top_idx: torch.Tensor, top_fea: torch.Tensor, group=None):
    """All-gather every rank's [C,k] candidates (values, global pixel ids, features) and concatenate them in rank
    order along dim 1 -> [C, R*k(,D)].  Same bytes on every rank, so the finalised prototypes are bit-identical."""
    _, world = rank_world(group)
    c, k = top_val.shape
    d = top_fea.shape[-1]
    # one packed message per rank: [C, k, D + 2] (value, index bits, features)
    packed = torch.empty(c, k, d + 2, dtype=torch.float32, device=top_val.device)
    packed[:, :, 0] = top_val
    packed[:, :, 1] = top_idx.view(torch.float32)
    packed[:, :, 2:] = top_fea
    out = torch.empty(world * c, k, d + 2, dtype=torch.float32, device=top_val.device)   # rank-major concat
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.reshape(world, c, k, d + 2).permute(1, 0, 2, 3).reshape(c, world * k, d + 2)
    val = out[:, :, 0].contiguous()
    idx = out[:, :, 1].contiguous().view(torch.int32)
    fea = out[:, :, 2:].contiguous()
    return val, idx, fea


class SymmExchange:
    """One-shot prototype-candidate exchange over NVLink peer memory (include/wsegk.h: wsegk_proto_push /
    wsegk_proto_finalize_symm): every rank's top-k kernel output is stored straight into every peer's buffer, so a step has
    no NCCL kernel and no host involvement on this path.  Built on ``torch.distributed._symmetric_memory`` for the
    allocation and the peer mapping only."""

    _cache = {}

    def __init__(self, group, c: int, k: int, d: int, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        nbytes = int(_lib.lib().wsegk_proto_exchange_bytes(self.world, c, k, d))
        self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()
        self.state = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.hdl.barrier()                       # everybody's flags are zero before anybody pushes
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.peer_bases = (ctypes.c_void_p * self.world)(*ptrs)
        self.dims = (c, k, d)

    @classmethod
    def get(cls, group, c: int, k: int, d: int, device):
        key = (id(group), c, k, d, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(group, c, k, d, device)
        return cls._cache[key]


class MineExchange:
    """Symmetric buffer for the GLOBAL hard-pixel mining (include/wsegk.h: wsegk_contrast_loss_fwd_bwd_global): every rank's
    per-pixel (similarity, key, pseudo label) of both views lands in every peer's buffer straight from the kernel."""

    _cache = {}

    def __init__(self, group, n: int, h: int, w: int, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        nbytes = int(_lib.lib().wsegk_mine_exchange_bytes(self.world, n, h, w))
        self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()
        self.state = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.hdl.barrier()                       # everybody's flags are zero before anybody pushes
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.peer_bases = (ctypes.c_void_p * self.world)(*ptrs)

    @classmethod
    def get(cls, group, n: int, h: int, w: int, device):
        key = (id(group), n, h, w, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(group, n, h, w, device)
        return cls._cache[key]
