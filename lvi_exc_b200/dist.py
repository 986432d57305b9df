"""One process per GPU: torch.distributed (NCCL) is the bootstrap, the library gets its own communicator.

`make_backend()` reads RANK / LOCAL_RANK / WORLD_SIZE (torchrun), initialises the process group when world > 1, broadcasts an
ncclUniqueId made by the library on rank 0 and returns (CudaBackend, dist module or None, rank, world)."""
from __future__ import annotations

import ctypes as C
import os


def make_backend():
    import torch
    from . import _capi
    from .backend import CudaBackend
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device — this library has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        lib = _capi.load()
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = C.create_string_buffer(128)
            _capi.check(lib.lvi_nccl_unique_id(raw))
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        nccl_id = bytes(idbuf.cpu().numpy().tobytes())
    return CudaBackend(local_rank, nccl_id=nccl_id, rank=rank, world=world), dist, rank, world


def shard_range(n: int, rank: int, world: int):
    """contiguous (= time-contiguous) chunk of n items for `rank` — lowering.hpp::shard_range"""
    return n * rank // world, n * (rank + 1) // world
