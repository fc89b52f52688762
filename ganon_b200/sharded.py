"""Bin-block sharded classification over several GPUs: one process per GPU, every rank holds the
bin-word columns [r*bw/N, (r+1)*bw/N) of every database row, sees the same read block, runs K2 + K3 on its columns,
and the sparse per-read tuples are all-gathered inside HBM by the library itself (NCCL C API, csrc/comm.cpp), followed
by the sort and K4 on every rank.  Every rank then holds the identical finished result, so no second exchange is
needed; rank 0 writes the output.  This module only creates the communicator (torch.distributed carries the NCCL id).

merge_tuples / gather_tuples_device are the host-array and torch forms of the same exchange for the level-wise C-ABI
calls (gnb_session_run_level .. gnb_session_finish_level): used by the CPU tests (gloo) and the single-GPU shard tests.

What crosses the link per batch is the tuple list -- 8 bytes per (read, target) candidate, typically < 1 per read --
not the per-bin count vectors (128 KiB/read at 65 536 bins) the reference layout would suggest.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .classify import Comm, Database, Session

TUPLE_KEY_SHIFT = 17  # (read, node) = bits [17, 64) of a tuple


def merge_tuples(local: np.ndarray, group=None, device: Optional[str] = None) -> np.ndarray:
    """All-gather variable-length uint64 tuple arrays and return them sorted by (read, node)."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        merged = local
    else:
        world = dist.get_world_size(group)
        dev = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
        n = torch.tensor([local.size], dtype=torch.int64, device=dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = [int(x.item()) for x in sizes]
        m = max(sizes)
        if m == 0:
            return np.empty(0, dtype=np.uint64)
        buf = torch.zeros(m, dtype=torch.int64, device=dev)
        if local.size:
            buf[: local.size] = torch.from_numpy(local.view(np.int64)).to(dev)
        parts = [torch.empty(m, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        merged = np.concatenate([p[:k].cpu().numpy().view(np.uint64) for p, k in zip(parts, sizes)])
    order = np.argsort(merged >> np.uint64(TUPLE_KEY_SHIFT), kind="stable")
    return merged[order]


class _DeviceWords:
    """A device pointer seen as a 1-d int64 array (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def gather_tuples_device(ptr: int, n: int, device: int, group=None):
    """All-gather of the ranks' tuple lists inside HBM (NCCL): returns a torch int64 tensor on `device` holding the
    concatenation (rank order; the library sorts it).  Lists are padded to the longest one for the collective."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", device)
    local = torch.as_tensor(_DeviceWords(ptr, n), device=dev) if n else torch.empty(0, dtype=torch.int64, device=dev)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    sizes = sizes.tolist()
    m = max(sizes)
    if m == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    buf = torch.zeros(m, dtype=torch.int64, device=dev)
    buf[:n] = local
    out = torch.empty(world * m, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    merged = out if all(k == m for k in sizes) else torch.cat([out[r * m : r * m + k] for r, k in enumerate(sizes)])
    torch.cuda.current_stream(dev).synchronize()
    return merged


def make_comm(device: int, group=None) -> Comm:
    """The library's communicator for this process, with torch.distributed as the plumbing that carries the id: rank 0
    draws it (gnb_comm_unique_id), broadcast_object_list hands it to the other ranks, gnb_comm_create is collective."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return Comm(Comm.unique_id(), 0, 1, device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return Comm(box[0], rank, world, device)


class ShardedSession(Session):
    """Session over this rank's column shards.  The tuple exchange between the ranks happens inside libganon_b200
    (NCCL, see include/ganon_b200.h): every entry point of Session -- classify, stage / run_staged / finish_staged,
    submit / collect -- works unchanged, provided all ranks make the same calls on the same blocks; every rank gets
    the complete result.  sliced_ingest: each rank copies only its 1/n of a block to its GPU (all-gathered over NVLink)."""

    def __init__(self, dbs: Sequence[Database], *args, comm: Optional[Comm] = None, group=None, sliced_ingest: bool = False, **kwargs):
        self.comm = comm if comm is not None else make_comm(int(kwargs.get("device", 0)), group)
        super().__init__(dbs, *args, comm=self.comm, sliced_ingest=sliced_ingest, **kwargs)

    @classmethod
    def open(cls, paths: Sequence[str], rank: int, world: int, device: int, *args, group=None, comm: Optional[Comm] = None, **kwargs) -> "ShardedSession":
        dbs = [Database.open(p, device=device, shard=rank, n_shards=world) for p in paths]
        return cls(dbs, *args, group=group, comm=comm, device=device, **kwargs)
