"""Bin-block sharded classification over several GPUs: one process per GPU, every rank holds the
bin-word columns [r*bw/N, (r+1)*bw/N) of every database row, stages the same read block, runs K2 + K3 on its columns,
and the sparse per-read tuples are all-gathered (torch.distributed; NCCL on GPUs -- inside HBM, followed by the sort and
K4 on every rank -- or gloo with host arrays in the CPU tests).  Every rank then holds the identical finished result,
so no second exchange is needed; rank 0 writes the output.

What crosses the link per batch is the tuple list -- 8 bytes per (read, target) candidate, typically < 1 per read --
not the per-bin count vectors (128 KiB/read at 65 536 bins) the reference layout would suggest.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .classify import Database, Session

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


class ShardedSession:
    """Session over column shards; `classify` has the semantics of Session.classify on every rank."""

    def __init__(self, dbs: Sequence[Database], *args, group=None, exchange: str = "auto", **kwargs):
        self.group = group
        self.exchange = exchange  # "auto": in HBM over NCCL when the backend is nccl; "host": numpy + all_gather (gloo tests)
        self.device = int(kwargs.get("device", 0))
        self.sess = Session(dbs, *args, **kwargs)
        self.level_labels = self.sess.level_labels

    @classmethod
    def open(cls, paths: Sequence[str], rank: int, world: int, device: int, *args, group=None, **kwargs) -> "ShardedSession":
        dbs = [Database.open(p, device=device, shard=rank, n_shards=world) for p in paths]
        return cls(dbs, *args, group=group, device=device, **kwargs)

    def _device_exchange(self) -> bool:
        if self.exchange == "host":
            return False
        import torch.distributed as dist

        return not (dist.is_available() and dist.is_initialized()) or dist.get_backend(self.group) == "nccl"

    def run_levels(self, prefix_id: int = 0):
        """K2/K3 per level on this rank's columns, tuple exchange, finishing stage; the batch must be staged."""
        s = self.sess
        exchanged = 0
        on_device = self._device_exchange()
        for li, nf in enumerate(s.n_filters_per_level):
            if on_device and nf == 1:
                # tuples never leave HBM: NCCL all-gather, sort + K4 on every rank
                s.run_level_device(li)
                ptr, n = s.level_tuples_device(li)
                merged = gather_tuples_device(ptr, n, self.device, self.group)
                exchanged += merged.numel() * 8
                s.set_level_tuples_device(li, merged.data_ptr() if merged.numel() else 0, merged.numel())
                s.finish_level_device(li, prefix_id)
                continue
            s.run_level(li)
            for fi in range(nf):
                merged = merge_tuples(s.level_tuples(li, fi), self.group)
                exchanged += merged.size * 8
                s.set_level_tuples(li, fi, merged)
            s.finish_level(li)
        self.last_exchanged_bytes = exchanged

    def classify(self, block1, block2=None, final: bool = True, prefix_id: int = 0, len1=None, len2=None):
        s = self.sess
        s.stage(block1, block2, final=final, len1=len1, len2=len2)
        self.run_levels(prefix_id)
        return s.collect_staged(prefix_id)

    def __getattr__(self, name):
        return getattr(self.sess, name)
