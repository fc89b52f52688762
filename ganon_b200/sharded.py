"""Bin-block sharded classification over several GPUs: one process per GPU, every rank holds the
bin-word columns [r*bw/N, (r+1)*bw/N) of every database row, stages the same read block, runs K2 + K3 on its columns,
and the sparse per-read tuples are all-gathered (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  Every rank
then runs the identical host finishing stage, so no second exchange is needed; rank 0 writes the output.

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


class ShardedSession:
    """Session over column shards; `classify` has the semantics of Session.classify on every rank."""

    def __init__(self, dbs: Sequence[Database], *args, group=None, **kwargs):
        self.group = group
        self.sess = Session(dbs, *args, **kwargs)
        self.level_labels = self.sess.level_labels

    @classmethod
    def open(cls, paths: Sequence[str], rank: int, world: int, device: int, *args, group=None, **kwargs) -> "ShardedSession":
        dbs = [Database.open(p, device=device, shard=rank, n_shards=world) for p in paths]
        return cls(dbs, *args, group=group, device=device, **kwargs)

    def classify(self, block1, block2=None, final: bool = True, prefix_id: int = 0, len1=None, len2=None):
        s = self.sess
        s.stage(block1, block2, final=final, len1=len1, len2=len2)
        exchanged = 0
        for li, nf in enumerate(s.n_filters_per_level):
            s.run_level(li)
            for fi in range(nf):
                merged = merge_tuples(s.level_tuples(li, fi), self.group)
                exchanged += merged.size * 8
                s.set_level_tuples(li, fi, merged)
            s.finish_level(li)
        res = s.collect_staged(prefix_id)
        self.last_exchanged_bytes = exchanged
        return res

    def __getattr__(self, name):
        return getattr(self.sess, name)
