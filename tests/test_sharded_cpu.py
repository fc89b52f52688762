"""Host-side logic of the N > 1 path on CPU: the tuple exchange of ganon_b200/sharded.py over torch.distributed with
the gloo backend, world_size 2."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ganon_b200.sharded import merge_tuples

    rng = np.random.default_rng(100 + rank)
    # rank r owns nodes with node % world == r; partial tuples of a straddling node (node 7) come from both ranks
    n = 50 + 30 * rank
    reads = rng.integers(0, 40, size=n).astype(np.uint64)
    nodes = (rng.integers(0, 20, size=n) * world + rank).astype(np.uint64)
    counts = rng.integers(1, 100, size=n).astype(np.uint64)
    tup = (reads << np.uint64(40)) | (nodes << np.uint64(17)) | counts
    strad = (np.arange(5, dtype=np.uint64) << np.uint64(40)) | (np.uint64(7) << np.uint64(17)) | np.uint64(1 << 16) | np.uint64(3 + rank)
    local = np.concatenate([tup, strad])
    merged = merge_tuples(local)
    np.save(os.path.join(out_dir, "local%d.npy" % rank), local)
    np.save(os.path.join(out_dir, "merged%d.npy" % rank), merged)
    # empty contribution from one rank
    e = merge_tuples(local if rank == 0 else np.empty(0, dtype=np.uint64))
    np.save(os.path.join(out_dir, "e%d.npy" % rank), e)
    z = merge_tuples(np.empty(0, dtype=np.uint64))
    assert z.size == 0
    dist.barrier()
    dist.destroy_process_group()


def test_merge_tuples_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    l0, l1 = np.load(tmp_path / "local0.npy"), np.load(tmp_path / "local1.npy")
    m0, m1 = np.load(tmp_path / "merged0.npy"), np.load(tmp_path / "merged1.npy")
    assert np.array_equal(m0, m1)  # every rank ends with the same list
    assert sorted(m0.tolist()) == sorted(np.concatenate([l0, l1]).tolist())
    key = m0 >> np.uint64(17)
    assert (np.diff(key.astype(np.int64)) >= 0).all()  # sorted by (read, node)
    # the straddling node's partial sums are adjacent, one per rank
    for r in range(5):
        k = (np.uint64(r) << np.uint64(23)) | np.uint64(7)
        sel = m0[key == k]
        assert sel.size == 2 and sorted((sel & np.uint64(0xFFFF)).tolist()) == [3, 4] and ((sel >> np.uint64(16)) & np.uint64(1)).all()
    e0, e1 = np.load(tmp_path / "e0.npy"), np.load(tmp_path / "e1.npy")
    assert np.array_equal(e0, e1) and sorted(e0.tolist()) == sorted(l0.tolist())
