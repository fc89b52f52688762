"""EM reassignment, reads that share an id (ganon_b200/csrc/em_merge.cpp compiled with g++, tests/native/em_merge_host.cpp):
src/ganon/reassign.py:78-85 collects the matches of the `.all` lines in a dictionary keyed by the read id, so reads with equal
ids are one read standing where the first of them stood, with their matches in file order.  The regrouped store must be what
that dictionary holds, and the restatement of the module (oracle/reassign_oracle.py, pinned to the reference module by
tests/test_reassign_cpu.py) must give the same `.one` text for the regrouped store as for the original lines."""
import ctypes as C
import os
import random
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("em_merge") / "em_merge_host.so")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "native", "em_merge_host.cpp")])
    lib = C.CDLL(so)
    lib.emh_merge.restype = C.c_uint64
    lib.emh_merge.argtypes = [C.c_uint64] + [C.c_void_p] * 10
    return lib


def merge(L, reads):
    """reads: [(id bytes, [(target, count)])] -> the same after em_merge_by_id."""
    n = len(reads)
    off = np.zeros(n + 1, dtype=np.uint64)
    id_off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for _i, m in reads])
    id_off[1:] = np.cumsum([len(i) for i, _m in reads])
    tgt = np.array([t for _i, m in reads for t, _c in m] + [0], dtype=np.uint32)
    cnt = np.array([c for _i, m in reads for _t, c in m] + [0], dtype=np.uint32)
    ids = np.frombuffer(b"".join(i for i, _m in reads) + b"\0", dtype=np.uint8).copy()
    o_off, o_id_off = np.zeros_like(off), np.zeros_like(id_off)
    o_tgt, o_cnt, o_ids = np.zeros_like(tgt), np.zeros_like(cnt), np.zeros_like(ids)
    g = L.emh_merge(n, *(a.ctypes.data for a in (off, tgt, cnt, id_off, ids, o_off, o_tgt, o_cnt, o_id_off, o_ids)))
    out = []
    for r in range(g):
        rid = bytes(o_ids[int(o_id_off[r]) : int(o_id_off[r + 1])])
        out.append((rid, [(int(o_tgt[j]), int(o_cnt[j])) for j in range(int(o_off[r]), int(o_off[r + 1]))]))
    return out


def random_reads(rng, n, n_ids, n_targets):
    pool = [b"read%d" % i if i % 3 else b"r%d extra text/%d" % (i, i % 2) for i in range(n_ids)] + [b""]
    return [(rng.choice(pool), [(rng.randrange(n_targets), rng.randrange(1, 200)) for _ in range(rng.choice((1, 1, 2, 3, 7)))]) for _ in range(n)]


@pytest.mark.parametrize("seed", range(12))
def test_regrouped_store_is_the_dictionary_of_the_reference(L, seed):
    rng = random.Random(seed)
    n = rng.choice((0, 1, 2, 50, 400))
    reads = random_reads(rng, n, max(1, n // rng.choice((1, 2, 10))), 9)
    want = {}
    for rid, m in reads:  # reassign.py:78-85
        want.setdefault(rid, []).extend(m)
    assert merge(L, reads) == list(want.items())


def test_unique_ids_come_back_unchanged(L):
    reads = [(b"a", [(1, 5)]), (b"b", [(2, 7), (1, 3)]), (b"ab", [(0, 1)]), (b"", [(4, 4)])]
    assert merge(L, reads) == reads


@pytest.mark.parametrize("seed", range(4))
def test_restatement_gives_the_same_one_file_for_the_regrouped_store(L, seed):
    from oracle import reassign_oracle as RO

    rng = random.Random(100 + seed)
    reads = [(i.replace(b" ", b"_") or b"x", m) for i, m in random_reads(rng, 300, 120, 6)]

    def text(rs):
        return "".join("%s\tT%d\t%d\n" % (rid.decode(), t, c) for rid, m in rs for t, c in m)

    rep = "".join("H1\tT%d\t1\t1\t0\n" % t for t in range(6)) + "#total_classified\t300\n#total_unclassified\t0\n"
    for threshold, max_iter in ((0, 10), (0.05, 0)):
        a = RO.reassign_texts(rep, {"": text(reads)}, threshold, max_iter)
        b = RO.reassign_texts(rep, {"": text(merge(L, reads))}, threshold, max_iter)
        assert a[0][""] == b[0][""]
