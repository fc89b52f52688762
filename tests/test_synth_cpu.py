"""The numpy data generators agree with the oracle's arithmetic (so the synthetic databases written for the reference
binary equal the ones generated in HBM)."""
import numpy as np

from ganon_b200 import synth
from oracle import oracle as O


def test_ibf_rows_and_emplace_match_oracle():
    rng = np.random.default_rng(0)
    for bin_size, h in [(1009, 3), (1 << 24, 4), (100003, 5), ((1 << 33) + 7, 2)]:
        ibf = O.OracleIBF(64, min(bin_size, 5003), h) if bin_size > 10**6 else O.OracleIBF(130, bin_size, h)
        hs = rng.integers(0, 1 << 38, size=200, dtype=np.uint64)
        # row function against the C oracle with the real bin_size
        c = O.GoIbf(64, 64, bin_size, 64 - int(bin_size).bit_length(), 1, h, None)
        import ctypes as C

        want = np.array([[O.lib().go_ibf_row(C.byref(c), int(x), i) for x in hs] for i in range(h)], dtype=np.uint64)
        assert np.array_equal(synth.ibf_rows(hs, h, bin_size), want)
    ibf = O.OracleIBF(130, 1543, 3)
    hs = rng.integers(0, 1 << 38, size=500, dtype=np.uint64)
    bins = rng.integers(0, 130, size=500)
    data = np.zeros_like(ibf.data)
    synth.emplace_numpy(data, ibf.bin_words, 1543, 3, hs, bins)
    for v, b in zip(hs, bins):
        ibf.emplace(int(v), int(b))
    assert np.array_equal(data, ibf.data)


def test_random_words_density_and_padding():
    w = synth.random_words(7, 2, 500, 3, 130).reshape(500, 3)
    assert (w[:, 2] >> np.uint64(2)).max() == 0
    bits = np.unpackbits(w[:, :2].view(np.uint8)).mean()
    assert 0.22 < bits < 0.28
    # row windows are consistent with the whole
    part = synth.random_words(7, 2, 500, 3, 130, row0=100, rows=50).reshape(50, 3)
    assert np.array_equal(part, w[100:150])


def test_fastq_block_and_reads():
    g = synth.random_genomes(1, 8, 2000)
    m1, m2, origin = synth.reads_from_genomes(2, g, 100, paired=True)
    assert m1.shape == (100, 150) and m2.shape == (100, 150) and (origin >= -1).all()
    blk = synth.fastq_block(m1, first_index=5, suffix=b"/1").tobytes()
    recs = blk.split(b"\n")
    assert recs[0] == b"@r000000005/1" and recs[1] == m1[0].tobytes() and recs[2] == b"+" and recs[3] == b"I" * 150
    assert len(recs) == 401 and recs[-1] == b""
    reads = O.parse_reads.__wrapped__ if hasattr(O.parse_reads, "__wrapped__") else None  # noqa: F841


def test_bench_hibf_layout_is_a_consistent_tree():
    """The synthetic 3-level HIBF of bench.py (workload c4): tables in the raptor layout, every user bin reachable from
    the top level through merged bins, chains list the bins that must hold a user bin's content."""
    import bench

    wl = dict(bench.WORKLOADS["c4tiny"])
    bins, rows, nxt, pos, chains = bench.hibf_layout(wl)
    n_ibf = len(bins)
    assert n_ibf == 1 + 256 + 256 * wl["child_merged"] and bins[0] == wl["top_bins"]
    assert sum(b * r for b, r in zip(bins, rows)) // 8 == (wl["top_bins"] * wl["top_rows"] + 256 * wl["child_bins"] * wl["child_rows"] + 1024 * wl["grand_bins"] * wl["grand_rows"]) // 8
    seen_users, seen_ibfs = set(), {0}
    stack = [0]
    while stack:
        i = stack.pop()
        assert len(nxt[i]) == len(pos[i]) == bins[i]
        for b in range(bins[i]):
            if pos[i][b] < 0:
                c = nxt[i][b]
                assert c not in seen_ibfs and 0 < c < n_ibf
                seen_ibfs.add(c)
                stack.append(c)
            else:
                assert nxt[i][b] == i
                seen_users.add(pos[i][b])
    assert seen_ibfs == set(range(n_ibf)) and seen_users == set(range(len(chains)))
    for u, chain in enumerate(chains):
        leaf, leaf_bins = chain[0]
        assert all(pos[leaf][b] == u for b in leaf_bins)
        child = leaf
        for parent, pb in chain[1:]:  # the merged bins above point down the chain
            assert len(pb) == 1 and pos[parent][pb[0]] < 0 and nxt[parent][pb[0]] == child
            child = parent
        assert child == 0
    assert any(len(c[0][1]) == 2 for c in chains)  # split user bins exist
