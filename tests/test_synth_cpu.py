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


def _synthdb(tmp_path, bins, bin_size, glen, h=4, k=19, w=31, threads=3):
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle"), "synthdb"])
    g = synth.random_genomes(1, bins, glen)
    gfile, out = str(tmp_path / "g.bin"), str(tmp_path / "t.ibf")
    g.tofile(gfile)
    done = subprocess.run([os.path.join(root, "oracle", "synthdb"), out, gfile, str(bins), str(bin_size), str(h), str(k), str(w), str(glen), "1", "1234", str(threads)],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert done.returncode == 0, done.stderr
    return g, out, [int(x) for x in done.stdout.split()]


def test_cpu_database_writer_equals_the_numpy_statement_of_the_device_generators(tmp_path):
    """oracle/synthdb (the reference arm's database writer) = random_words (gnb_db_fill_random) OR emplace of every
    genome's minimisers into its bin -- the filter bench.build_database creates in HBM -- in the reference's .ibf layout."""
    from ganon_b200 import formats

    for bins, bin_size, glen in ((200, 50021, 3000), (70, 1 << 15, 500), (64, 1 << 23, 400)):  # the last one spans several 32 MiB chunks
        h, k, w = 4, 19, 31
        g, path, (n_words, xor, total, _planted) = _synthdb(tmp_path, bins, bin_size, glen)
        f = formats.read_ibf(path)
        bw = (bins + 63) // 64
        want = synth.random_words(1, 1, bin_size, bw, bins)
        hs, bb = [], []
        for i in range(bins):
            u = O.minimiser_hash(g[i].tobytes(), k, w)
            hs.append(u)
            bb.append(np.full(u.size, i, np.uint32))
        synth.emplace_numpy(want, bw, bin_size, h, np.concatenate(hs), np.concatenate(bb))
        assert n_words == want.size and np.array_equal(f.ibf.data, want)
        assert xor == int(np.bitwise_xor.reduce(want)) and total == int(want.sum(dtype=np.uint64))
        assert (f.kmer_size, f.window_size, f.ibf.hash_funs, f.max_hashes_bin) == (k, w, h, 1234)
        assert f.bin_map == [(b, "T%d" % b) for b in range(bins)] and f.hashes_count == [("T%d" % b, 1234) for b in range(bins)]


def test_reference_arm_runs_without_the_product_library(tmp_path):
    """bench.py --impl reference: one invocation of the unmodified binary on a database written on the CPU; the process
    reports which libganon_b200 objects it mapped -- none."""
    import json
    import os
    import subprocess
    import sys

    import pytest

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "ganon-classify")):
        pytest.skip("oracle/_ref/ganon-classify not built")
    env = dict(os.environ, GANON_B200_BENCH_DIR=str(tmp_path))
    done = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert done.returncode == 0, done.stderr[-2000:]
    line = json.loads(done.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["native_so_loaded"] == [] and line["value"] > 0
    assert line["config"]["invocations"] == 1 and line["e2e"]["h2d_bytes_per_step"] == 0
