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
