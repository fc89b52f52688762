"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (ganon_b200._lib);
the checker is oracle/ (pinned to the reference in tests/test_oracle.py) and the committed outputs of the unmodified
reference binary (tests/golden/expected).  Integer work: bit-exact."""
import ctypes as C
import glob
import gzip
import os

import numpy as np
import pytest

from ganon_b200 import _lib, cli, formats
from ganon_b200.classify import Database, Session, minimisers, minimisers_batch, result_text
from oracle import oracle as O
from tests import scenario_util as SU

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------------------------ K2
def test_minimisers_seqan3_kats():
    # libs/seqan3/test/unit/search/views/minimiser_hash_test.cpp:62-79 with the ganon seed (adjust_seed)
    for seq, k, w in [(b"ACGGCGACGTTTAG", 4, 8), (b"ACGTCGACGTTTAG", 4, 8), (b"A" * 19, 4, 8), (b"ACGGCGACG", 4, 8), (b"A" * 19, 19, 19)]:
        assert minimisers(seq, k, w).tolist() == O.minimiser_hash(seq, k, w).tolist(), (seq, k, w)
    assert minimisers(b"AC", 4, 8).tolist() == []
    assert minimisers(b"A" * 19, 19, 19).tolist() == [min(0 ^ O.adjust_seed(19), (4**19 - 1) ^ O.adjust_seed(19))]


@pytest.mark.parametrize("k,w", [(19, 31), (10, 10), (4, 8), (32, 40), (21, 63), (5, 260), (28, 35), (31, 31), (27, 29), (26, 33), (12, 14)])
def test_minimisers_random_and_adversarial(k, w):
    rng = np.random.default_rng(k * 1000 + w)
    seqs = []
    for L in [w, w + 1, 75, 150, 151, 255, 256, 257, 300, 1000, 5000]:
        if L >= w:
            seqs.append(bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8)))
    seqs += [b"A" * 200, b"C" * 150, b"AC" * 100, b"ACG" * 80, b"ACGT" * 70, b"AT" * 300, (b"ACGGT" * 7 + b"T") * 30, b"N" * 150]
    seqs.append(bytes(rng.choice(list(b"ACGTNRYSWKMBDHVUacgtn"), size=400).astype(np.uint8)))
    seqs.append(bytes(rng.choice(list(b"AT"), size=700).astype(np.uint8)))  # many ties
    seqs.append((bytes(rng.choice(list(b"ACGT"), size=37).astype(np.uint8)) * 40))  # long period repeat, multi-tile
    for s in seqs:
        if len(s) < w:
            continue
        assert minimisers(s, k, w).tolist() == O.minimiser_hash(s, k, w).tolist(), (k, w, len(s), s[:40])


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_either_minimiser_kernel_alone_passes_the_k2_and_scenario_tests(kernel):
    """For k <= 29, w-k+1 <= 32 the library runs the thread-per-read kernel (k2_thread.cuh), over segments when the batch holds
    long reads, and the warp-per-read kernel for every other (k, w).  GANON_B200_K2=warp / =thread (read once per process) pin
    one kernel on whole reads: re-run the K2 tests, the golden scenarios (single, paired, FASTA, several levels) and the oracle
    session tests of this file in a child process under each, so that all three stay pinned on short and on long sequences."""
    import subprocess
    import sys

    if os.environ.get("GANON_B200_K2", ""):
        pytest.skip("already inside the child run")
    env = dict(os.environ, GANON_B200_K2=kernel)
    sel = "minimisers_seqan3 or minimisers_random or long_sequence or long_read_session or (golden_scenarios and device) or session_matches_oracle or device_and_host_record_index"
    done = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k", sel, "-p", "no:cacheprovider"], env=env,
                          cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert done.returncode == 0, done.stdout[-2000:]


def test_long_sequence_batches_hash_like_the_oracle():
    """A batch of a few long reads and the same reads inside batches of many short ones (K2t over segments of 512 windows;
    GANON_B200_K2=thread / =warp walk every read whole): same minimisers as the oracle either way."""
    rng = np.random.default_rng(77)
    long_reads = [bytes(rng.choice(list(b"ACGTN"), p=[0.2499, 0.2499, 0.2499, 0.2499, 0.0004], size=n).astype(np.uint8)) for n in (30_000, 2_000, 70_001, 640)]
    want = [O.minimiser_hash(s, 19, 31) for s in long_reads]

    def batch(seqs):
        hoff, h = minimisers_batch(seqs, 19, 31)
        return h, hoff

    h, hoff = batch(long_reads)
    for i, wnt in enumerate(want):
        assert np.array_equal(h[int(hoff[i]) : int(hoff[i + 1])], wnt), i
    short = [bytes(rng.choice(list(b"ACGT"), size=40).astype(np.uint8)) for _ in range(70_000)]
    h, hoff = batch(short[:35_000] + long_reads + short[35_000:])
    for i, wnt in enumerate(want):
        j = 35_000 + i
        assert np.array_equal(h[int(hoff[j]) : int(hoff[j + 1])], wnt), i
    assert np.array_equal(h[: int(hoff[1])], O.minimiser_hash(short[0], 19, 31))
    # repeats: homopolymers and tandem repeats keep equal values in every window -- segments inside them cannot know the state
    # of the walk and hand their read to one thread; around their ends either outcome must be exact
    rnd = lambda n: bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))
    odd = [b"A" * 9000, b"TTAGGG" * 1500, rnd(1700) + b"C" * 2600 + rnd(3000), rnd(5000) + b"AT" * 900, b"GA" * 333 + rnd(4000) + b"N" * 1500 + rnd(30),
           bytes(rng.choice(list(b"AT"), size=6000).astype(np.uint8)), rnd(37) * 200]
    for k, w in ((19, 31), (4, 8), (29, 60), (12, 12), (10, 41)):
        hoff, h = minimisers_batch(odd + long_reads, k, w)
        for i, s in enumerate(odd + long_reads):
            assert np.array_equal(h[int(hoff[i]) : int(hoff[i + 1])], O.minimiser_hash(s, k, w)), (k, w, i)
    # a big batch keeps reads of a few thousand bases on a thread each and cuts only from 4096 windows on
    many = short + short
    hoff, h = minimisers_batch(many[:100] + [long_reads[1], long_reads[0]] + many[100:], 19, 31)
    for j, s in ((100, long_reads[1]), (101, long_reads[0]), (102, many[100]), (len(many) + 1, many[-1])):
        assert np.array_equal(h[int(hoff[j]) : int(hoff[j + 1])], O.minimiser_hash(s, 19, 31)), j


# ------------------------------------------------------------------------------------------------------------------ K3
def _random_db(rng, bins, bin_size, h, density_terms=2, k=19, w=31):
    db = Database.create(bins, bin_size, h, k, w)
    db.fill_random(int(rng.integers(1, 1 << 40)), density_terms)
    return db


def _oracle_ibf(db):
    i = db.info()
    return O.OracleIBF(i.bins, i.bin_size_bits, i.hash_functions, db.read_words(0, i.bin_size_bits * i.bin_words))


@pytest.mark.parametrize(
    "bins,bin_size,h,nmax",
    [(64, 1009, 4, 40), (70, 5003, 3, 40), (130, 100003, 3, 60), (192, 4099, 5, 30), (1000, 2053, 2, 50), (4096, 1031, 4, 40), (4100, 521, 1, 20), (9000, 263, 4, 300), (65, 997, 4, 700)],
)
def test_bulk_count_matches_oracle(bins, bin_size, h, nmax):
    rng = np.random.default_rng(bins * 7 + h)
    db = _random_db(rng, bins, bin_size, h)
    oibf = _oracle_ibf(db)
    # padding bins must be empty (IBF.hpp:238-240)
    words = oibf.data.reshape(bin_size, -1)
    if bins % 64:
        assert (words[:, -1] >> np.uint64(bins % 64)).max() == 0
    lens = [0, 1, 2, 3, 4, 5, 7, 8, 9, nmax] + list(rng.integers(0, nmax, size=12))
    lists = [rng.integers(0, 1 << 38, size=int(n), dtype=np.uint64) for n in lens]
    # some repeated hashes (duplicates are counted, GC.cpp:693-703)
    lists.append(np.repeat(lists[-1][:3], 5))
    off = np.zeros(len(lists) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in lists])
    got = db.bulk_count(np.concatenate(lists), off)
    for i, hs in enumerate(lists):
        want = oibf.bulk_count(hs)
        assert np.array_equal(got[i], want), (i, hs.size)
    db.close()


def test_fill_random_matches_numpy_generator_and_save_load(tmp_path):
    from ganon_b200 import synth

    db = Database.create(130, 1543, 3, 19, 31)
    db.fill_random(99, 2)
    i = db.info()
    words = db.read_words(0, i.bin_size_bits * i.bin_words)
    assert np.array_equal(words, synth.random_words(99, 2, i.bin_size_bits, i.bin_words, 130))
    hashes = np.arange(1000, 1100, dtype=np.uint64)
    db.emplace(hashes, np.arange(100, dtype=np.uint32))
    o = _oracle_ibf(db)
    for h_, b in zip(hashes, range(100)):
        assert o.bulk_count(np.array([h_], dtype=np.uint64))[b] == 1
    p = str(tmp_path / "x.ibf")
    db.save(p)
    f = formats.read_ibf(p)  # the file parses with the independent python reader
    assert f.ibf.bins == 130 and np.array_equal(f.ibf.data, db.read_words(0, i.bin_size_bits * i.bin_words))
    db2 = Database.open(p)
    assert np.array_equal(db2.read_words(0, i.bin_size_bits * i.bin_words), f.ibf.data)
    assert [t[0] for t in db2.targets()] == [t[0] for t in db.targets()]
    # column shards hold exactly their slice
    sh = Database.open(p, shard=1, n_shards=3)
    si = sh.info()
    full = f.ibf.data.reshape(i.bin_size_bits, i.bin_words)
    assert (si.shard_word_begin, si.shard_word_end) == (1, 2)
    assert np.array_equal(sh.read_words(0, i.bin_size_bits), full[:, 1])


# ------------------------------------------------------------------------------------------------------------------ end to end
def _read_sorted(path):
    with open(path) as f:
        return sorted(l.rstrip("\n") for l in f)


FINISH_MODES = {
    # K4 on the device for every single-filter level (the default)
    "device": {},
    # the host finishing stage for every level
    "host": {"GANON_B200_HOST_FINISH": "1"},
    # K4 first; every --fpr-query value counts as "too close to the threshold", so levels with --fpr-query < 1 are
    # handed to the host finishing stage after the device pass (the path a borderline libm value takes)
    "fallback": {"GANON_B200_FPR_BAND": "2"},
}


@pytest.fixture(params=sorted(FINISH_MODES))
def finish_mode(request, monkeypatch):
    for k, v in FINISH_MODES[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


@pytest.mark.parametrize("name", sorted(SU.load_scenarios()))
def test_golden_scenarios_match_reference_outputs(name, golden_dbs, tmp_path, finish_mode):
    """The `ganon-classify` drop-in on the arguments the reference binary was run with; every output file the
    reference wrote must exist with identical (sorted) lines, and no extra file may appear."""
    args = SU.expand(SU.load_scenarios()[name], golden_dbs)
    pre = str(tmp_path / name)
    assert cli.main(args + ["-o", pre, "-t", "4", "--quiet"]) == 0
    want_files = sorted(os.path.basename(p)[len(name) :] for p in glob.glob(os.path.join(SU.GOLDEN, "expected", name + ".*")))
    got_files = sorted(os.path.basename(p)[len(name) :] for p in glob.glob(pre + ".*"))
    assert got_files == want_files
    for ext in want_files:
        assert _read_sorted(pre + ext) == SU.expected_lines(name, ext[1:]), ext


def test_session_matches_oracle_multibin_targets_and_blocks(finish_mode):
    """Synthetic DB whose targets span 1..5 bins (crossing 32-bin registers, lanes and 4096-bin chunks), paired
    reads, fed in several blocks; structured result and text vs the oracle."""
    rng = np.random.default_rng(3)
    k, w, h = 19, 31, 3
    bins = 4300
    db = Database.create(bins, 3001, h, k, w)
    db.fill_random(5, 3)
    bin_target, names, b = [], [], 0
    while b < bins:
        nb = int(min(rng.choice([1, 1, 2, 3, 5, 40]), bins - b))
        names.append("tgt%d" % len(names))
        bin_target += [len(names) - 1] * nb
        b += nb
    genomes = [bytes(rng.choice(list(b"ACGT"), size=1500).astype(np.uint8)) for _ in names]
    hs, bs, counts = [], [], []
    first = np.searchsorted(bin_target, np.arange(len(names)))
    nbins = np.bincount(bin_target)
    for t, g in enumerate(genomes):
        u = np.unique(O.minimiser_hash(g, k, w))
        hs.append(u)
        bs.append((first[t] + np.arange(u.size) % nbins[t]).astype(np.uint32))
        counts.append(u.size)
    db.emplace(np.concatenate(hs), np.concatenate(bs))
    db.set_targets(names, bin_target, counts, int(max(-(-c // n) for c, n in zip(counts, nbins))))
    reads = []
    for i in range(600):
        g = genomes[int(rng.integers(0, len(names)))]
        p = int(rng.integers(0, len(g) - 300))
        frag = g[p : p + 300]
        m2 = frag[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))[:150]
        m1 = bytearray(frag[:150])
        if i % 3 == 0:
            m1[int(rng.integers(0, 150))] = ord("N")
        if i % 50 == 0:
            m2 = m2[:20]  # mate shorter than the window: only mate 1 counts
        reads.append((b"p%d/1" % i, bytes(m1), m2))
    reads.append((b"short", b"ACGTACGT", b"ACGT" * 40))  # mate 1 shorter than the window: skipped
    fq1 = [b"@%s\n%s\n+\n%s\n" % (i, a, b"F" * len(a)) for i, a, _ in reads]
    fq2 = [b"@%s\n%s\n+\n%s\n" % (i, c, b"F" * len(c)) for i, _, c in reads]
    info = db.info()
    oibf = O.OracleIBF(info.bins, info.bin_size_bits, info.hash_functions, db.read_words(0, info.bin_size_bits * info.bin_words))
    tb = [[j for j, t in enumerate(bin_target) if t == ti] for ti in range(len(names))]
    fpr = [t[1] for t in db.targets()]
    assert [t[0] for t in db.targets()] == names
    for cutoff, relf, fq in [(0.0, 1.0, 1.0), (0.3, 0.2, 1.0), (0.1, 0.5, 0.2)]:
        filt = O.OracleFilter(oibf, names, tb, fpr, cutoff, k, w)
        want = O.classify_level([filt], reads, relf, fq)
        sess = Session([db], [cutoff], [relf], [fq], output_all=True, output_unclassified=True)
        got_all, got_unc, n_hashes = [], [], []
        cuts = [0, 200, 201, len(reads)]
        for a, b_ in zip(cuts[:-1], cuts[1:]):
            # blocks carry a partial record at the end unless final
            tail = b"" if b_ == len(reads) else fq1[b_][:15]
            tail2 = b"" if b_ == len(reads) else fq2[b_][:9]
            res = sess.classify(b"".join(fq1[a:b_]) + tail, b"".join(fq2[a:b_]) + tail2, final=b_ == len(reads))
            assert res.n_reads == b_ - a
            assert res.consumed1 == sum(map(len, fq1[a:b_])) and res.consumed2 == sum(map(len, fq2[a:b_]))
            got_all += result_text(res, "all").decode().splitlines()
            got_unc += result_text(res, "unc").decode().splitlines()
            n_hashes += [res.n_hashes[i] for i in range(res.n_reads)]
            # structured CSR agrees with the text
            nm = res.match_off[res.n_reads]
            assert nm == len(result_text(res, "all").decode().splitlines())
            lines = result_text(res, "all").decode().splitlines()
            for i in range(res.n_reads):  # CSR rows are the read's lines, in order
                for j in range(res.match_off[i], res.match_off[i + 1]):
                    assert lines[j].split("\t")[1:] == [names[res.match_target[j]], str(res.match_count[j])]
                assert (res.read_level[i] == 0) == (res.match_off[i + 1] > res.match_off[i])
            want_dev = {"device": 1, "host": 0, "fallback": 1 if fq >= 1.0 else 0}[finish_mode]
            assert res.levels_on_device == want_dev, (finish_mode, fq)
        assert n_hashes == [r["n_hashes"] for r in want]
        assert sorted(got_all) == O.all_lines(want), (cutoff, relf, fq)
        assert sorted(got_unc) == sorted(r["id"].decode() for r in want if not r["matches"])
        t = sess.totals()
        assert t.input_seqs == len(reads) and t.seqs_skipped_small == 1 and t.seqs_classified == sum(1 for r in want if r["matches"])
        assert t.discarded_matches_filter == sum(len(r["discarded_filter"]) for r in want)
        assert t.discarded_matches_fprquery == sum(len(r["discarded_fpr"]) for r in want)
        sess.close()
    db.close()


def test_long_read_session_matches_oracle():
    """Long reads (tens of kbp: K2t over segments of 512 windows), single and paired, mixed with short ones and with reads that
    hold homopolymers and tandem repeats (segments inside them are flagged and their read is walked from the start): hash counts
    and classification like the oracle's."""
    rng = np.random.default_rng(41)
    k, w, h = 19, 31, 3
    names = ["g%d" % i for i in range(24)]
    genomes = [bytes(rng.choice(list(b"ACGT"), size=40_000).astype(np.uint8)) for _ in names]
    db = Database.create(len(names), 60_013, h, k, w)
    hs, bs, counts = [], [], []
    for t, g in enumerate(genomes):
        u = np.unique(O.minimiser_hash(g, k, w))
        hs.append(u)
        bs.append(np.full(u.size, t, dtype=np.uint32))
        counts.append(u.size)
    db.emplace(np.concatenate(hs), np.concatenate(bs))
    db.set_targets(names, list(range(len(names))), counts, max(counts))
    rc = lambda x: x[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))
    reads = []
    for i in range(60):
        g = genomes[int(rng.integers(0, len(names)))]
        n = int(rng.choice([150, 600, 1100, 5_000, 12_345, 30_000]))
        p = int(rng.integers(0, len(g) - n + 1))
        m1 = bytearray(g[p : p + n])
        for _ in range(n // 400):
            m1[int(rng.integers(0, n))] = rng.choice(list(b"ACGTN"))
        if i % 7 == 0 and n > 3000:  # a homopolymer / a tandem repeat over a few segments, or up to the end of the read
            a = int(rng.integers(0, n // 2))
            rep = (b"A", b"TTAGGG", b"CA")[i // 7 % 3]
            ln = n - a if i % 14 == 0 else int(rng.integers(700, 2500))
            m1[a : a + ln] = (rep * (ln // len(rep) + 1))[: min(ln, n - a)]
        m2 = rc(g[p : p + n])[: int(rng.choice([20, 150, 4000, n]))]
        reads.append((b"lr%d" % i, bytes(m1), m2))
    info = db.info()
    oibf = O.OracleIBF(info.bins, info.bin_size_bits, info.hash_functions, db.read_words(0, info.bin_size_bits * info.bin_words))
    fpr = [t[1] for t in db.targets()]
    filt = O.OracleFilter(oibf, names, [[t] for t in range(len(names))], fpr, 0.2, k, w)
    for paired in (False, True):
        rs = reads if paired else [(i, a, None) for i, a, _ in reads]
        want = O.classify_level([filt], rs, 0.1, 1.0)
        fq1 = b"".join(b"@%s\n%s\n+\n%s\n" % (i, a, b"F" * len(a)) for i, a, _ in rs)
        fq2 = b"".join(b"@%s\n%s\n+\n%s\n" % (i, c, b"F" * len(c)) for i, _, c in rs) if paired else None
        sess = Session([db], [0.2], [0.1], [1.0], output_all=True, output_unclassified=True)
        res = sess.classify(fq1, fq2, final=True)
        assert [res.n_hashes[i] for i in range(res.n_reads)] == [r["n_hashes"] for r in want], paired
        assert max(r["n_hashes"] for r in want) > 3000
        assert sorted(result_text(res, "all").decode().splitlines()) == O.all_lines(want)
        assert sorted(result_text(res, "unc").decode().splitlines()) == sorted(r["id"].decode() for r in want if not r["matches"])
        sess.close()
    db.close()


def test_fasta_multiline_and_gz_inputs(golden_dbs, tmp_path):
    import gzip
    import shutil

    # same reads as FASTQ, gz FASTQ, multi-line FASTQ: identical results
    src = os.path.join(SU.GOLDEN, "reads.se.fq")
    gz = str(tmp_path / "r.fq.gz")
    with open(src, "rb") as fi, gzip.open(gz, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    ml = str(tmp_path / "ml.fq")
    with open(src, "rb") as fi, open(ml, "wb") as fo:
        lines = fi.read().split(b"\n")
        for j in range(0, len(lines) - 1, 4):
            i_, s, p, q = lines[j : j + 4]
            wrap = lambda x, n: b"\n".join(x[o : o + n] for o in range(0, len(x), n))
            fo.write(i_ + b"\n" + wrap(s, 50) + b"\n" + p + b"\n" + wrap(q, 70) + b"\n")
    outs = []
    for f in (src, gz, ml):
        pre = str(tmp_path / ("o%d" % len(outs)))
        assert cli.main(["-r", f, "-i", golden_dbs["synth"], "-c", "0.25", "-d", "0.5", "-o", pre, "-a", "-u", "--quiet"]) == 0
        outs.append((_read_sorted(pre + ".all"), _read_sorted(pre + ".unc"), _read_sorted(pre + ".rep")))
    assert outs[0] == outs[1]
    assert outs[0] == outs[2]
    assert len(outs[0][0]) > 100


def test_parse_error_truncation_rule(golden_dbs, tmp_path):
    """A malformed record ends the file; the reference loses the --n-reads chunk being assembled (GC.cpp:1240-1283)."""
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 10, b"I" * 40) for i in range(1000)]
    recs[850] = b"@bad\nACGTXXXX\n+\nIIIIIIII\n"  # X is not a dna15 letter
    f = str(tmp_path / "bad.fq")
    open(f, "wb").write(b"".join(recs))
    pre = str(tmp_path / "o")
    assert cli.main(["-r", f, "-i", golden_dbs["synth"], "-o", pre, "-u", "--quiet", "--n-reads", "400"]) == 0
    rep = dict(l.split("\t") for l in open(pre + ".rep").read().splitlines() if l.startswith("#"))
    assert int(rep["#total_unclassified"]) + int(rep["#total_classified"]) == 800


@pytest.mark.parametrize("bad_at,kept", [(0, 0), (1, 0), (399, 0), (400, 0), (401, 400), (800, 400), (801, 800)])
def test_parse_error_chunk_rule_at_chunk_boundaries(golden_dbs, tmp_path, bad_at, kept):
    """The reference's reader looks one record ahead: record e fails inside the chunk that holds record e - 1, so
    floor((e - 1) / n_reads) * n_reads records survive (numbers confirmed with the reference binary on the CPU); a run
    without a single parsed read writes a `.rep` without totals."""
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 10, b"I" * 40) for i in range(1000)]
    recs[bad_at] = b"@bad\nACGTXXXX\n+\nIIIIIIII\n"
    f = str(tmp_path / "bad.fq")
    open(f, "wb").write(b"".join(recs))
    pre = str(tmp_path / "o")
    assert cli.main(["-r", f, "-i", golden_dbs["synth"], "-o", pre, "-u", "--quiet", "--n-reads", "400"]) == 0
    rep = dict(l.split("\t") for l in open(pre + ".rep").read().splitlines() if l.startswith("#"))
    assert int(rep.get("#total_unclassified", 0)) + int(rep.get("#total_classified", 0)) == kept
    assert bool(rep) == (kept > 0)


def test_device_and_host_record_index_agree(golden_dbs, tmp_path, monkeypatch):
    """K1 (device FASTQ index) and the host reader give identical results, including a final block without a
    trailing newline and blocks that end in the middle of a record."""
    args = SU.expand(SU.load_scenarios()["pe_synth_all"], golden_dbs)
    outs = []
    for host_index in ("0", "1"):
        monkeypatch.setenv("GANON_B200_HOST_INDEX", host_index)
        pre = str(tmp_path / ("o" + host_index))
        assert cli.main(args + ["-o", pre, "--quiet"]) == 0
        outs.append({e: _read_sorted(pre + "." + e) for e in ("all", "unc", "rep", "sta")})
    assert outs[0] == outs[1]
    assert outs[0]["all"] == SU.expected_lines("pe_synth_all", "all")
    # block boundaries inside records + missing final newline, through the session API
    fq1 = open(os.path.join(SU.GOLDEN, "reads.1.fq"), "rb").read().rstrip(b"\n")
    fq2 = open(os.path.join(SU.GOLDEN, "reads.2.fq"), "rb").read().rstrip(b"\n")
    db = Database.open(golden_dbs["synth"])
    res_all = []
    for host_index in ("0", "1"):
        monkeypatch.setenv("GANON_B200_HOST_INDEX", host_index)
        sess = Session([db], [0.0], [1.0], [1.0], output_all=True, output_unclassified=True)
        lines, o1, o2, total = [], 0, 0, 0
        step = 40000
        while True:
            e1, e2 = min(len(fq1), o1 + step), min(len(fq2), o2 + step + 777)
            final = e1 == len(fq1) and e2 == len(fq2)
            r = sess.classify(fq1[o1:e1], fq2[o2:e2], final=final)
            lines += result_text(r, "all").decode().splitlines()
            total += r.n_reads
            o1 += r.consumed1
            o2 += r.consumed2
            if final:
                break
            assert r.n_reads > 0
        res_all.append((total, sorted(lines)))
        sess.close()
    assert res_all[0] == res_all[1]
    assert res_all[0][1] == SU.expected_lines("pe_synth_all", "all")


# ------------------------------------------------------------------------------------------------------------------ EM reassignment
EM_SCENARIOS = ["pe_real4_all", "se_synth_all", "pe_synth_all", "pe_two_filters", "pe_three_filters", "hier_two_levels", "hier_two_levels_single", "se_synth"]
EM_CASES = [(n, "device") for n in EM_SCENARIOS] + [(n, m) for n in ("pe_synth_all", "hier_two_levels", "hier_two_levels_single") for m in ("host", "fallback")]


@pytest.mark.parametrize("name,mode", EM_CASES)
def test_em_reassign_in_hbm_matches_restatement(name, mode, golden_dbs, tmp_path, monkeypatch):
    """`--reassign-em` (matches kept in HBM, iterations on the device) against the restatement of src/ganon/reassign.py
    (pinned to the reference module by tests/test_reassign_cpu.py) run on the `.all` / `.rep` files of the same run.
    Single- and multi-filter levels, two hierarchy levels with split and single output files; K4 and host finishing."""
    from oracle import reassign_oracle as RO

    for k, v in FINISH_MODES[mode].items():
        monkeypatch.setenv(k, v)
    args = SU.expand(SU.load_scenarios()[name], golden_dbs)
    plain = str(tmp_path / "plain")
    assert cli.main(args + ["-o", plain, "-t", "4", "--quiet"]) == 0
    rep = open(plain + ".rep").read()
    have = [os.path.basename(p)[len("plain") + 1 :] for p in glob.glob(plain + ".*all")]
    labels = RO.all_files_of(rep, have)
    texts = {h: open(plain + ("." + h if h else "") + ".all").read() for h in labels}
    for threshold, max_iter in [(0, 10), (0, 1), (0.05, 0)]:
        em = str(tmp_path / ("em_%s_%s" % (threshold, max_iter)))
        assert cli.main(args + ["-o", em, "-t", "4", "--quiet", "--reassign-em", "--em-max-iter", str(max_iter), "--em-threshold", str(threshold)]) == 0
        ones, new_rep = RO.reassign_texts(rep, texts, threshold, max_iter)
        assert open(em + ".rep").read() == new_rep
        for h, one in ones.items():
            assert open(em + ("." + h if len(ones) > 1 else "") + ".one").read() == one, h
        for h in labels:  # the .all files of the EM run are the plain run's
            suffix = ("." + h if h else "") + ".all"
            assert open(em + suffix).read() == open(plain + suffix).read()


def test_em_reassign_makes_one_read_of_reads_that_share_an_id(golden_dbs, tmp_path):
    """src/ganon/reassign.py:78-85 keys the matches of the `.all` lines by read id: reads with equal ids are ONE read to it
    (its matches are theirs in file order, it counts once in the total weight, it gets one `.one` line).  The store in HBM
    keeps reads by position; `gnb_session_reassign` finds equal ids (sorted hashes on the device) and regroups the store
    (csrc/em_merge.cpp).  Every third record takes the id of the record before it, some ids appear three times."""
    from oracle import reassign_oracle as RO

    fq = open(os.path.join(SU.GOLDEN, "reads.se.fq"), "rb").read().split(b"\n")
    recs = [[fq[i], fq[i + 1], fq[i + 2], fq[i + 3]] for i in range(0, len(fq) - 1, 4)]
    for i in range(len(recs)):
        if i % 3 == 2 or i % 50 == 1:
            recs[i][0] = recs[i - 1][0]
    f = str(tmp_path / "dup.fq")
    open(f, "wb").write(b"".join(b"\n".join(r) + b"\n" for r in recs))
    args = SU.expand(SU.load_scenarios()["se_synth_all"], golden_dbs)
    args[args.index("-r") + 1] = f
    plain = str(tmp_path / "plain")
    assert cli.main(args + ["-o", plain, "-t", "4", "--quiet"]) == 0
    rep, all_text = open(plain + ".rep").read(), open(plain + ".all").read()
    classified = int(dict(l.split("\t") for l in rep.splitlines() if l.startswith("#"))["#total_classified"])
    for threshold, max_iter in [(0, 10), (0, 1), (0.05, 0)]:
        em = str(tmp_path / ("em_%s_%s" % (threshold, max_iter)))
        assert cli.main(args + ["-o", em, "-t", "4", "--quiet", "--reassign-em", "--em-max-iter", str(max_iter), "--em-threshold", str(threshold)]) == 0
        ones, new_rep = RO.reassign_texts(rep, {"": all_text}, threshold, max_iter)
        assert len(ones[""].splitlines()) < classified  # reads were merged
        assert open(em + ".one").read() == ones[""]
        assert open(em + ".rep").read() == new_rep
        assert open(em + ".all").read() == all_text


# ------------------------------------------------------------------------------------------------------------------ HIBF
def test_hibf_sub_ibf_counts_match_oracle(golden_dbs):
    h = formats.read_hibf(golden_dbs["synth_hibf"])
    db = Database.open(golden_dbs["synth_hibf"], hibf=True)
    assert db.info().n_ibfs == len(h.ibfs) and db.info().is_hibf == 1
    rng = np.random.default_rng(8)
    lists = [rng.integers(0, 1 << 38, size=int(n), dtype=np.uint64) for n in (0, 1, 5, 17, 40)]
    off = np.zeros(len(lists) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in lists])
    L = _lib.lib()
    for idx in (0, 1, len(h.ibfs) - 1):
        i = h.ibfs[idx]
        o = O.OracleIBF(i.bins, i.bin_size, i.hash_funs, i.data)
        got = np.zeros((len(lists), i.technical_bins), dtype=np.uint16)
        hs = np.concatenate(lists)
        _lib.check(L.gnb_db_bulk_count(db.handle, idx, hs.ctypes.data, off.ctypes.data, len(lists), got.ctypes.data))
        for j, x in enumerate(lists):
            assert np.array_equal(got[j], o.bulk_count(x)), (idx, j)


@pytest.mark.parametrize("name", sorted(SU.load_hibf_scenarios()))
def test_hibf_scenarios_match_reference_outputs(name, golden_dbs, tmp_path, finish_mode):
    args = SU.expand(SU.load_hibf_scenarios()[name], golden_dbs)
    pre = str(tmp_path / name)
    assert cli.main(args + ["-o", pre, "-t", "4", "--quiet"]) == 0
    want_files = sorted(os.path.basename(p)[len(name) :] for p in glob.glob(os.path.join(SU.GOLDEN, "expected", name + ".*")))
    got_files = sorted(os.path.basename(p)[len(name) :] for p in glob.glob(pre + ".*"))
    assert got_files == want_files
    for ext in want_files:
        assert _read_sorted(pre + ext) == SU.expected_lines(name, ext[1:]), ext


def test_hibf_created_in_hbm_matches_oracle(tmp_path, finish_mode):
    """gnb_db_create_hibf / emplace_ibf / save -> reopen: a 3-level HIBF whose sub-IBFs take every traversal kernel
    (1, 2 and 4 lanes per item, and a 2100-bin sub-IBF on the warp-per-item path), with user bins split inside a
    32-bin register, across registers and across bin-words; results vs the oracle's HIBF restatement."""
    rng = np.random.default_rng(21)
    k, w, h = 19, 31, 3
    bins, rows, nxt, pos, chains = [], [], [], [], []

    def new_ibf(nb, nr):
        bins.append(nb)
        rows.append(nr)
        nxt.append([len(bins) - 1] * nb)
        pos.append([0] * nb)
        return len(bins) - 1

    def fill(ibf, first, up, runs):
        b = first
        for r in runs:
            if b + r > bins[ibf]:
                break
            u = len(chains)
            for x in range(b, b + r):
                pos[ibf][x] = u
            chains.append([(ibf, list(range(b, b + r)))] + up)
            b += r
        while b < bins[ibf]:
            u = len(chains)
            pos[ibf][b] = u
            chains.append([(ibf, [b])] + up)
            b += 1

    top = new_ibf(200, 4001)
    c1, c2 = new_ibf(64, 3001), new_ibf(2100, 1009)
    g1 = new_ibf(130, 2003)
    nxt[top][0], pos[top][0] = c1, -1
    nxt[top][1], pos[top][1] = c2, -1
    nxt[c1][5], pos[c1][5] = g1, -1
    fill(top, 2, [], [1, 2, 3, 1, 40, 1, 1, 70])  # runs inside a register, across registers and across words
    fill(c1, 0, [(top, [0])], [2, 3])  # bins 0..4, bin 5 is merged
    fill(c1, 6, [(top, [0])], [1, 30, 1])
    fill(c2, 0, [(top, [1])], [1] * 50 + [3, 33, 1, 65])
    fill(g1, 0, [(c1, [5]), (top, [0])], [1, 1, 2, 64, 1])
    names = ["u%d" % u for u in range(len(chains))]
    db = Database.create_hibf(bins, rows, h, k, w, nxt, pos, names, fpr=0.01)
    db.fill_random(9, 3)
    genomes = [bytes(rng.choice(list(b"ACGT"), size=700).astype(np.uint8)) for _ in names]
    per = {}
    for u, g in enumerate(genomes):
        hs = np.unique(O.minimiser_hash(g, k, w))
        for ibf, bs in chains[u]:
            a, b_ = per.setdefault(ibf, ([], []))
            a.append(hs)
            b_.append(np.asarray(bs, dtype=np.uint32)[np.arange(hs.size) % len(bs)])
    for ibf, (a, b_) in per.items():
        db.emplace(np.concatenate(a), np.concatenate(b_), ibf_index=ibf)
    path = str(tmp_path / "made.hibf")
    db.save(path)
    f = formats.read_hibf(path)
    assert [i.bins for i in f.ibfs] == bins and f.next_ibf_id == nxt and f.bin_to_user == pos
    assert [formats.hibf_target_name(p[0]) for p in f.bin_path] == names
    for i, ib in enumerate(f.ibfs):
        assert np.array_equal(ib.data, db.read_words(0, ib.bin_size * ib.bin_words, ibf_index=i))
    ohibf = O.OracleHIBF([O.OracleIBF(i.bins, i.bin_size, i.hash_funs, i.data) for i in f.ibfs], nxt, pos, len(names))
    reads = []
    for i in range(500):
        if i % 5 == 4:
            s = bytes(rng.choice(list(b"ACGT"), size=150).astype(np.uint8))
        else:
            g = genomes[int(rng.integers(0, len(names)))]
            p0 = int(rng.integers(0, len(g) - 150))
            s = bytearray(g[p0 : p0 + 150])
            for _ in range(int(rng.integers(0, 4))):
                s[int(rng.integers(0, 150))] = ord("ACGT"[int(rng.integers(0, 4))])
            s = bytes(s)
        reads.append((b"r%d" % i, s, None))
    fq = b"".join(b"@%s\n%s\n+\n%s\n" % (i, a, b"F" * len(a)) for i, a, _ in reads)
    reopened = Database.open(path, hibf=True)
    for cutoff, relf, fpq in [(0.05, 1.0, 1.0), (0.4, 0.3, 1e-4)]:
        filt = O.OracleFilter(ohibf, names, [[u] for u in range(len(names))], [0.01] * len(names), cutoff, k, w)
        want = O.all_lines(O.classify_level([filt], reads, relf, fpq))
        assert len(want) > 300
        for d in (db, reopened):
            sess = Session([d], [cutoff], [relf], [fpq], output_all=True)
            res = sess.classify(fq, final=True)
            assert sorted(result_text(res, "all").decode().splitlines()) == want, (cutoff, relf, fpq)
            sess.close()
    db.close()
    reopened.close()


# ------------------------------------------------------------------------------------------------------------------ build drop-in
def test_build_dropin_on_gpu_writes_the_oracle_backend_file(tmp_path):
    """`ganon-build` drop-in with K2 and the insertion on the device against the same orchestration with the oracle
    standing in for the device (tests/test_build_cpu.py pins that one to the reference builder): identical files."""
    import json

    from ganon_b200 import build as B
    from tests.build_util import OracleBackend

    cases = [c for c in json.load(open(os.path.join(SU.GOLDEN, "build_cases.json"))) if c["genomes"]]
    for ci, c in enumerate(cases[:6]):
        d = tmp_path / ("c%d" % ci)
        d.mkdir()
        tsv = str(d / "in.tsv")
        with open(tsv, "w") as t:
            for name, seq in c["genomes"].items():
                p = str(d / (name + ".fa"))
                open(p, "w").write(">%s\n%s\n" % (name, seq))
                t.write("%s\t%s\n" % (p, name))
        pr = c["params"]
        mk = lambda out: B.GanonBuildConfig(input_file=tsv, output_file=out, kmer_size=pr["k"], window_size=pr["w"], max_fp=pr["max_fp"], filter_size=pr["filter_size"],
                                            hash_functions=pr["hash_functions"], mode=pr["mode"], quiet=True)
        # the device leg through the command line, in its own process
        import subprocess
        import sys

        argv = [sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin", "ganon-build"), "-i", tsv, "-o", str(d / "gpu.ibf"), "-k", str(pr["k"]),
                "-w", str(pr["w"]), "-s", str(pr["hash_functions"]), "-j", pr["mode"], "--quiet"] + (["-f", str(pr["filter_size"])] if pr["filter_size"] else ["-p", str(pr["max_fp"])])
        done = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert done.returncode == 0, done.stderr[-500:]
        assert B.run_build(mk(str(d / "cpu.ibf")), backend=OracleBackend())
        assert open(str(d / "gpu.ibf"), "rb").read() == open(str(d / "cpu.ibf"), "rb").read()


# ------------------------------------------------------------------------------------------------------------------ pipeline + shards
def test_async_submit_collect_matches_sync(golden_dbs):
    fq1 = open(os.path.join(SU.GOLDEN, "reads.1.fq"), "rb").read()
    fq2 = open(os.path.join(SU.GOLDEN, "reads.2.fq"), "rb").read()
    recs1 = [b"@" + r for r in fq1.split(b"\n@")]
    recs1[0] = recs1[0][1:]
    recs2 = [b"@" + r for r in fq2.split(b"\n@")]
    recs2[0] = recs2[0][1:]
    fix = lambda r: r if r.endswith(b"\n") else r + b"\n"
    recs1, recs2 = [fix(r) for r in recs1], [fix(r) for r in recs2]
    db = Database.open(golden_dbs["synth"])
    ref = Session([db], [0.0], [1.0], [1.0], output_all=True, output_unclassified=True)
    r = ref.classify(b"".join(recs1), b"".join(recs2), final=True)
    want_all, want_unc = sorted(result_text(r, "all").decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines())
    want_rep = ref.report()
    sess = Session([db], [0.0], [1.0], [1.0], output_all=True, output_unclassified=True)
    _n, cap = sess.in_flight()
    assert cap >= 2
    blocks, step = [], 90
    for a in range(0, len(recs1), step):
        blocks.append((b"".join(recs1[a : a + step]), b"".join(recs2[a : a + step])))
    got_all, got_unc, pending = [], [], 0
    for i, (b1, b2) in enumerate(blocks):
        info = sess.submit(b1, b2, final=i == len(blocks) - 1)
        assert info.n_reads == min(step, len(recs1) - i * step) and info.consumed1 == len(b1) and info.consumed2 == len(b2)
        pending += 1
        if pending == cap:
            res = sess.collect()
            got_all += result_text(res, "all").decode().splitlines()
            got_unc += result_text(res, "unc").decode().splitlines()
            pending -= 1
    while pending:
        res = sess.collect()
        got_all += result_text(res, "all").decode().splitlines()
        got_unc += result_text(res, "unc").decode().splitlines()
        pending -= 1
    assert sorted(got_all) == want_all and sorted(got_unc) == want_unc
    assert sess.report() == want_rep
    with pytest.raises(_lib.GnbError):
        sess.collect()


def test_column_shards_on_one_gpu_equal_the_whole(golden_dbs):
    """Three shard handles of the same .ibf (targets of 1-3 bins straddle the shard borders): merged tuples give the
    unsharded result."""
    from ganon_b200.sharded import merge_tuples

    fq = open(os.path.join(SU.GOLDEN, "reads.se.fq"), "rb").read()
    whole = Session([Database.open(golden_dbs["synth"])], [0.1], [0.5], [1.0], output_all=True, output_unclassified=True)
    r = whole.classify(fq, final=True)
    want = (sorted(result_text(r, "all").decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), whole.report())
    n_shards = 3
    sessions = [Session([Database.open(golden_dbs["synth"], shard=i, n_shards=n_shards)], [0.1], [0.5], [1.0], output_all=True, output_unclassified=True) for i in range(n_shards)]
    for s in sessions:
        assert s.stage(fq, final=True) == r.n_reads
        s.run_level(0)
    parts = [s.level_tuples(0, 0) for s in sessions]
    assert any(((p >> np.uint64(16)) & np.uint64(1)).any() for p in parts)  # partial sums exist
    merged = merge_tuples(np.concatenate(parts))
    for s in sessions:
        s.set_level_tuples(0, 0, merged)
        s.finish_level(0)
        res = s.collect_staged()
        got = (sorted(result_text(res, "all").decode().splitlines()), sorted(result_text(res, "unc").decode().splitlines()), s.report())
        assert got == want


def test_column_shards_device_exchange_equals_the_whole(golden_dbs):
    """The same three shards with the exchange inside HBM: device tuples concatenated with torch (what the NCCL
    all-gather yields), sorted by the library, finished by K4 on every shard session."""
    import torch

    from ganon_b200.sharded import _DeviceWords

    fq = open(os.path.join(SU.GOLDEN, "reads.se.fq"), "rb").read()
    whole = Session([Database.open(golden_dbs["synth"])], [0.1], [0.5], [1e-3], output_all=True, output_unclassified=True)
    r = whole.classify(fq, final=True)
    want = (sorted(result_text(r, "all").decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), whole.report())
    n_shards = 3
    sessions = [Session([Database.open(golden_dbs["synth"], shard=i, n_shards=n_shards)], [0.1], [0.5], [1e-3], output_all=True, output_unclassified=True) for i in range(n_shards)]
    parts = []
    for s in sessions:
        assert s.stage(fq, final=True) == r.n_reads
        s.run_level_device(0)
        ptr, n = s.level_tuples_device(0)
        parts.append(torch.as_tensor(_DeviceWords(ptr, n), device="cuda").clone() if n else torch.empty(0, dtype=torch.int64, device="cuda"))
    merged = torch.cat(parts)
    torch.cuda.synchronize()
    for s in sessions:
        s.set_level_tuples_device(0, merged.data_ptr(), merged.numel())
        s.finish_level_device(0)
        res = s.collect_staged()
        assert res.levels_on_device == 1
        got = (sorted(result_text(res, "all").decode().splitlines()), sorted(result_text(res, "unc").decode().splitlines()), s.report())
        assert got == want


def test_sharded_create_fill_emplace_equals_slice():
    full = Database.create(1000, 997, 4, 19, 31)
    full.fill_random(3, 2)
    hs = np.arange(5000, 5400, dtype=np.uint64)
    bs = (np.arange(400) * 7 % 1000).astype(np.uint32)
    full.emplace(hs, bs)
    fi = full.info()
    whole = full.read_words(0, fi.bin_size_bits * fi.bin_words).reshape(fi.bin_size_bits, fi.bin_words)
    for sh in range(3):
        part = Database.create(1000, 997, 4, 19, 31, shard=sh, n_shards=3)
        part.fill_random(3, 2)
        part.emplace(hs, bs)
        pi = part.info()
        w = pi.shard_word_end - pi.shard_word_begin
        got = part.read_words(0, pi.bin_size_bits * w).reshape(pi.bin_size_bits, w)
        assert np.array_equal(got, whole[:, pi.shard_word_begin : pi.shard_word_end])


def _sharded_cases(golden_dbs):
    """(name, database paths, labels, cutoffs, rel_filter, fpr_query, reads1, reads2) run sharded and unsharded."""
    g = SU.GOLDEN
    return [
        ("se", [golden_dbs["synth"]], None, [0.1], [0.5], [1.0], os.path.join(g, "reads.se.fq"), None),
        ("pe_fpr", [golden_dbs["synth"]], None, [0.0], [1.0], [1e-3], os.path.join(g, "reads.1.fq"), os.path.join(g, "reads.2.fq")),
        ("fasta_host_reader", [golden_dbs["synth"]], None, [0.1], [0.5], [1.0], os.path.join(g, "reads.fa"), None),
        # (a column shard cannot be narrower than one 64-bin word, so the one-word real4 fixtures cannot be split)
        ("two_levels", [golden_dbs["synth"], golden_dbs["synth"]], ["A", "B"], [0.6, 0.1], [0.2, 0.5], [1.0, 1.0], os.path.join(g, "reads.se.fq"), None),
        ("two_filters_one_level", [golden_dbs["synth"], golden_dbs["synth"]], ["A", "A"], [0.3, 0.1], [0.5], [1.0], os.path.join(g, "reads.se.fq"), None),
    ]


def _cut_blocks(fq1, fq2, n_blocks):
    """Whole FASTQ / FASTA records grouped into n_blocks consecutive blocks (mates cut at the same record)."""
    def recs(b):
        mark = b[:1]
        out = [mark + r for r in b.split(b"\n" + mark)]
        out[0] = out[0][1:]
        return [r if r.endswith(b"\n") else r + b"\n" for r in out]

    r1 = recs(fq1)
    r2 = recs(fq2) if fq2 is not None else None
    step = (len(r1) + n_blocks - 1) // n_blocks
    return [(b"".join(r1[a : a + step]), b"".join(r2[a : a + step]) if r2 is not None else None) for a in range(0, len(r1), step)]


def _run_case(make_session, fq1, fq2):
    """The three forms of the public call on one session factory: synchronous classify, stage + run + finish, and the
    pipelined submit / collect over several blocks.  Returns their (.all lines sorted, .unc lines sorted, report)."""
    out = []
    s = make_session()
    r = s.classify(fq1, fq2, final=True)
    out.append((sorted(b"".join(result_text(r, "all", lv) for lv in range(len(s.level_labels))).decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), s.report()))
    s.close()
    s = make_session()
    s.stage(fq1, fq2, final=True)
    s.run_staged()
    r = s.finish_staged()
    out.append((sorted(b"".join(result_text(r, "all", lv) for lv in range(len(s.level_labels))).decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), s.report()))
    s.close()
    s = make_session()
    _n, cap = s.in_flight()
    blocks = _cut_blocks(fq1, fq2, 5)
    alls, uncs, pending = [], [], 0
    t1 = t2 = b""
    keep = []  # submitted blocks stay alive until collected
    for i, (b1, b2) in enumerate(blocks):
        # what a block did not consume goes in front of the next one (a FASTA record only ends where the next begins)
        b1 = t1 + b1
        b2 = t2 + b2 if b2 is not None else None
        keep.append((b1, b2))
        info = s.submit(b1, b2, final=i == len(blocks) - 1)
        t1 = b1[info.consumed1 :]
        t2 = b2[info.consumed2 :] if b2 is not None else b""
        pending += 1
        while pending >= cap or (i == len(blocks) - 1 and pending):
            r = s.collect()
            alls += b"".join(result_text(r, "all", lv) for lv in range(len(s.level_labels))).decode().splitlines()
            uncs += result_text(r, "unc").decode().splitlines()
            pending -= 1
    out.append((sorted(alls), sorted(uncs), s.report()))
    s.close()
    return out


def _two_gpu_worker(rank, world, port, cases, out_dir):
    import pickle

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ganon_b200.sharded import ShardedSession, make_comm

    comm = make_comm(rank)
    assert comm.n_ranks == world and comm.nccl_version() > 20000
    results = {}
    for name, paths, labels, cutoff, rel_filter, fpr, p1, p2 in cases:
        fq1 = open(p1, "rb").read()
        fq2 = open(p2, "rb").read() if p2 else None
        dbs = [Database.open(p, device=rank, shard=rank, n_shards=world) for p in paths]
        for sliced in (False, True):
            mk = lambda: ShardedSession(dbs, cutoff, rel_filter, fpr, hierarchy_labels=labels, output_all=True, output_unclassified=True, device=rank, comm=comm, sliced_ingest=sliced)
            results[(name, sliced)] = _run_case(mk, fq1, fq2)
        for d in dbs:
            d.close()
    with open(os.path.join(out_dir, "r%d.pkl" % rank), "wb") as f:
        pickle.dump(results, f)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_sharded_two_gpus_nccl(golden_dbs, tmp_path):
    """Bin-sharded sessions over 2 GPUs with the exchange inside the library (NCCL): every form of the public call, with
    and without sliced ingest, single / paired / FASTA (host reader) input, two hierarchy levels, two filters on one level
    (host finishing stage) -- each rank's result equals the unsharded session's."""
    import pickle

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    cases = _sharded_cases(golden_dbs)
    want = {}
    for name, paths, labels, cutoff, rel_filter, fpr, p1, p2 in cases:
        fq1 = open(p1, "rb").read()
        fq2 = open(p2, "rb").read() if p2 else None
        dbs = [Database.open(p) for p in paths]
        want[name] = _run_case(lambda: Session(dbs, cutoff, rel_filter, fpr, hierarchy_labels=labels, output_all=True, output_unclassified=True), fq1, fq2)
        assert want[name][0][0], name  # the case classifies something
    mp.spawn(_two_gpu_worker, args=(2, port, cases, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = pickle.load(open(tmp_path / ("r%d.pkl" % rank), "rb"))
        for (name, sliced), forms in got.items():
            for form, res in zip(("classify", "staged", "pipelined"), forms):
                assert res == want[name][0], (rank, name, sliced, form)
    for name in want:
        assert want[name][0] == want[name][1] == want[name][2], name  # the three forms agree unsharded as well


@pytest.mark.parametrize("name", ["se_synth_all", "hier_two_levels_single"])
def test_command_line_devices_two_gpus(name, golden_dbs, tmp_path):
    """`ganon-classify --devices 0,1`: one process per GPU on column shards of every .ibf (sliced reads of the plain and the
    gzip-compressed input, exchange inside the library); rank 0 writes the reference's outputs, bit for bit as one GPU does."""
    import gzip as gz
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    args = SU.expand(SU.load_scenarios()[name], golden_dbs)
    if any(a.endswith("real4.ibf") or a.endswith("real4b.ibf") or "real4.ibf," in a or "real4b.ibf," in a for a in args):
        # the one-word fixtures cannot be split into two column shards: use the 3-word fixture in their place
        args = [a.replace(golden_dbs["real4b"], golden_dbs["synth"]).replace(golden_dbs["real4"], golden_dbs["synth"]) for a in args]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    one = str(tmp_path / "one")
    assert cli.main(args + ["-o", one, "-t", "4", "--quiet"]) == 0
    for tag, conv in (("two", lambda a: a), ("two_gz", None)):
        a2 = list(args)
        if conv is None:  # the same reads gzip-compressed
            i = a2.index("-r")
            files = []
            for f in a2[i + 1].split(","):
                g = str(tmp_path / (os.path.basename(f) + ".gz"))
                with open(f, "rb") as fi, gz.open(g, "wb") as fo:
                    fo.write(fi.read())
                files.append(g)
            a2[i + 1] = ",".join(files)
        pre = str(tmp_path / tag)
        done = subprocess.run([sys.executable, os.path.join(root, "bin", "ganon-classify")] + a2 + ["-o", pre, "-t", "4", "--quiet", "--devices", "0,1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert done.returncode == 0, done.stderr[-2000:]
        want_files = sorted(os.path.basename(p)[len("one") :] for p in glob.glob(one + ".*"))
        assert sorted(os.path.basename(p)[len(tag) :] for p in glob.glob(pre + ".*")) == want_files
        for ext in want_files:
            assert _read_sorted(pre + ext) == _read_sorted(one + ext), (tag, ext)


def test_paged_filter_host_resident_tier_equals_the_whole(golden_dbs, tmp_path, monkeypatch):
    """a filter above its HBM budget is cut into column pages, some resident, the others streamed from
    page-locked host memory through two staging buffers; every form of the call gives the unpaged result bit for bit
    (targets of 1-3 bins straddle the page borders of the 3-word fixture)."""
    fq1 = open(os.path.join(SU.GOLDEN, "reads.1.fq"), "rb").read()
    fq2 = open(os.path.join(SU.GOLDEN, "reads.2.fq"), "rb").read()
    path = golden_dbs["synth"]
    whole = Database.open(path)
    wi = whole.info()
    assert wi.n_pages == 0
    col = wi.bin_size_bits * 8  # bytes of one bin-word column
    mk = lambda db: (lambda: Session([db], [0.1], [0.5], [1e-3], output_all=True, output_unclassified=True))
    want = _run_case(mk(whole), fq1, fq2)
    assert want[0] == want[1] == want[2] and want[0][0]
    for budget, n_res in ((2 * col + 100, 0), (5 * col // 2, 0)):  # (three one-word pages: the two staging buffers fill the budget)
        paged = Database.open(path, hbm_budget=budget)
        pi = paged.info()
        assert (pi.n_pages, pi.n_resident_pages) == (3, n_res) and pi.host_bytes == (3 - n_res) * col and pi.device_bytes == (n_res + 2) * col
        got = _run_case(mk(paged), fq1, fq2)
        assert got == want, budget
        with pytest.raises(_lib.GnbError):
            paged.read_words(0, 8)
        paged.close()
    # a filter made in HBM, then paged out: 128 words per row in 11 pages of 12 words (6 resident, 5 streamed), a 12-bin
    # target across the border of two pages and a background target spread over all of them
    rng = np.random.default_rng(9)
    k, w = 19, 31
    db = Database.create(8192, 4099, 3, k, w)
    db.fill_random(4, 2)
    genomes = [bytes(rng.choice(list(b"ACGT"), size=2000).astype(np.uint8)) for _ in range(64)]
    first = [int(x) for x in rng.choice(8100, size=64, replace=False)]
    first[0] = 3835  # bins 3835..3846 straddle word 59 | 60, the border of pages 4 and 5
    bin_target = np.full(8192, 64, dtype=np.uint32)  # the rest: one background target
    names = ["g%d" % i for i in range(64)] + ["bg"]
    hs, bs, counts = [], [], []
    for t, g in enumerate(genomes):
        nb = 12 if t == 0 else 1
        bin_target[first[t] : first[t] + nb] = t
        u = np.unique(O.minimiser_hash(g, k, w))
        hs.append(u)
        bs.append((first[t] + np.arange(u.size) % nb).astype(np.uint32))
        counts.append(u.size)
    db.emplace(np.concatenate(hs), np.concatenate(bs))
    db.set_targets(names, bin_target, counts + [1000], 500)
    reads = b"".join(b"@q%d\n%s\n+\n%s\n" % (i, genomes[i % 64][100 + i : 250 + i], b"I" * 150) for i in range(640))
    mk2 = lambda: Session([db], [0.2], [0.5], [1.0], output_all=True, output_unclassified=True)
    want2 = _run_case(mk2, reads, None)
    assert want2[0][0]
    db.page_out(100 * 4099 * 8)  # room for 100 of the 128 columns: pages of 100 / 8 = 12 words, 8 fit, 2 of them staging
    di = db.info()
    assert (di.n_pages, di.n_resident_pages) == (11, 6)
    assert _run_case(mk2, reads, None) == want2
    # the command line takes the budget from the environment
    monkeypatch.setenv("GANON_B200_HBM_BUDGET_GB", "%.9f" % ((2 * col + 100) / (1 << 30)))
    pre = str(tmp_path / "paged")
    assert cli.main(["-r", os.path.join(SU.GOLDEN, "reads.se.fq"), "-i", path, "-c", "0.1", "-d", "0.5", "-o", pre, "-a", "-u", "--quiet"]) == 0
    monkeypatch.delenv("GANON_B200_HBM_BUDGET_GB")
    pre3 = str(tmp_path / "paged_flag")  # the same through the command-line flag
    assert cli.main(["-r", os.path.join(SU.GOLDEN, "reads.se.fq"), "-i", path, "-c", "0.1", "-d", "0.5", "-o", pre3, "-a", "-u", "--quiet", "--hbm-budget-gb", "%.9f" % ((2 * col + 100) / (1 << 30))]) == 0
    pre2 = str(tmp_path / "whole")
    assert cli.main(["-r", os.path.join(SU.GOLDEN, "reads.se.fq"), "-i", path, "-c", "0.1", "-d", "0.5", "-o", pre2, "-a", "-u", "--quiet"]) == 0
    for ext in (".all", ".unc", ".rep"):
        assert _read_sorted(pre + ext) == _read_sorted(pre2 + ext)
        assert _read_sorted(pre3 + ext) == _read_sorted(pre2 + ext)


def test_levels_with_several_filters_finish_on_the_device(golden_dbs):
    """The cross-filter merge of select_matches (GC.cpp:528-539: store only a strictly greater count, max / min updated at
    every store -- the stale-min behaviour) runs in K4: two and three databases on one level report levels_on_device == 1
    and give the host finishing stage's result (which the golden scenarios pin to the reference binary)."""
    fq1 = open(os.path.join(SU.GOLDEN, "reads.1.fq"), "rb").read()
    fq2 = open(os.path.join(SU.GOLDEN, "reads.2.fq"), "rb").read()
    for names, cutoffs, rel_filter, fpr in ((("real4", "real4b"), [0.1, 0.3], [0.2], [1.0]), (("real4b", "synth", "real4"), [0.05] * 3, [0.3], [0.5]), (("synth", "synth"), [0.3, 0.1], [0.5], [1e-3])):
        dbs = [Database.open(golden_dbs[n]) for n in names]
        out = {}
        for mode in ("device", "host"):
            if mode == "host":
                os.environ["GANON_B200_HOST_FINISH"] = "1"
            try:
                s = Session(dbs, cutoffs, rel_filter, fpr, output_all=True, output_unclassified=True, output_lca=False)
                r = s.classify(fq1, fq2, final=True)
                out[mode] = (sorted(result_text(r, "all").decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), s.report(), r.levels_on_device)
                s.close()
            finally:
                os.environ.pop("GANON_B200_HOST_FINISH", None)
        assert out["device"][3] == 1 and out["host"][3] == 0, names
        assert out["device"][:3] == out["host"][:3] and out["device"][0], names


def test_build_file_hashes_equal_the_oracle_minimiser_sets(tmp_path):
    """gnb_build_file_hashes (ganon-build's count_hashes for one file): sequences cut into segments of 2048 windows, K2, sort
    + unique in HBM -- the distinct minimisers equal the set the oracle finds over the whole sequences; wrapped FASTA, gzip,
    --min-length, sequences shorter than the window (clamped window) and shorter than k."""
    import gzip as gz

    from ganon_b200.classify import build_file_hashes

    rng = np.random.default_rng(21)
    k, w = 19, 31
    seqs = [bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8)) for n in (50_000, 7, 25, 30, 31, 2048 + 30, 2049 + 30, 4096 + 31, 300, 12_345)]
    seqs[8] = b"ACGTN" * 60  # IUPAC letters count as A
    fa = b"".join(b">s%d some text\n" % i + b"\n".join(s[a : a + 70] for a in range(0, len(s), 70)) + b"\n" for i, s in enumerate(seqs))
    (tmp_path / "g.fa").write_bytes(fa)
    with gz.open(tmp_path / "g.fa.gz", "wb") as f:
        f.write(fa)
    fq = b"".join(b"@q%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs))
    (tmp_path / "g.fq").write_bytes(fq)

    def oracle_set(min_length):
        parts = []
        for s in seqs:
            if len(s) < min_length or len(s) < k:
                continue
            parts.append(O.minimiser_hash(s, k, min(w, len(s))))
        return np.unique(np.concatenate(parts))

    for name in ("g.fa", "g.fa.gz", "g.fq"):
        for min_length in (0, 26, 301):
            got, st = build_file_hashes(str(tmp_path / name), k, w, min_length)
            want = oracle_set(min_length)
            assert np.array_equal(got, want), (name, min_length, got.size, want.size)
            assert st.n_sequences == sum(len(s) >= min_length for s in seqs) and st.n_skipped == sum(len(s) < min_length for s in seqs)
            assert st.n_bases == sum(len(s) for s in seqs if len(s) >= min_length) and st.n_unique == want.size and not st.parse_error
    # a parse error drops the file's hashes, the counts stay
    (tmp_path / "bad.fq").write_bytes(fq + b"@broken\nACGT\n+\nII\n")
    got, st = build_file_hashes(str(tmp_path / "bad.fq"), k, w, 0)
    assert got is None and st.parse_error and st.n_sequences == len(seqs)


@pytest.mark.parametrize("bad_at", [2801, 2912, 2913, 5000, 5601, 9999])
def test_parse_error_chunk_rule_across_blocks(golden_dbs, tmp_path, monkeypatch, bad_at):
    """The file is read in blocks (here 256 KiB ~ 2900 records); a parse error in a later block must still retract exactly
    the reference's chunks: a block that does not end the file hands on floor((end - 1) / n_reads) * n_reads records and
    keeps the rest for the next block, so nothing already classified has to be taken back."""
    import ganon_b200.classify as K

    monkeypatch.setattr(K, "BLOCK_BYTES", 256 << 10)
    recs = [b"@r%05d\n%s\n+\n%s\n" % (i, b"ACGT" * 10, b"I" * 40) for i in range(10000)]
    recs[bad_at] = b"@bad\nACGTXXXX\n+\nIIIIIIII\n"
    f = str(tmp_path / "bad.fq")
    open(f, "wb").write(b"".join(recs))
    pre = str(tmp_path / "o")
    assert cli.main(["-r", f, "-i", golden_dbs["synth"], "-o", pre, "-u", "--quiet", "--n-reads", "400"]) == 0
    rep = dict(l.split("\t") for l in open(pre + ".rep").read().splitlines() if l.startswith("#"))
    assert int(rep.get("#total_unclassified", 0)) + int(rep.get("#total_classified", 0)) == (bad_at - 1) // 400 * 400
    assert len(open(pre + ".unc").read().splitlines()) == int(rep.get("#total_unclassified", 0))


def test_file_format_follows_the_file_name_like_seqan3(golden_dbs, tmp_path):
    """seqan3 picks the reader from the extension (compression suffix stripped): FASTQ content in a .fa file is a parse error
    on the first record (nothing classified, the run goes on), an unknown extension ends the reference's run -- here an error."""
    fq = open(os.path.join(SU.GOLDEN, "reads.se.fq"), "rb").read()
    ok, wrong, unknown = str(tmp_path / "r.fastq"), str(tmp_path / "r.fa"), str(tmp_path / "r.txt")
    for p in (ok, wrong, unknown):
        open(p, "wb").write(fq)
    pre = str(tmp_path / "o")
    assert cli.main(["-r", ok, "-i", golden_dbs["synth"], "-o", pre, "-u", "--quiet"]) == 0
    n_ok = len(open(pre + ".rep").read().splitlines())
    assert n_ok > 2
    assert cli.main(["-r", wrong + "," + ok, "-i", golden_dbs["synth"], "-o", pre + "2", "-u", "--quiet"]) == 0
    assert open(pre + "2.rep").read() == open(pre + ".rep").read()  # the mis-named file contributes nothing, the next one is read
    assert cli.main(["-r", unknown, "-i", golden_dbs["synth"], "-o", pre + "3", "-u", "--quiet"]) != 0


def test_embl_genbank_and_sam_read_files_like_seqan3(golden_dbs, tmp_path):
    """seqan3::sequence_file_input also reads EMBL, GenBank and SAM, chosen by the file name (csrc/seqformats.cpp rewrites them
    as two-line FASTA for K1): the same records in every format give the same outputs as the FASTQ file -- GenBank ids are the
    whole LOCUS line, as the reference returns them -- and a record seqan3 throws on loses the reference's chunk of --n-reads."""
    from tests import test_seqformats_cpu as SF
    import random

    fq = open(os.path.join(SU.GOLDEN, "reads.se.fq"), "rb").read().split(b"\n")
    recs = [(fq[i][1:], fq[i + 1]) for i in range(0, len(fq) - 1, 4)]
    pre = str(tmp_path / "fq")
    assert cli.main(["-r", os.path.join(SU.GOLDEN, "reads.se.fq"), "-i", golden_dbs["synth"], "-o", pre, "-a", "-u", "--quiet"]) == 0
    want_all = sorted(open(pre + ".all", "rb").read().splitlines())
    want_unc = sorted(open(pre + ".unc", "rb").read().splitlines())
    assert want_all
    rng = random.Random(2)
    for fmt, ext in (("sam", "sam"), ("embl", "embl"), ("genbank", "gbk"), ("sam", "sam.gz")):
        data = SF.dress(rng, recs, fmt, {"header": True, "tags": True, "minimal": fmt == "genbank"})
        f = str(tmp_path / ("reads." + ext))
        open(f, "wb").write(gzip.compress(data) if ext.endswith(".gz") else data)
        out = str(tmp_path / ("o_" + ext))
        assert cli.main(["-r", f, "-i", golden_dbs["synth"], "-o", out, "-a", "-u", "--quiet"]) == 0
        strip = (lambda l: l) if fmt != "genbank" else (lambda l: b"\t".join([l.split(b"\t")[0].split(b" ")[0]] + l.split(b"\t")[1:]))
        assert sorted(strip(l) for l in open(out + ".all", "rb").read().splitlines()) == want_all, fmt
        assert sorted(strip(l) for l in open(out + ".unc", "rb").read().splitlines()) == want_unc, fmt
        assert open(out + ".rep").read() == open(pre + ".rep").read(), fmt
    # a record without a sequence in a SAM file: parse error at record 450 of --n-reads 400 -> 400 records kept
    bad = list(recs[:1000]) if len(recs) >= 1000 else [recs[i % len(recs)] for i in range(1000)]
    bad = [(b"r%d" % i, s) for i, (_id, s) in enumerate(bad)]
    bad[450] = (b"bad", b"")
    f = str(tmp_path / "bad.sam")
    open(f, "wb").write(SF.dress(rng, bad, "sam", {}))
    out = str(tmp_path / "o_bad")
    assert cli.main(["-r", f, "-i", golden_dbs["synth"], "-o", out, "-u", "--quiet", "--n-reads", "400"]) == 0
    rep = dict(l.split("\t") for l in open(out + ".rep").read().splitlines() if l.startswith("#"))
    assert int(rep.get("#total_unclassified", 0)) + int(rep.get("#total_classified", 0)) == 400


def test_bzip2_read_files(golden_dbs, tmp_path):
    """`.bz2` read files (the reference reads them through seqan3 + libbz2; csrc/bz2stream.cpp decodes the blocks on all host
    threads): same outputs as the plain files, single-end and paired, one stream and several."""
    import bz2

    se = os.path.join(SU.GOLDEN, "reads.se.fq")
    pre = str(tmp_path / "plain")
    assert cli.main(["-r", se, "-i", golden_dbs["synth"], "-o", pre, "-a", "-u", "--quiet"]) == 0
    data = open(se, "rb").read()
    half = data.index(b"\n@", len(data) // 2) + 1
    z = str(tmp_path / "reads.se.fq.bz2")
    open(z, "wb").write(bz2.compress(data[:half], 1) + bz2.compress(data[half:], 9))
    out = str(tmp_path / "bz")
    assert cli.main(["-r", z, "-i", golden_dbs["synth"], "-o", out, "-a", "-u", "--quiet"]) == 0
    for ext in (".all", ".unc", ".rep"):
        assert open(out + ext).read() == open(pre + ext).read(), ext
    p1, p2 = os.path.join(SU.GOLDEN, "reads.1.fq"), os.path.join(SU.GOLDEN, "reads.2.fq")
    assert cli.main(["-p", p1 + "," + p2, "-i", golden_dbs["synth"], "-o", pre + "_pe", "-a", "-u", "--quiet"]) == 0
    zs = []
    for f in (p1, p2):
        zs.append(str(tmp_path / (os.path.basename(f) + ".bz2")))
        open(zs[-1], "wb").write(bz2.compress(open(f, "rb").read(), 1))
    assert cli.main(["-p", ",".join(zs), "-i", golden_dbs["synth"], "-o", out + "_pe", "-a", "-u", "--quiet"]) == 0
    for ext in (".all", ".unc", ".rep"):
        assert open(out + "_pe" + ext).read() == open(pre + "_pe" + ext).read(), ext


def test_long_reads_finish_on_the_device_when_fpr_query_is_off(golden_dbs):
    """Reads of tens of thousands of minimisers (long-read data) stay on K4 unless --fpr-query needs the device's libm-style
    evaluation (trusted up to 4096 minimisers): same result as the host stage; with --fpr-query the level goes to the host."""
    rng = np.random.default_rng(31)
    reads = b"".join(b"@long%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8)), b"I" * n) for i, n in enumerate((60_000, 150, 120_000, 9_000)))
    db = Database.open(golden_dbs["synth"])
    out = {}
    for mode, fpr in (("device", 1.0), ("host", 1.0), ("device_fpr", 1e-3)):
        if mode == "host":
            os.environ["GANON_B200_HOST_FINISH"] = "1"
        try:
            s = Session([db], [0.0], [1.0], [fpr], output_all=True, output_unclassified=True)
            r = s.classify(reads, final=True)
            out[mode] = (sorted(result_text(r, "all").decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), s.report(), r.levels_on_device, max(r.n_hashes[i] for i in range(4)))
            s.close()
        finally:
            os.environ.pop("GANON_B200_HOST_FINISH", None)
    assert out["device"][4] > 10_000
    assert out["device"][3] == 1 and out["host"][3] == 0 and out["device_fpr"][3] == 0
    assert out["device"][:3] == out["host"][:3] and out["device"][0]


def test_unwrapped_fasta_is_indexed_on_the_device(golden_dbs, monkeypatch):
    """K1 takes 2-line FASTA records (">id\\nSEQ\\n"): same result as the host reader over blocks that end anywhere, with and
    without a final newline; a wrapped FASTA block is recognised (the line after a sequence line is not a header) and goes to
    the host reader."""
    fq = open(os.path.join(SU.GOLDEN, "reads.se.fq"), "rb").read().split(b"\n")
    recs = [(fq[i][1:], fq[i + 1]) for i in range(0, len(fq) - 1, 4)]
    flat = b"".join(b">%s\n%s\n" % (i, s) for i, s in recs)
    wrapped = b"".join(b">%s\n%s\n" % (i, b"\n".join(s[a : a + 60] for a in range(0, len(s), 60))) for i, s in recs)
    db = Database.open(golden_dbs["synth"])
    out = {}
    for name, text, host_index in (("flat_dev", flat, "0"), ("flat_dev_no_newline", flat.rstrip(b"\n"), "0"), ("flat_host", flat, "1"), ("wrapped", wrapped, "0")):
        monkeypatch.setenv("GANON_B200_HOST_INDEX", host_index)
        sess = Session([db], [0.0], [1.0], [1.0], output_all=True, output_unclassified=True)
        lines, uncs, pos, total, on_dev, flags = [], [], 0, 0, 0, []
        while True:
            end = min(len(text), pos + 9011)
            final = end == len(text)
            r = sess.classify(text[pos:end], final=final)
            lines += result_text(r, "all").decode().splitlines()
            uncs += result_text(r, "unc").decode().splitlines()
            total += r.n_reads
            on_dev += r.ms_index > 0
            flags.append(int(r.ms_index > 0))
            pos += r.consumed1
            if final:
                break
            assert r.n_reads > 0
        out[name] = (total, sorted(lines), sorted(uncs), sess.report())
        sess.close()
        if name.startswith("flat_dev"):
            assert on_dev >= len(flags) - 1 and on_dev > 3, (name, flags)  # the device index took the blocks
        else:
            assert on_dev == 0, name
    assert out["flat_dev"][0] == len(recs) and out["flat_dev"][1]
    assert out["flat_dev"] == out["flat_host"] == out["flat_dev_no_newline"] == out["wrapped"]


def test_hibf_levels_with_several_filters_finish_on_the_device(golden_dbs):
    """Two and three HIBFs on one hierarchy level (ordinary ganon usage: several --db-prefix): every traversal appends behind
    the tuples of the filters before it, K4 merges the filters of a node; same result as the host finishing stage."""
    fq = open(os.path.join(SU.GOLDEN, "reads.hibf.fq"), "rb").read()
    for cutoffs, rel_filter, fpr in (([0.3, 0.05], [0.5], [1.0]), ([0.0, 0.2, 0.6], [1.0], [1e-2])):
        dbs = [Database.open(golden_dbs["synth_hibf"], hibf=True) for _ in cutoffs]
        out = {}
        for mode in ("device", "host"):
            if mode == "host":
                os.environ["GANON_B200_HOST_FINISH"] = "1"
            try:
                s = Session(dbs, cutoffs, rel_filter, fpr, output_all=True, output_unclassified=True)
                r = s.classify(fq, final=True)
                out[mode] = (sorted(result_text(r, "all").decode().splitlines()), sorted(result_text(r, "unc").decode().splitlines()), s.report(), r.levels_on_device)
                s.close()
            finally:
                os.environ.pop("GANON_B200_HOST_FINISH", None)
        assert out["device"][3] == 1 and out["host"][3] == 0, cutoffs
        assert out["device"][:3] == out["host"][:3] and out["device"][0], cutoffs
