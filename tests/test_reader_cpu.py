"""The library's host record reader (ganon_b200/csrc/reads.cpp) on the CPU: the same source compiled with
g++ (tests/native/reads_host.cpp), checked on hand-written inputs, for independence of the block size, and -- where
oracle/_ref exists -- differentially against the UNMODIFIED reference binary on randomly formatted FASTA / FASTQ files
(wrapped lines, blanks, digits, CRLF, `;` headers, blanks before ids, missing final newline, illegal letters with the
chunk-loss rule of GanonClassify.cpp:1220-1287).  tests/reader_util.py holds the machinery; `python -m tests.reader_util N`
runs a long campaign (round 1: 2850 seeds, no mismatch; a dozen inputs made the reference itself crash or hang)."""
import os
import subprocess

import pytest

from tests import fuzz_util as F
from tests import reader_util as R

needs_ref = pytest.mark.skipif(not os.path.exists(F.REF_BIN), reason="oracle/_ref not built (only in the build container)")


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    lib = R.lib(str(tmp_path_factory.mktemp("reads_host")))
    if lib is None:
        pytest.skip("no g++")
    return lib


def test_plain_and_dressed_records(L):
    recs, consumed, err, _ = R.index_block(L, b"@r1 extra\nACGT\n+\nIIII\n@r2\nNNAC\n+r2\n!!!!\n", True)
    assert recs == [(b"r1 extra", b"ACGT"), (b"r2", b"NNAC")] and err is None and consumed == 40
    # wrapped FASTQ, blanks inside the sequence, qualities that start with '@' and '+'
    recs, _c, err, _ = R.index_block(L, b"@w\nAC GT\nAC\tGT\n+\n@+II\nIIII\n@x\nAC\n+\nII\n", True)
    assert recs == [(b"w", b"ACGTACGT"), (b"x", b"AC")] and err is None
    # FASTA: several lines, blank lines, digits and blanks in the sequence, blanks before the id, ';' header, no final newline
    recs, _c, err, _ = R.index_block(L, b">  a b\nAC\n\nGT 10 ac\n;c\nTTTT\n>d\nGG", True)
    assert recs == [(b"a b", b"ACGTac"), (b"c", b"TTTT"), (b"d", b"GG")] and err is None
    # every dna15 letter and U in both cases is legal, anything else is a parse error at that record
    recs, _c, err, _ = R.index_block(L, b">ok\nABCDGHKMNRSTVWYUabcdghkmnrstvwyu\n>bad\nACGTEACGT\n>after\nAC\n", True)
    assert recs[0] == (b"ok", b"ABCDGHKMNRSTVWYUabcdghkmnrstvwyu") and err == 1
    recs, _c, err, _ = R.index_block(L, b"@a\nAC\n+\nII\n@b\nAXC\n+\nIII\n", True)
    assert recs == [(b"a", b"AC")] and err == 1


def test_incomplete_records_wait_for_the_next_block(L):
    data = b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\nJJJJ\n"
    for cut in range(1, len(data)):
        recs, consumed, err, _ = R.index_block(L, data[:cut], False)
        assert err is None and consumed <= cut
        assert recs == [(b"r1", b"ACGT"), (b"r2", b"GGCC")][: len(recs)]
        assert data[:consumed].count(b"@r") == len(recs)
    fa = b">a\nACGT\nAC\n>b\nGG\n"
    for cut in range(1, len(fa)):
        recs, consumed, err, _ = R.index_block(L, fa[:cut], False)
        assert err is None and recs == [(b"a", b"ACGTAC")][: len(recs)]  # the last record may still grow


def test_block_size_does_not_matter(L):
    import random

    rng = random.Random(7)
    genomes = {"g": F._seq(rng, 3000)}
    for fmt in ("fasta", "fastq"):
        for trial in range(6):
            recs = F.make_reads(rng, genomes, 80, 31)
            style = {s: rng.random() < 0.4 for s in R.STYLES}
            data = R.dress(rng, recs, fmt, style)
            whole, e0 = R.read_file(L, data, 1 << 30, 400)
            # (wrapped quality lines with CRLF ends are a parse error in the reference too: "Qualitites longer than sequence.")
            assert e0 or len(whole) == len(recs), (fmt, trial, style)
            for bs in (37, 64, 1000, 4096):
                part, e1 = R.read_file(L, data, bs, 400)
                assert (part, e1) == (whole, e0), (fmt, trial, bs, style)


@needs_ref
@pytest.mark.parametrize("seed", range(24))
def test_reader_differential_against_reference_binary(L, seed, tmp_path):
    ok, desc = R.differential_case(L, seed, str(tmp_path), F.REF_BIN)
    if ok is None:
        pytest.skip(desc)
    assert ok, desc


@needs_ref
@pytest.mark.parametrize("bad_at", [0, 1, 2, 399, 400, 401, 799, 800, 801])
def test_parse_error_loses_the_chunk_of_the_previous_record(L, bad_at, tmp_path):
    """--n-reads 400: record e is read while the chunk holding record e - 1 is still being assembled
    (seqan3::views::chunk looks one record ahead), so the reference keeps floor((e - 1) / 400) * 400 records."""
    import random

    rng = random.Random(1)
    ibf = str(tmp_path / "db.ibf")
    F.make_db(rng, ibf, 12, 16)
    recs = [(b"r%d" % i, b"ACGTTGCAAGCTTGCAATGC") for i in range(1000)]
    recs[bad_at] = (b"bad", b"ACGTXACGTACGTACGTACG")
    data = R.dress(rng, recs, "fastq", {})
    f = str(tmp_path / "bad.fq")
    open(f, "wb").write(data)
    pre = str(tmp_path / "o")
    pr = subprocess.run([F.REF_BIN, "-r", f, "-i", ibf, "-o", pre, "-u", "--quiet", "--n-reads", "400", "-t", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert pr.returncode == 0, pr.stderr
    rep = dict(l.split("\t") for l in open(pre + ".rep").read().splitlines() if l.startswith("#"))
    kept_ref = int(rep.get("#total_unclassified", 0)) + int(rep.get("#total_classified", 0))  # no totals at all without a single read
    mine, err = R.read_file(L, data, 1 << 16, 400)
    assert err and len(mine) == kept_ref == (max(bad_at - 1, 0) // 400 * 400 if bad_at else 0)
