"""seqan3's other sequence formats (EMBL, GenBank, SAM -- `seqan3::sequence_file_input` picks them by file name, SURVEY 8a
row A0) through the library's byte-stream reader (ganon_b200/csrc/seqformats.cpp behind `gnb_reads_file_*`, no GPU involved):
the stream comes out as two-line FASTA.  Checked on hand-written files and -- where oracle/_ref exists -- differentially
against the UNMODIFIED reference binary: the reference classifies the EMBL / GenBank / SAM file, the oracle classifies the
records the host record reader (tests/reader_util.py) finds in the rewritten stream, chunk-loss rule of a parse error
included (GanonClassify.cpp:1220-1287).  `python -m tests.test_seqformats_cpu N` runs a longer campaign."""
import ctypes as C
import gzip
import os
import random
import subprocess

import pytest

from tests import fuzz_util as F
from tests import reader_util as R

needs_ref = pytest.mark.skipif(not os.path.exists(F.REF_BIN), reason="oracle/_ref not built (only in the build container)")


def library_stream(path, piece=1 << 16):
    """The bytes `gnb_reads_file_read` returns for a file, read in pieces."""
    from ganon_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    assert L.gnb_reads_file_open(path.encode(), 2, C.byref(h)) == 0, L.gnb_last_error()
    buf = C.create_string_buffer(piece)
    out = []
    try:
        while True:
            n = L.gnb_reads_file_read(h, buf, piece)
            assert n >= 0, L.gnb_last_error()
            if n == 0:
                break
            out.append(buf.raw[:n])
    finally:
        L.gnb_reads_file_close(h)
    return b"".join(out)


def _grouped(s, eol, lower=True, numbers="front"):
    """Sequence lines of 60 letters in groups of ten, numbered in front (GenBank) or behind (EMBL)."""
    s = s.lower() if lower else s
    lines = []
    for o in range(0, len(s), 60):
        row = s[o : o + 60]
        body = b" ".join(row[j : j + 10] for j in range(0, len(row), 10))
        lines.append((b"%9d " % (o + 1) + body) if numbers == "front" else (b"     " + body + b" %9d" % (o + len(row))))
    return eol.join(lines) + eol if lines else b""


def dress(rng, recs, fmt, style):
    eol = b"\r\n" if style.get("crlf") else b"\n"
    out = []
    if fmt == "sam" and style.get("header"):
        out.append(b"@HD\tVN:1.6\tSO:unsorted" + eol + b"@PG\tID:x\tPN:x" + eol)
    for rid, s in recs:
        if fmt == "genbank":
            if style.get("minimal"):  # what seqan3 itself writes
                out.append(b"LOCUS       " + rid + b"                 %d bp" % len(s) + eol + b"ORIGIN" + eol)
            else:
                out.append(b"LOCUS       " + rid + b"   %d bp    DNA     linear   UNK 01-JAN-1980" % len(s) + eol + b"DEFINITION  a read." + eol + b"ACCESSION   " + rid + eol +
                           b"FEATURES             Location/Qualifiers" + eol + b"     source          1..%d" % len(s) + eol + b"ORIGIN" + eol)
            out.append(_grouped(s, eol, lower=not style.get("upper")))
            out.append(b"//" + eol)
            if style.get("blank_lines") and rng.random() < 0.3:
                out.append(eol)
        elif fmt == "embl":
            if style.get("minimal"):
                out.append(b"ID " + rid + b"; %d BP." % len(s) + eol + b"SQ Sequence %d BP;" % len(s) + eol)
            else:
                out.append(b"ID   " + rid + b"; SV 1; linear; genomic DNA; STD; UNC; %d BP." % len(s) + eol + b"XX" + eol + b"DE   a read" + eol + b"XX" + eol +
                           b"SQ   Sequence %d BP; 0 A; 0 C; 0 G; 0 T; 0 other;" % len(s) + eol)
            out.append(_grouped(s, eol, lower=not style.get("upper"), numbers="behind"))
            out.append(b"//" + eol)
        else:
            flag = rng.choice((b"4", b"77", b"141", b"0"))
            tags = b"\tNM:i:0\tXS:Z:a b" if style.get("tags") and rng.random() < 0.5 else b""
            seq = s if s else b"*"
            out.append(rid + b"\t" + flag + b"\t*\t0\t0\t*\t*\t0\t0\t" + seq + b"\t" + (b"I" * len(s) if s and rng.random() < 0.8 else b"*") + tags + eol)
    data = b"".join(out)
    if style.get("trailing_newline"):
        data += eol
    if style.get("no_final_newline") and fmt == "sam":
        data = data.rstrip(b"\r\n")
    return data


STYLES = ["crlf", "header", "minimal", "upper", "blank_lines", "tags"]
RARE = ["trailing_newline", "empty_seq"]
EXT = {"genbank": ("gb", "gbk", "genbank"), "embl": ("embl",), "sam": ("sam",)}


def differential_case(L, seed, tmp, ref_bin):
    """-> (ok, description); None = the reference binary itself died on the input (seqan3 throws exceptions other than
    parse_error on truncated records and malformed numbers; ganon-classify does not catch them)."""
    from ganon_b200 import formats
    from oracle import oracle as O

    rng = random.Random(seed)
    k = rng.choice((8, 12, 19))
    w = k + rng.choice((0, 4, 12))
    ibf = os.path.join(tmp, "f%d.ibf" % seed)
    genomes = F.make_db(rng, ibf, k, w)
    recs = F.make_reads(rng, genomes, rng.choice((1, 7, 60)), w)
    fmt = rng.choice(("genbank", "embl", "sam"))
    style = {s: rng.random() < 0.35 for s in STYLES}
    style.update({s: rng.random() < 0.12 for s in RARE})
    if style["empty_seq"] and recs:
        j = rng.randrange(len(recs))
        recs[j] = (recs[j][0], b"")
    bad_at = None
    if rng.random() < 0.35 and recs:
        bad_at = rng.randrange(len(recs))
        rid, s = recs[bad_at]
        p = rng.randrange(max(1, len(s)))
        recs[bad_at] = (rid, s[:p] + rng.choice((b"X", b"E", b"*", b"-", b"Z", b"@", b">")) + s[p + 1 :])
    data = dress(rng, recs, fmt, style)
    zipped = rng.random() < 0.25
    path = os.path.join(tmp, "f%d.%s%s" % (seed, rng.choice(EXT[fmt]), ".gz" if zipped else ""))
    with open(path, "wb") as f:
        f.write(gzip.compress(data) if zipped else data)
    n_reads = rng.choice((1, 3, 400))
    cutoff = rng.choice((0.0, 0.2, 0.6))
    out = os.path.join(tmp, "f%d_ref" % seed)
    desc = "seed %d %s%s k=%d w=%d reads=%d n_reads=%d bad_at=%s style=%s" % (seed, fmt, " gz" if zipped else "", k, w, len(recs), n_reads, bad_at, [s for s in STYLES + RARE if style[s]])
    try:
        pr = subprocess.run([ref_bin, "-r", path, "-i", ibf, "-c", str(cutoff), "-d", "1", "-a", "-u", "-o", out, "-t", "2", "--quiet", "--n-reads", str(n_reads)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=8)
    except subprocess.TimeoutExpired:
        return None, desc + ": the reference binary hangs on this input"
    if pr.returncode < 0 or "terminate called" in pr.stderr:
        return None, desc + " reference died: " + pr.stderr[-160:].strip()
    if pr.returncode != 0:
        return False, desc + " reference failed: " + pr.stderr[-200:]
    want_all, want_unc = R._lines(out + ".all"), R._lines(out + ".unc")
    fasta = library_stream(path, rng.choice((61, 4096, 1 << 20)))
    mine, _err = R.read_file(L, fasta, rng.choice((64, 257, 4096, 1 << 20)), n_reads)
    filt = O.OracleFilter.from_ibf_file(formats.read_ibf(ibf), cutoff)
    res = O.classify_level([filt], [(i, s, None) for i, s in mine], 1.0, 1.0)
    got_all = sorted(b"%s\t%s\t%d" % (r["id"], t.encode(), c) for r in res for t, c in r["matches"])
    got_unc = sorted(r["id"] for r in res if not r["matches"])
    ok = got_all == want_all and got_unc == want_unc
    if not ok:
        desc += " | ref all=%d unc=%d, mine all=%d unc=%d" % (len(want_all), len(want_unc), len(got_all), len(got_unc))
    return ok, desc


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    lib = R.lib(str(tmp_path_factory.mktemp("reads_host")))
    if lib is None:
        pytest.skip("no g++")
    return lib


def test_rewritten_as_two_line_fasta(tmp_path):
    gb = (b"LOCUS       r1 some text   8 bp    DNA\nDEFINITION  x\nORIGIN\n        1 acgtac gt\n//\n"
          b"LOCUS       r2                 4 bp\nORIGIN      \n        1 NNAC\n//\n\n")
    p = tmp_path / "a.gbk"
    p.write_bytes(gb)
    assert library_stream(str(p), 7) == b">r1 some text   8 bp    DNA\nacgtacgt\n>r2                 4 bp\nNNAC\n"
    embl = b"ID   r1; SV 1; 8 BP.\nXX\nDE   Some Sequence\nSQ   Sequence 8 BP;\n     acgtacgt         8\n//\nID r2; 4 BP.\nSQ Sequence 4 BP;\nNNAC\n//\n"
    p = tmp_path / "a.embl"
    p.write_bytes(embl)
    assert library_stream(str(p), 5) == b">r1\nacgtacgt\n>r2\nNNAC\n"
    sam = b"@HD\tVN:1.6\n@SQ\tSN:x\tLN:10\nr1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\nr2 b\t77\t*\t0\t0\t*\t*\t0\t0\tGGNN\t*\tNM:i:0\n"
    p = tmp_path / "a.sam"
    p.write_bytes(sam)
    assert library_stream(str(p)) == b">r1\nACGT\n>r2 b\nGGNN\n"
    q = tmp_path / "b.sam.gz"
    q.write_bytes(gzip.compress(sam))
    assert library_stream(str(q)) == b">r1\nACGT\n>r2 b\nGGNN\n"
    # FASTA / FASTQ names pass through untouched
    p = tmp_path / "a.fa"
    p.write_bytes(b">x\nAC\nGT\n")
    assert library_stream(str(p)) == b">x\nAC\nGT\n"


def test_what_seqan3_throws_on_becomes_a_failing_record(tmp_path, L):
    cases = {
        "a.gb": b"LOCUS       r1   4 bp\nORIGIN\n        1 acgt\n//\nLOKUS r2\nORIGIN\n 1 ac\n//\n",  # not the code word
        "b.gb": b"LOCUS       r1   4 bp\nORIGIN\n        1 acgt\n//\nLOCUS       r2   4 bp\nORIGIN\n        1 acxt\n//\n",  # illegal letter
        "c.embl": b"ID r1; 4 BP.\nSQ Sequence 4 BP;\nacgt\n//\n\n",  # a blank line after the last record: no code word
        "d.embl": b"ID r1; 4 BP.\nSQ Sequence 4 BP;\nacgt\n//\nID r2; 4 BP.\nSQ Sequence 4 BP;\nac-t\n//\n",
        "e.sam": b"r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\nr2\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n",  # no sequence
        "f.sam": b"r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\nr2\t4\t*\t0\t0\t*\t*\t0\t0\tAC.T\tIIII\n",
    }
    for name, data in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        fasta = library_stream(str(p))
        assert fasta.startswith(b">r1") and fasta.endswith(b">\n!\n"), (name, fasta)
        recs, _consumed, err, _msg = R.index_block(L, fasta, True)
        assert [r[0][:2] for r in recs] == [b"r1"] and err == 1, name


def test_records_across_refills(tmp_path):
    """Records larger than one refill of the rewriting stream's buffer, and many small ones."""
    rng = random.Random(3)
    big = F._seq(rng, 9 << 20)
    recs = [(b"big", big)] + [(b"s%d" % i, F._seq(rng, 50)) for i in range(2000)] + [(b"big2", big[::-1])]
    for fmt, ext in (("genbank", "gb"), ("embl", "embl"), ("sam", "sam")):
        p = tmp_path / ("big." + ext)
        p.write_bytes(dress(rng, recs, fmt, {"upper": True}))
        want = b"".join(b">" + (rid + b"   %d bp    DNA     linear   UNK 01-JAN-1980" % len(s) if fmt == "genbank" else rid) + b"\n" + s + b"\n" for rid, s in recs)
        assert library_stream(str(p), 1 << 22) == want, fmt


@needs_ref
@pytest.mark.parametrize("seed", range(30))
def test_formats_differential_against_reference_binary(L, seed, tmp_path):
    ok, desc = differential_case(L, seed, str(tmp_path), F.REF_BIN)
    if ok is None:
        pytest.skip(desc)
    assert ok, desc


if __name__ == "__main__":
    import sys
    import tempfile

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tmp = tempfile.mkdtemp()
    lib = R.lib(tmp)
    bad = died = 0
    for seed in range(first, first + n):
        ok, desc = differential_case(lib, seed, tmp, F.REF_BIN)
        if ok is None:
            died += 1
            print("SKIP", desc)
        elif not ok:
            bad += 1
            print("MISMATCH", desc)
    print("%d cases, %d mismatches, %d skipped (reference died)" % (n, bad, died))
