"""bzip2-compressed read files through the library's byte-stream reader (ganon_b200/csrc/bz2stream.cpp behind
`gnb_reads_file_*`, no GPU involved; SURVEY 8a row A0: the reference reads them through seqan3's transparent decompression,
misc_input.hpp:145-153, and cannot be built without libbz2, CMakeLists.txt:114).  Checked against Python's bz2 module: every
block size, block-aligned and bit-shifted blocks, long runs, incompressible data, concatenated streams, the other sequence
formats under a .bz2 suffix, damaged and truncated files (an error, never wrong bytes), and segments cut at a place that is no
block start (what a chance occurrence of the block number inside compressed data would do)."""
import bz2
import ctypes as C
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stream(path, threads=4, piece=1 << 20):
    """-> (bytes read before the end or an error, error message or None)"""
    from ganon_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    assert L.gnb_reads_file_open(path.encode(), threads, C.byref(h)) == 0, L.gnb_last_error()
    buf = C.create_string_buffer(piece)
    out, err = [], None
    try:
        while True:
            n = L.gnb_reads_file_read(h, buf, piece)
            if n < 0:
                err = L.gnb_last_error()
                break
            if n == 0:
                break
            out.append(buf.raw[:n])
    finally:
        L.gnb_reads_file_close(h)
    return b"".join(out), err


def fastq(n, seed, length=100):
    r = random.Random(seed)
    return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(r.choice(b"ACGT") for _ in range(length)), b"I" * length) for i in range(n))


CASES = {
    "empty": lambda: b"",
    "one_record": lambda: b"@a\nACGT\n+\nIIII\n",
    "fastq": lambda: fastq(3000, 1),
    "fastq_many_blocks": lambda: fastq(12000, 2),  # 2.6 MB: 27 blocks at level 1
    "runs": lambda: b"A" * 2_000_000 + b"C" * 10 + b"G" * 3_000_000,  # blocks that expand 50-fold
    "incompressible": lambda: random.Random(3).randbytes(400_000),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_equals_pythons_bz2(name, tmp_path):
    data = CASES[name]()
    for level in (1, 5, 9):
        p = str(tmp_path / ("%s_%d.fq.bz2" % (name, level)))
        open(p, "wb").write(bz2.compress(data, level))
        for threads, piece in ((1, 1 << 20), (4, 4099), (8, 1 << 22)):
            if piece == 4099 and len(data) > 1_000_000:
                continue
            got, err = stream(p, threads, piece)
            assert err is None and got == data, (name, level, threads, piece, err, len(got))


def test_concatenated_streams(tmp_path):
    a, b = fastq(2500, 4), fastq(9000, 5)
    p = str(tmp_path / "cat.fq.bz2")
    open(p, "wb").write(bz2.compress(a, 9) + bz2.compress(b"", 9) + bz2.compress(b, 1) + bz2.compress(b"x", 3))
    assert stream(p) == (a + b + b"x", None)


def test_other_sequence_formats_under_bz2(tmp_path):
    sam = b"@HD\tVN:1.6\nr1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\nr2\t4\t*\t0\t0\t*\t*\t0\t0\tGGNN\t*\n"
    p = str(tmp_path / "a.sam.bz2")
    open(p, "wb").write(bz2.compress(sam))
    assert stream(p) == (b">r1\nACGT\n>r2\nGGNN\n", None)


def test_damaged_and_truncated_files_are_errors(tmp_path):
    data = fastq(12000, 6)
    comp = bz2.compress(data, 1)
    rng = random.Random(7)
    for trial in range(18):
        c = bytearray(comp)
        if trial % 3 == 0:
            c = c[: rng.randrange(20, len(c) - 1)]
        elif trial % 3 == 1:
            i = rng.randrange(10, len(c) - 10)
            c[i] ^= 1 << rng.randrange(8)
        else:
            c = c[: -rng.randrange(1, 11)]  # the end-of-stream trailer cut
        p = str(tmp_path / ("bad%d.fq.bz2" % trial))
        open(p, "wb").write(bytes(c))
        try:
            want = bz2.decompress(bytes(c))
        except (OSError, ValueError, EOFError):
            want = None
        got, err = stream(p)
        if want is None:
            assert err is not None and data.startswith(got), (trial, err, len(got))
        else:  # a flipped bit in the padding or the (unchecked) combined CRC of the stream
            assert got == want or (err is not None and data.startswith(got)), (trial, err)


def test_segments_cut_where_no_block_starts_are_put_together_again(tmp_path):
    data = fastq(12000, 8)
    p = str(tmp_path / "split.fq.bz2")
    open(p, "wb").write(bz2.compress(data, 1))
    code = ("import sys; sys.path.insert(0, %r); from tests.test_bz2_cpu import stream; import bz2; "
            "got, err = stream(%r, 4); assert err is None and got == bz2.decompress(open(%r, 'rb').read()), err; print('ok')" % (ROOT, p, p))
    env = dict(os.environ, GANON_B200_BZ2_SPLIT="1")
    done = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert done.returncode == 0 and done.stdout.strip() == "ok", done.stderr[-2000:]
