"""The library's read-file byte stream (gnb_reads_file_*, csrc/gzstream.cpp; reader side of GC.cpp:1220-1287): plain files
by parallel preads, gzip files -- single-member, multi-member, BGZF-like, stored / fixed / dynamic blocks -- by the
chunk-parallel inflater.  Whatever the chunk size and thread count, the consumer sees exactly the bytes zlib produces;
damaged files give an error, never wrong bytes."""
import ctypes as C
import gzip
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

from ganon_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_all(path, threads=0, cap=1 << 20):
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.gnb_reads_file_open(str(path).encode(), threads, C.byref(h))
    assert rc == 0, L.gnb_last_error()
    out = []
    buf = C.create_string_buffer(cap)
    try:
        while True:
            n = L.gnb_reads_file_read(h, buf, cap)
            if n < 0:
                return None, L.gnb_last_error().decode()
            if n == 0:
                break
            out.append(buf.raw[:n])
    finally:
        L.gnb_reads_file_close(h)
    return b"".join(out), None


def read_all_with_chunk(path, threads, chunk):
    """GANON_B200_GZ_CHUNK is read when a file is opened: run in this process with the variable set."""
    old = os.environ.get("GANON_B200_GZ_CHUNK")
    try:
        if chunk:
            os.environ["GANON_B200_GZ_CHUNK"] = str(chunk)
        else:
            os.environ.pop("GANON_B200_GZ_CHUNK", None)
        return read_all(path, threads)
    finally:
        if old is None:
            os.environ.pop("GANON_B200_GZ_CHUNK", None)
        else:
            os.environ["GANON_B200_GZ_CHUNK"] = old


@pytest.fixture(scope="module")
def fastq():
    rng = np.random.default_rng(1)
    g = synth.random_genomes(3, 32, 3000)
    m1, _, _ = synth.reads_from_genomes(5, g, 40000)
    fq = synth.fastq_block(m1)
    rec = fq.size // 40000
    fq = fq.reshape(40000, rec).copy()
    fq[:, rec - 151 : rec - 1] = rng.choice(np.frombuffer(b"FFFFFFF:,#", dtype=np.uint8), size=(40000, 150))
    return fq.tobytes()


def test_plain_and_gzip_variants(tmp_path, fastq):
    rng = np.random.default_rng(2)
    cases = {}
    (tmp_path / "plain.fq").write_bytes(fastq)
    cases["plain.fq"] = fastq
    for lvl in (1, 6, 9):
        with gzip.open(tmp_path / ("l%d.fq.gz" % lvl), "wb", compresslevel=lvl) as f:
            f.write(fastq)
        cases["l%d.fq.gz" % lvl] = fastq
    with open(tmp_path / "members.fq.gz", "wb") as f:  # concatenated members
        for a in range(0, len(fastq), 1_700_000):
            f.write(gzip.compress(fastq[a : a + 1_700_000], 6))
    cases["members.fq.gz"] = fastq
    with open(tmp_path / "bgzf_like.fq.gz", "wb") as f:  # <= 64 KiB members, as bgzip / bcl2fastq write
        for a in range(0, len(fastq), 65280):
            f.write(gzip.compress(fastq[a : a + 65280], 6))
    cases["bgzf_like.fq.gz"] = fastq
    # incompressible bytes (stored blocks), text, a run of zeros (long matches, distance 1), tiny tail (fixed block)
    blob = rng.integers(0, 256, size=700_000, dtype=np.uint8).tobytes() + fastq[:900_000] + bytes(500_000) + b"tail"
    with gzip.open(tmp_path / "mix.gz", "wb") as f:
        f.write(blob)
    cases["mix.gz"] = blob
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)  # fixed Huffman blocks only
    (tmp_path / "fixed.gz").write_bytes(co.compress(fastq[:300_000]) + co.flush())
    cases["fixed.gz"] = fastq[:300_000]
    (tmp_path / "empty.gz").write_bytes(gzip.compress(b""))
    cases["empty.gz"] = b""
    (tmp_path / "small.gz").write_bytes(gzip.compress(b"@r\nACGT\n+\nIIII\n"))
    cases["small.gz"] = b"@r\nACGT\n+\nIIII\n"
    (tmp_path / "garbage_after.gz").write_bytes(gzip.compress(fastq[:100_000]) + b"\0" * 37)  # gzip ignores trailing padding
    cases["garbage_after.gz"] = fastq[:100_000]
    for name, want in cases.items():
        for threads, chunk in ((1, None), (4, None), (4, 65536), (3, 20000), (2, 4096)):
            got, err = read_all_with_chunk(tmp_path / name, threads, chunk)
            assert err is None, (name, threads, chunk, err)
            assert got == want, (name, threads, chunk, len(got), len(want))


def test_damaged_gzip_files_are_errors(tmp_path, fastq):
    z = gzip.compress(fastq[:2_000_000], 6)
    rng = np.random.default_rng(3)
    bad = 0
    (tmp_path / "trunc.gz").write_bytes(z[: len(z) // 2])
    got, err = read_all_with_chunk(tmp_path / "trunc.gz", 3, 65536)
    assert got is None and err
    (tmp_path / "crc.gz").write_bytes(z[:-8] + bytes([z[-8] ^ 1]) + z[-7:])
    got, err = read_all_with_chunk(tmp_path / "crc.gz", 3, 65536)
    assert got is None and "CRC" in err
    for k in range(12):  # a flipped bit in the deflate data: an error, or (if the flip is harmless to the structure) a CRC failure
        pos = int(rng.integers(64, len(z) - 16))
        b = bytearray(z)
        b[pos] ^= 1 << int(rng.integers(0, 8))
        (tmp_path / "flip.gz").write_bytes(bytes(b))
        got, err = read_all_with_chunk(tmp_path / "flip.gz", 3, 32768)
        try:
            want = zlib.decompress(bytes(b), 31)
        except zlib.error:
            want = None
        if want is None:
            assert got is None and err, k
            bad += 1
        else:
            assert got == want, k
    assert bad >= 8


def test_bytes_match_gzip_command_on_a_real_gzip_file(tmp_path, fastq):
    """A file written by the gzip program (not Python's zlib binding): header with file name, its block splitting."""
    p = tmp_path / "reads.fq"
    p.write_bytes(fastq)
    subprocess.check_call(["gzip", "-6", "-k", str(p)])
    for threads, chunk in ((1, None), (4, 100000), (4, 30000)):
        got, err = read_all_with_chunk(str(p) + ".gz", threads, chunk)
        assert err is None and got == fastq


def _bgzf(data: bytes, block: int = 65280) -> bytes:
    """Real BGZF (htslib / bgzip / bcl2fastq layout): every member carries the 'BC' extra field with its compressed size and
    the file ends with the empty end-of-file member."""
    import struct

    out = []
    for a in list(range(0, len(data), block)) + [None]:
        piece = b"" if a is None else data[a : a + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(piece) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + body + struct.pack("<II", zlib.crc32(piece), len(piece)))
    return b"".join(out)


def test_bgzf_and_flush_points(tmp_path, fastq):
    """BGZF with its extra fields and end-of-file member; streams with sync / full flush points (empty stored blocks,
    byte-aligned restarts: what pigz and streaming writers emit); a header with name and comment."""
    rng = np.random.default_rng(5)
    (tmp_path / "real.bgzf.fq.gz").write_bytes(_bgzf(fastq[:3_000_000]))
    for threads, chunk in ((4, None), (3, 30000), (2, 4096)):
        got, err = read_all_with_chunk(tmp_path / "real.bgzf.fq.gz", threads, chunk)
        assert err is None and got == fastq[:3_000_000]
    for seed in range(4):
        data = fastq[: 1_500_000 + 1000 * seed]
        co = zlib.compressobj(int(rng.integers(1, 10)), zlib.DEFLATED, 31)
        parts, pos = [], 0
        while pos < len(data):
            n = int(rng.integers(1, 200_000))
            parts.append(co.compress(data[pos : pos + n]))
            pos += n
            if rng.random() < 0.5:
                parts.append(co.flush(zlib.Z_FULL_FLUSH if rng.random() < 0.5 else zlib.Z_SYNC_FLUSH))
        parts.append(co.flush())
        z = b"".join(parts)
        # a header with FNAME and FCOMMENT in place of the plain one
        z = b"\x1f\x8b\x08\x18\x00\x00\x00\x00\x00\x03" + b"reads.fq\x00" + b"a comment\x00" + z[10:]
        (tmp_path / "flush.gz").write_bytes(z)
        assert zlib.decompress(z, 31) == data
        for threads, chunk in ((4, 20000), (2, 4096), (4, None)):
            got, err = read_all_with_chunk(tmp_path / "flush.gz", threads, chunk)
            assert err is None and got == data, (seed, threads, chunk)


def test_fuzz_against_zlib(tmp_path):
    """Random payload kinds (noise, DNA text, long runs, FASTQ, zeros, byte runs) x random deflate settings (level 0-9, window
    2^9..2^15, memLevel, strategies incl. RLE / Huffman-only / fixed), 1-3 members, random sync / full flush points, random
    thread counts and chunk sizes: the bytes must be zlib's.  (A 1000-case campaign of the same generator ran clean in
    development.)"""
    rng = np.random.default_rng(2024)

    def make(kind, n):
        if kind == 0:
            return rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        if kind == 1:
            return bytes(rng.choice(list(b"ACGT\n"), size=n).astype(np.uint8))
        if kind == 2:
            return (b"A" * int(rng.integers(1, 5000)) + b"\n") * max(1, n // 2500)
        if kind == 3:
            recs, size = [], 0
            while size < n:
                L = int(rng.integers(30, 300))
                s = bytes(rng.choice(list(b"ACGTN"), size=L).astype(np.uint8))
                q = bytes(rng.choice(list(b"FFFF:,#"), size=L).astype(np.uint8))
                recs.append(b"@read%d some/1\n%s\n+\n%s\n" % (len(recs), s, q))
                size += len(recs[-1])
            return b"".join(recs)
        if kind == 4:
            return bytes(n)
        return b"".join(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 400)) for _ in range(n // 200 + 1))[:n]

    for case in range(24):
        data = make(int(rng.integers(0, 6)), int(rng.integers(1, 1_200_000)))
        level, wbits, mem = int(rng.integers(0, 10)), int(rng.integers(9, 16)), int(rng.integers(1, 10))
        strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        members = int(rng.integers(1, 4))
        cuts = sorted(rng.integers(0, len(data) + 1, size=members - 1).tolist()) + [len(data)]
        parts, pos = [], 0
        for e in cuts:
            co = zlib.compressobj(level, zlib.DEFLATED, 16 + wbits, mem, strat)
            piece, p = data[pos:e], 0
            pos = e
            while p < len(piece):
                m = int(rng.integers(1, 300000))
                parts.append(co.compress(piece[p : p + m]))
                p += m
                if rng.random() < 0.2:
                    parts.append(co.flush(zlib.Z_SYNC_FLUSH if rng.random() < 0.5 else zlib.Z_FULL_FLUSH))
            parts.append(co.flush())
        (tmp_path / "f.gz").write_bytes(b"".join(parts))
        threads, chunk = int(rng.integers(1, 7)), int(rng.choice([4096, 8000, 20000, 65536, 300000, 0])) or None
        got, err = read_all_with_chunk(tmp_path / "f.gz", threads, chunk)
        assert err is None and got == data, (case, level, wbits, mem, strat, members, threads, chunk, err)


def test_closing_a_gzip_file_early_stops_its_workers(tmp_path, fastq):
    """A reader that is closed while finders and decoders are still at work (right after opening, or after a few bytes) must
    come down cleanly -- the workers, the sequencer and the queued pieces -- and a fresh reader of the same file still works."""
    L = _lib.lib()
    (tmp_path / "r.fq.gz").write_bytes(gzip.compress(fastq, 6))
    buf = C.create_string_buffer(5000)
    for threads, chunk, reads in ((4, 65536, 0), (8, 4096, 1), (3, None, 2), (6, 20000, 0), (1, None, 1), (16, 65536, 3)):
        old = os.environ.get("GANON_B200_GZ_CHUNK")
        if chunk:
            os.environ["GANON_B200_GZ_CHUNK"] = str(chunk)
        try:
            for _ in range(3):
                h = C.c_void_p()
                assert L.gnb_reads_file_open(str(tmp_path / "r.fq.gz").encode(), threads, C.byref(h)) == 0
                for _ in range(reads):
                    assert L.gnb_reads_file_read(h, buf, 5000) == 5000
                    assert buf.raw[:5000] == fastq[:5000] or reads > 1
                L.gnb_reads_file_close(h)
        finally:
            if old is None:
                os.environ.pop("GANON_B200_GZ_CHUNK", None)
            else:
                os.environ["GANON_B200_GZ_CHUNK"] = old
    got, err = read_all(tmp_path / "r.fq.gz", 4)
    assert err is None and got == fastq
