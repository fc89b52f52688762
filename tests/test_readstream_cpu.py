"""Host-side block reader of the command line (ganon_b200/classify.py:_ReadStream; reader of GC.cpp:1220-1287): whatever
the block size, the headroom and the compression, the consumer sees the file's bytes exactly once and in order."""
import gzip
import os

import numpy as np
import pytest

from ganon_b200 import classify as K


def _fastq(n, rng, long_every=0):
    recs = []
    for i in range(n):
        L = int(rng.integers(20, 300)) if not (long_every and i % long_every == 0) else 5000
        s = bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8))
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * L))
    return recs


def _consume(stream, records_total):
    """Take whole 4-line records from every block, as Session.submit reports them through `consumed`."""
    out, grows = [], 0
    while True:
        stream.next_block()
        if stream.fill == 0:
            break
        block = bytes(stream.block_bytes_view())
        final = stream.eof
        lines = block.count(b"\n")
        n_rec = lines // 4
        if final and not block.endswith(b"\n"):
            n_rec = (lines + 1) // 4
        if n_rec == 0 and not final:
            stream.whole_block_to_tail()
            stream.grow()
            grows += 1
            continue
        pos = 0
        for _ in range(n_rec * 4):
            nl = block.find(b"\n", pos)
            pos = len(block) if nl < 0 else nl + 1
        out.append(block[:pos])
        stream.consume(pos)
        if final:
            break
    return b"".join(out), grows


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("block,head", [(1000, 64), (4096, 1 << 20), (50000, 300), (1 << 20, 1 << 20)])
def test_blocks_cover_the_file_in_order(tmp_path, monkeypatch, gz, block, head):
    rng = np.random.default_rng(block + head + gz)
    recs = _fastq(400, rng, long_every=97)
    data = b"".join(recs)
    p = str(tmp_path / ("r.fq.gz" if gz else "r.fq"))
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(data)
    monkeypatch.setattr(K, "_HEADROOM", head)
    monkeypatch.setattr(K, "_IO_SLICE", 777)
    s = K._ReadStream(p, block, 5)
    try:
        got, grows = _consume(s, len(recs))
    finally:
        s.close()
    assert got == data
    if block < 5000:
        assert grows >= 1  # the 5 kbp records do not fit the first block size


def test_missing_final_newline_and_empty_tail(tmp_path):
    data = b"@a\nACGT\n+\nIIII\n@b\nAC\n+\nII"
    p = str(tmp_path / "x.fq")
    open(p, "wb").write(data)
    s = K._ReadStream(p, 1 << 16, 3)
    try:
        got, _ = _consume(s, 2)
    finally:
        s.close()
    assert got == data


def _bgzf(data: bytes, rng) -> bytes:
    """BGZF as bgzip writes it: gzip members of <= 64 KiB with the BC extra field, then the empty EOF block."""
    import struct
    import zlib

    out, pos = [], 0
    while pos <= len(data):
        n = int(rng.integers(1, 65000)) if pos < len(data) else 0
        chunk = data[pos : pos + n]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 18 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + body + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        if pos == len(data):
            break
        pos += n
    return b"".join(out)


@pytest.mark.parametrize("block", [3000, 100000, 1 << 20])
def test_bgzf_blocks_are_inflated_in_parallel_and_in_order(tmp_path, monkeypatch, block):
    rng = np.random.default_rng(block)
    data = b"".join(_fastq(3000, rng))
    raw = _bgzf(data, rng)
    p = str(tmp_path / "r.fq.gz")
    open(p, "wb").write(raw)
    assert gzip.open(p, "rb").read() == data  # a valid multi-member gzip for everyone else
    assert K._BgzfReader.is_bgzf(p)
    monkeypatch.setattr(K._BgzfReader, "CHUNK", 150000)
    s = K._ReadStream(p, block, 5)
    assert isinstance(s.f, K._BgzfReader)
    try:
        got, _ = _consume(s, 3000)
    finally:
        s.close()
    assert got == data
    # a plain .gz keeps the gzip module
    q = str(tmp_path / "plain.fq.gz")
    with gzip.open(q, "wb") as f:
        f.write(data[:5000])
    assert not K._BgzfReader.is_bgzf(q)
