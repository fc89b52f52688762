"""The per-thread core of the K2t kernel (ganon_b200/csrc/k2_thread.cuh, one thread per read) compiled for the host
(tests/native/k2t_host.cpp: the same source the device runs, plain memory in place of shared / global memory) against the
oracle, which tests/test_oracle.py pins to seqan3's known-answer tests and to the reference binary.  Bit-exact."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEG, WARM = 512, 64  # k2t::kSegWindows, kSegWarm


@pytest.fixture(scope="module")
def k2t(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("k2t") / "k2t_host.so")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "native", "k2t_host.cpp")])
    lib = ctypes.CDLL(so)
    lib.k2t_host_minimisers.restype = ctypes.c_long
    lib.k2t_host_minimisers.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64]

    lib.k2t_host_batch.restype = ctypes.c_long
    lib.k2t_host_batch.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    def run(seq, k, w, misalign=0, stride=1, lane=0):
        out = np.empty(len(seq) + 1, dtype=np.uint64)
        n = lib.k2t_host_minimisers(seq, len(seq), k, w, misalign, stride, lane, out.ctypes.data, out.size)
        assert n >= 0, (n, k, w, len(seq))
        return out[:n]

    lib.k2t_host_segments.restype = ctypes.c_long
    lib.k2t_host_segments.argtypes = [ctypes.c_char_p] + [ctypes.c_uint32] * 6 + [ctypes.c_void_p, ctypes.c_void_p]

    def segments(seq, k, w, misalign=0, stride=1, lane=0):
        """[(values emitted by the segment, flagged)] of k2t::segment over every segment of seq."""
        windows = len(seq) - w + 1
        out = np.full(windows + 1, 0xDEAD, dtype=np.uint64)
        cnt = np.zeros((windows + SEG - 1) // SEG + 1, dtype=np.uint32)
        n = lib.k2t_host_segments(seq, len(seq), k, w, misalign, stride, lane, out.ctypes.data, cnt.ctypes.data)
        assert n == (windows + SEG - 1) // SEG
        return [(out[g * SEG : g * SEG + int(cnt[g] & 0x7FFFFFFF)], bool(cnt[g] >> 31)) for g in range(n)]

    run.batch = lib.k2t_host_batch
    run.segments = segments
    return run


def test_seqan3_known_answers(k2t):
    # libs/seqan3/test/unit/search/views/minimiser_hash_test.cpp:62-79 with the ganon seed (adjust_seed)
    for seq, k, w in [(b"ACGGCGACGTTTAG", 4, 8), (b"ACGTCGACGTTTAG", 4, 8), (b"A" * 19, 4, 8), (b"ACGGCGACG", 4, 8), (b"A" * 19, 19, 19)]:
        for mis in range(8):
            assert k2t(seq, k, w, mis).tolist() == O.minimiser_hash(seq, k, w).tolist(), (seq, k, w, mis)


@pytest.mark.parametrize("k,w", [(19, 31), (10, 10), (4, 8), (29, 60), (29, 29), (1, 1), (1, 32), (28, 35), (27, 29), (26, 33), (12, 14), (16, 47), (17, 48), (5, 36)])
def test_adversarial_sequences(k2t, k, w):
    rng = np.random.default_rng(k * 1000 + w)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8)) for L in [w, w + 1, w + 7, w + 8, w + 9, 75, 150, 151, 255, 256, 257, 300, 1000] if L >= w]
    seqs += [b"A" * 200, b"C" * 150, b"AC" * 100, b"ACG" * 80, b"ACGT" * 70, b"AT" * 300, (b"ACGGT" * 7 + b"T") * 30, b"N" * 150]
    seqs.append(bytes(rng.choice(list(b"ACGTNRYSWKMBDHVUacgtn"), size=400).astype(np.uint8)))
    seqs.append(bytes(rng.choice(list(b"AT"), size=700).astype(np.uint8)))  # many ties
    seqs.append(bytes(rng.choice(list(b"ACGT"), size=37).astype(np.uint8)) * 40)  # long-period repeat
    for i, s in enumerate(seqs):
        if len(s) < w:
            continue
        assert k2t(s, k, w, i & 7, 1 + i % 3, i % (1 + i % 3)).tolist() == O.minimiser_hash(s, k, w).tolist(), (k, w, len(s), s[:40])


def test_random_parameters(k2t):
    rng = np.random.default_rng(20261017)
    alphabets = [b"ACGT", b"AC", b"A", b"ACGTN", b"ACGTacgtNnRYKMSWBDHVUu", b"AT", b"ACGTRYKMSWBDHVN"]
    for it in range(3000):
        k = int(rng.integers(1, 30))
        W = int(rng.integers(1, 33))
        w = k + W - 1
        L = w + int(rng.integers(0, 5)) if rng.random() < 0.2 else int(rng.integers(w, w + 400))
        a = alphabets[int(rng.integers(0, len(alphabets)))]
        mode = rng.random()
        if mode < 0.3:  # periodic: many ties
            per = int(rng.integers(1, 9))
            seq = (bytes(rng.choice(list(a), size=per).astype(np.uint8)) * (L // per + 1))[:L]
        elif mode < 0.4:  # a sequence followed by its reverse complement: forward == reverse ties
            half = bytes(rng.choice(list(b"ACGT"), size=L // 2 + 1).astype(np.uint8))
            seq = (half + half[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA")))[:L]
        else:
            seq = bytes(rng.choice(list(a), size=L).astype(np.uint8))
        stride = int(rng.integers(1, 6))
        got = k2t(seq, k, w, int(rng.integers(0, 8)), stride, int(rng.integers(0, stride)))
        assert got.tolist() == O.minimiser_hash(seq, k, w).tolist(), (it, k, w, L, seq[:60])


def test_every_byte_value_decodes_like_dna4(k2t):
    # the character table against the oracle's char_to_rank on all 256 byte values (non-letters count as 'A')
    k, w = 3, 5
    for c in range(1, 256):  # ctypes c_char_p stops at NUL
        seq = b"ACGTTGCA" + bytes([c]) * 3 + b"GATTACA"
        assert k2t(seq, k, w).tolist() == O.minimiser_hash(seq, k, w).tolist(), c


def _block(rng, seqs):
    """A FASTQ block like the ones K1 indexes: (bytes with 64 bytes of slack on both sides, offsets, lengths)."""
    parts, off, ln, pos = [b"#" * 64], [], [], 64
    for i, s in enumerate(seqs):
        head = b"@r%d some text\n" % i
        parts += [head, s, b"\n+\n", b"I" * len(s), b"\n"]
        off.append(pos + len(head))
        ln.append(len(s))
        pos += len(head) + 2 * len(s) + 4
    parts.append(b"#" * 64)
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy(), np.asarray(off, dtype=np.uint32), np.asarray(ln, dtype=np.uint32)


@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("k,w", [(19, 31), (4, 8), (29, 60), (12, 12)])
def test_batches_follow_the_classify_rules(k2t, k, w, paired):
    """k2t::read_pair (the per-thread body of the kernel) over whole batches, in the kernel's three modes, against the hash lists
    ganon-classify builds (GanonClassify.cpp:690-700: read shorter than the window skipped, short mate ignored, mate 1 ++ mate 2)."""
    rng = np.random.default_rng(k * 100 + w + paired)
    n = 300
    lens = lambda: [int(rng.choice([w - 1, w, w + 1, 10, 75, 150, 151, 260])) for _ in range(n)]
    mk = lambda L: bytes(rng.choice(list(b"ACGTN"), size=max(L, 1)).astype(np.uint8))
    s1 = [mk(L) for L in lens()]
    s2 = [mk(L) for L in lens()] if paired else None
    b1, o1, l1 = _block(rng, s1)
    b2, o2, l2 = _block(rng, s2) if paired else (None, None, None)
    want = [O.read_hashes(a, s2[i] if paired else None, k, w) for i, a in enumerate(s1)]
    want = [np.empty(0, dtype=np.uint64) if h is None else h for h in want]
    ptr = lambda a: a.ctypes.data if a is not None else None
    # mode 0: counts
    counts = np.full(n, 77, dtype=np.uint32)
    total = k2t.batch(ptr(b1), ptr(o1), ptr(l1), ptr(b2), ptr(o2), ptr(l2), n, k, w, 0, ptr(counts), None, None)
    assert counts.tolist() == [h.size for h in want] and total == sum(h.size for h in want)
    # mode 1: exact offsets
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(counts)
    hashes = np.zeros(int(off[-1]) + 1, dtype=np.uint64)
    k2t.batch(ptr(b1), ptr(o1), ptr(l1), ptr(b2), ptr(o2), ptr(l2), n, k, w, 1, None, ptr(off), ptr(hashes))
    assert hashes[: int(off[-1])].tolist() == np.concatenate(want).tolist()
    # mode 2: upper-bound offsets (one slot per window, k_hash_upper_bounds), counts written alongside
    ub = np.asarray([max(a - w + 1, 0) + (max(int(l2[i]) - w + 1, 0) if paired else 0) for i, a in enumerate(l1.tolist())], dtype=np.uint64)
    off2 = np.zeros(n + 1, dtype=np.uint64)
    off2[1:] = np.cumsum(ub)
    hashes2 = np.zeros(int(off2[-1]) + 1, dtype=np.uint64)
    counts2 = np.full(n, 99, dtype=np.uint32)
    k2t.batch(ptr(b1), ptr(o1), ptr(l1), ptr(b2), ptr(o2), ptr(l2), n, k, w, 2, ptr(counts2), ptr(off2), ptr(hashes2))
    assert counts2.tolist() == counts.tolist()
    for i in range(n):
        assert hashes2[int(off2[i]) : int(off2[i]) + int(counts2[i])].tolist() == want[i].tolist(), i


# --------------------------------------------------------------------------------------------------- segments of long sequences
def _emissions(seq, k, w):
    """(window index, value) of every minimiser the reference's state machine emits (minimiser.hpp:421-472), written out window by
    window: the k-mer values come from the oracle (window = k: one value per k-mer); pinned to the oracle's list below."""
    vals = O.minimiser_hash(seq, k, k).tolist()
    assert len(vals) == len(seq) - k + 1
    W = w - k + 1
    rightmost_min = lambda a: max(i for i, v in enumerate(a) if v == min(a))
    pos = rightmost_min(vals[:W])  # absolute k-mer index of the tracked minimiser
    out = [(0, vals[pos])]
    for j in range(W, len(vals)):  # k-mer j enters, window index j - W + 1
        if pos == j - W:
            pos = j - W + 1 + rightmost_min(vals[j - W + 1 : j + 1])
            out.append((j - W + 1, vals[pos]))
        elif vals[j] < vals[pos]:
            pos = j
            out.append((j - W + 1, vals[j]))
    assert [v for _, v in out] == O.minimiser_hash(seq, k, w).tolist()
    return out


def _check_segments(k2t, seq, k, w, **kw):
    """Every segment that is not flagged holds exactly the reference's minimisers of its windows; returns the flags."""
    want = _emissions(seq, k, w)
    got = k2t.segments(seq, k, w, **kw)
    for g, (vals, flagged) in enumerate(got):
        if not flagged:
            assert vals.tolist() == [v for j, v in want if g * SEG <= j < (g + 1) * SEG], (g, k, w, len(seq), seq[:50])
    assert not got[0][1]  # the first segment starts where the sequence does
    return [f for _, f in got]


@pytest.mark.parametrize("k,w", [(19, 31), (4, 8), (29, 60), (12, 12), (1, 32), (10, 41), (27, 29)])
def test_segments_of_random_sequences_are_exact_and_never_flagged(k2t, k, w):
    rng = np.random.default_rng(k * 77 + w)
    for i, L in enumerate([w + SEG - 1, w + SEG, w + SEG + 1, w + 2 * SEG - 1, w + 2 * SEG, 3000, 5 * SEG + w + 17]):
        seq = bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8))
        flags = _check_segments(k2t, seq, k, w, misalign=i & 7, stride=1 + i % 3, lane=i % (1 + i % 3))
        if k >= 10:  # no repeated k-mer within a window in random text of this k: every warm-up reaches a certain state
            assert not any(flags), (k, w, L)


def test_segments_inside_repeats_are_flagged_not_wrong(k2t):
    """Homopolymers and tandem repeats with a period below the window keep equal values in every window: the state of a walk that
    starts inside depends on where the repeat began, so those segments must come back flagged; the segments after the repeat
    ended (a warm-up past it) are exact again."""
    rng = np.random.default_rng(5)
    rnd = lambda n: bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))
    k, w = 19, 31
    flags = _check_segments(k2t, b"A" * 3000, k, w)
    assert flags == [False] + [True] * (len(flags) - 1)
    flags = _check_segments(k2t, b"TTAGGG" * 500, k, w)
    assert all(flags[1:])
    flags = _check_segments(k2t, rnd(700) + b"A" * 1500 + rnd(1500), k, w)  # repeat over windows ~[690, 2200)
    assert flags[:2] == [False, False] and flags[2:5] == [True, True, True] and not any(flags[5:])
    flags = _check_segments(k2t, rnd(600) + b"AC" * 40 + rnd(1400), k, w)  # a short repeat away from the segment boundaries
    assert not any(flags)
    # a repeat that ends inside the warm-up of a segment: exact without a flag if a certain event follows, flagged otherwise
    for tail in range(0, WARM + 8, 3):
        _check_segments(k2t, rnd(100) + b"CA" * 180 + rnd(2 * SEG + 40)[: SEG + 200 + tail], k, w)
        seq = rnd(SEG - 200 - tail) + b"G" * (200 + k) + rnd(SEG + 300)
        _check_segments(k2t, seq, k, w)


def test_segments_random_parameters(k2t):
    rng = np.random.default_rng(2026101718)
    alphabets = [b"ACGT", b"AC", b"A", b"ACGTN", b"AT", b"ACGTRYKMSWBDHVN"]
    n_flagged = n_seg = 0
    for it in range(250):
        k = int(rng.integers(1, 30))
        W = int(rng.integers(1, 33))
        w = k + W - 1
        L = w + int(rng.integers(SEG - 2, 4 * SEG))
        a = alphabets[int(rng.integers(0, len(alphabets)))]
        parts, left = [], L
        while left > 0:  # random text interleaved with repeats of random period and length
            n = min(left, int(rng.integers(1, 900)))
            if rng.random() < 0.4:
                per = int(rng.integers(1, 12))
                parts.append((bytes(rng.choice(list(a), size=per).astype(np.uint8)) * (n // per + 1))[:n])
            else:
                parts.append(bytes(rng.choice(list(a), size=n).astype(np.uint8)))
            left -= n
        stride = int(rng.integers(1, 6))
        flags = _check_segments(k2t, b"".join(parts), k, w, misalign=int(rng.integers(0, 8)), stride=stride, lane=int(rng.integers(0, stride)))
        n_flagged += sum(flags)
        n_seg += len(flags)
    assert 0 < n_flagged < n_seg  # both outcomes were exercised
