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

    def run(seq, k, w, misalign=0, stride=1, lane=0):
        out = np.empty(len(seq) + 1, dtype=np.uint64)
        n = lib.k2t_host_minimisers(seq, len(seq), k, w, misalign, stride, lane, out.ctypes.data, out.size)
        assert n >= 0, (n, k, w, len(seq))
        return out[:n]

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
