"""Seeded synthetic databases and reads for the benchmark and the parity tests.

Everything here is data generation in numpy: bit-identical to the device-side generators of libganon_b200
(gnb_db_fill_random uses the same splitmix64 construction), so that the same database can be written to a .ibf
file for the reference binary and generated directly in HBM for the GPU arm.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

SEEDS = np.array([13572355802537770549, 13043817825332782213, 10650232656628343401, 16499269484942379435, 4893150838803335377], dtype=np.uint64)
MUL = np.uint64(11400714819323198485)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def random_words(seed: int, and_terms: int, bin_size: int, bin_words: int, bins: int, row0: int = 0, rows: int = None) -> np.ndarray:
    """Rows [row0, row0+rows) of the bitvector gnb_db_fill_random(seed, and_terms) produces, flattened."""
    rows = bin_size - row0 if rows is None else rows
    i = np.arange(row0 * bin_words, (row0 + rows) * bin_words, dtype=np.uint64)
    v = np.full(i.size, np.uint64(0xFFFFFFFFFFFFFFFF))
    with np.errstate(over="ignore"):
        for t in range(and_terms):
            v &= splitmix64(np.uint64(seed) + i * np.uint64(8) + np.uint64(t))
    if bins % 64:
        v = v.reshape(rows, bin_words)
        v[:, bins // 64] &= np.uint64((1 << (bins % 64)) - 1)
        v[:, bins // 64 + 1 :] = 0
        v = v.reshape(-1)
    return v


def ibf_rows(hashes: np.ndarray, h: int, bin_size: int) -> np.ndarray:
    """hash_and_fit rows (IBF.hpp:173-187) for every hash and hash function: uint64[h, n]."""
    hashes = np.asarray(hashes, dtype=np.uint64)
    shift = np.uint64(64 - int(bin_size).bit_length())
    out = np.empty((h, hashes.size), dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(h):
            x = hashes * SEEDS[i]
            x ^= x >> shift
            x *= MUL
            # mulhi(x, bin_size) with 32-bit limbs
            lo, hi = x & np.uint64(0xFFFFFFFF), x >> np.uint64(32)
            b = np.uint64(bin_size)
            blo, bhi = b & np.uint64(0xFFFFFFFF), b >> np.uint64(32)
            t = lo * blo
            t1 = hi * blo + (t >> np.uint64(32))
            t2 = lo * bhi + (t1 & np.uint64(0xFFFFFFFF))
            out[i] = hi * bhi + (t1 >> np.uint64(32)) + (t2 >> np.uint64(32))
    return out


def emplace_numpy(data: np.ndarray, bin_words: int, bin_size: int, h: int, hashes: np.ndarray, bins: np.ndarray) -> None:
    """IBF emplace (IBF.hpp:271-286) into a host bitvector (flattened [row][word])."""
    rows = ibf_rows(hashes, h, bin_size)
    bins = np.asarray(bins, dtype=np.uint64)
    for i in range(h):
        idx = rows[i] * np.uint64(bin_words) + (bins >> np.uint64(6))
        np.bitwise_or.at(data, idx.astype(np.int64), np.uint64(1) << (bins & np.uint64(63)))


def random_genomes(seed: int, n: int, length: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, length), dtype=np.uint8)]


_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def reads_from_genomes(seed: int, genomes: np.ndarray, n: int, length: int = 150, frac_planted: float = 0.5, sub_rate: float = 0.01, n_rate: float = 0.001, paired: bool = False, insert: int = 300):
    """n reads (or pairs): `frac_planted` sampled from the genomes (uniform position, random strand, substitutions),
    the rest uniform random; a fraction n_rate of all bases becomes 'N'.  Returns uint8[n, length] (and mates)."""
    rng = np.random.default_rng(seed)
    G, glen = genomes.shape
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    span = insert if paired else length
    n_pl = int(n * frac_planted)
    g = rng.integers(0, G, size=n_pl)
    p = rng.integers(0, glen - span, size=n_pl)
    idx = p[:, None] + np.arange(span)[None, :]
    frag = genomes[g[:, None], idx]
    strand = rng.random(n_pl) < 0.5
    rc = _COMP[frag[:, ::-1]]
    frag = np.where(strand[:, None], rc, frag)
    rnd = acgt[rng.integers(0, 4, size=(n - n_pl, span), dtype=np.uint8)]
    frag = np.concatenate([frag, rnd], axis=0)
    perm = rng.permutation(n)
    frag = frag[perm]
    origin = np.concatenate([g, np.full(n - n_pl, -1)])[perm]

    def noise(m):
        m = m.copy()
        sub = rng.random(m.shape) < sub_rate
        m[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
        m[rng.random(m.shape) < n_rate] = ord("N")
        return m

    m1 = noise(frag[:, :length])
    if not paired:
        return m1, None, origin
    m2 = noise(_COMP[frag[:, ::-1]][:, :length])
    return m1, m2, origin


def fastq_block(seqs: np.ndarray, first_index: int = 0, suffix: bytes = b"") -> np.ndarray:
    """Fixed-width 4-line FASTQ records `@r<9 digits><suffix>\\n<seq>\\n+\\n<qual>\\n` as one uint8 array."""
    n, L = seqs.shape
    idw = 2 + 9 + len(suffix)
    rec = np.empty((n, idw + 1 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    idx = np.arange(first_index, first_index + n, dtype=np.int64)
    for d in range(9):
        rec[:, 2 + 8 - d] = (idx % 10 + ord("0")).astype(np.uint8)
        idx //= 10
    for j, c in enumerate(suffix):
        rec[:, 11 + j] = c
    o = idw
    rec[:, o] = ord("\n")
    rec[:, o + 1 : o + 1 + L] = seqs
    o += 1 + L
    rec[:, o : o + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, o + 3 : o + 3 + L] = ord("I")
    rec[:, o + 3 + L] = ord("\n")
    return rec.reshape(-1)


def planted_hashes(genomes: np.ndarray, minimiser_fn: Callable[[bytes], np.ndarray], bins_of_genome: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray, List[int]]:
    """Unique minimiser hashes of every genome spread round-robin over the genome's bins (create_bin_map_hash,
    GanonBuild.cpp:619-653 distributes by contiguous ranges; any split works for a synthetic database)."""
    hs, bs, counts = [], [], []
    for g, bins in zip(genomes, bins_of_genome):
        u = np.unique(minimiser_fn(g.tobytes()))
        hs.append(u)
        b = np.asarray(bins, dtype=np.uint32)
        bs.append(b[np.arange(u.size) % b.size])
        counts.append(int(u.size))
    return np.concatenate(hs), np.concatenate(bs), counts

