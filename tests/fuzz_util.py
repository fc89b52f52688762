"""Randomised differential cases for the oracle: small random databases / reads / parameters are written to disk, the
UNMODIFIED reference binary (oracle/_ref/ganon-classify, compiled from /root/reference by oracle/Makefile) classifies them,
and the oracle restatement must give the same `.all` and `.unc` lines.  Used by tests/test_oracle_fuzz.py (a few seeds per
run) and by `python -m tests.fuzz_util N` (a long campaign; round 1: 400 seeds, no mismatch)."""
import os
import random
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from ganon_b200 import formats  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ganon-classify")
IUPAC = b"NRYSWKMBDHVUnacgtryswkmbdhvu"
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(alphabet) for _ in range(n))


def make_db(rng, path, k, w):
    """A random flat .ibf: non-power-of-two sizes, targets of 1..4 bins (not necessarily adjacent), shuffled bin map,
    some targets missing from hashes_count (fpr 0), background noise."""
    n_bins = rng.choice((1, 3, 17, 64, 65, 130, 200))
    bin_size = rng.choice((521, 1009, 4099, 20011, 65537))
    hf = rng.randint(1, 5)
    ibf = O.OracleIBF(n_bins, bin_size, hf)
    nprng = np.random.default_rng(rng.randrange(1 << 30))
    dens = rng.choice((0, 2, 3))
    if dens:
        noise = nprng.integers(0, 1 << 63, size=ibf.data.size, dtype=np.uint64)
        for _ in range(dens - 1):
            noise &= nprng.integers(0, 1 << 63, size=ibf.data.size, dtype=np.uint64)
        noise = noise.reshape(bin_size, ibf.bin_words)
        rem = n_bins - 64 * (ibf.bin_words - 1)
        if rem < 64:  # padding bins stay zero (IBF.hpp:238-240)
            noise[:, -1] &= np.uint64((1 << rem) - 1)
        ibf.data[:] = noise.reshape(-1)
    order = list(range(n_bins))
    if rng.random() < 0.5:
        rng.shuffle(order)  # a target's bins need not be adjacent
    genomes, bin_map, hashes_count = {}, [], []
    b = t = 0
    while b < n_bins:
        nb = min(rng.choice((1, 1, 2, 3, 4)), n_bins - b)
        name = "T%d.%d" % (rng.randrange(1000), t)
        g = _seq(rng, rng.choice((200, 600, 2000)))
        genomes[name] = g
        uniq = sorted(set(int(x) for x in O.minimiser_hash(g, k, w)))
        mine = order[b : b + nb]
        for i, h in enumerate(uniq):
            ibf.emplace(h, mine[i % nb])
        bin_map += [(x, name) for x in mine]
        if rng.random() < 0.9:
            hashes_count.append((name, max(1, len(uniq))))
        b += nb
        t += 1
    rng.shuffle(bin_map)
    per_bin = [max(1, -(-c // sum(1 for _b, tt in bin_map if tt == n))) for n, c in hashes_count] or [1]
    db = formats.IBFFile(formats.IBF(n_bins, bin_size, hf, ibf.data), k, w, max(per_bin), hashes_count, bin_map)
    formats.write_ibf(path, db)
    return genomes


def make_reads(rng, genomes, n, w):
    gl = list(genomes.values())
    out = []
    for i in range(n):
        kind = rng.random()
        ln = rng.choice((w - 1, w, w + 1, 50, 100, 150, 151, 260))
        if kind < 0.55 and gl:
            g = rng.choice(gl)
            if len(g) > ln:
                p = rng.randrange(0, len(g) - ln)
                s = bytearray(g[p : p + ln])
            else:
                s = bytearray(g)
            if rng.random() < 0.5:
                s = bytearray(bytes(s).translate(COMP)[::-1])
            for _ in range(rng.choice((0, 0, 1, 3, 8))):
                s[rng.randrange(len(s))] = rng.choice(b"ACGT")
            if rng.random() < 0.2:
                for _ in range(rng.randrange(1, 6)):
                    s[rng.randrange(len(s))] = rng.choice(IUPAC)
            s = bytes(s)
        elif kind < 0.8:
            s = _seq(rng, max(1, ln))
        elif kind < 0.9:
            s = (_seq(rng, rng.choice((1, 2, 3, 5, 13))) * 300)[: max(1, ln)]
        else:
            s = _seq(rng, max(1, ln), b"ACGTN")
        out.append((("q%d_%d" % (i, rng.randrange(99))).encode(), s))
    return out


def write_fastq(path, recs):
    with open(path, "wb") as f:
        for rid, s in recs:
            f.write(b"@" + rid + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


def make_case(seed, tmp):
    """Writes the databases and reads of one random case; returns (ganon-classify argv without -o/-t, description,
    a function computing the oracle's sorted `.all` / `.unc` lines)."""
    rng = random.Random(seed)
    k = rng.choice((8, 12, 19, 19, 21, 27, 31, 32))
    w = k + rng.choice((0, 1, 4, 12, 30))
    n_levels = rng.choice((1, 1, 2))
    n_filters = [rng.choice((1, 1, 2)) for _ in range(n_levels)]
    labels, ibfs, cut, genomes = [], [], [], {}
    for li, nf in enumerate(n_filters):
        for fi in range(nf):
            p = os.path.join(tmp, "s%d_l%d_f%d.ibf" % (seed, li, fi))
            genomes.update(make_db(rng, p, k, w))
            ibfs.append(p)
            labels.append("L%d" % li)
            cut.append(rng.choice((0.0, 0.05, 0.2, 0.5, 0.75, 1.0)))
    relf = [rng.choice((0.0, 0.1, 0.5, 1.0)) for _ in range(n_levels)]
    fprq = [rng.choice((1.0, 1.0, 0.5, 1e-2, 1e-5)) for _ in range(n_levels)]
    paired = rng.random() < 0.4
    reads1 = make_reads(rng, genomes, 120, w)
    f1 = os.path.join(tmp, "s%d.1.fq" % seed)
    write_fastq(f1, reads1)
    args = ["-i", ",".join(ibfs), "-y", ",".join(labels), "-c", ",".join(map(str, cut)), "-d", ",".join(map(str, relf)), "-f", ",".join(map(str, fprq)), "-a", "-u", "-s"]
    if paired:
        reads2 = make_reads(rng, genomes, 120, w)
        f2 = os.path.join(tmp, "s%d.2.fq" % seed)
        write_fastq(f2, reads2)
        args = ["-p", f1 + "," + f2] + args
        reads = [(a[0], a[1], b[1]) for a, b in zip(reads1, reads2)]
    else:
        args = ["-r", f1] + args
        reads = [(a[0], a[1], None) for a in reads1]
    desc = "seed %d k=%d w=%d filters=%s cut=%s relf=%s fprq=%s paired=%s" % (seed, k, w, n_filters, cut, relf, fprq, paired)

    def oracle_lines():
        # the oracle, level by level (unclassified reads move on, GC.cpp:811-830)
        left, got_all, fi = reads, [], 0
        for li, nf in enumerate(n_filters):
            filters = [O.OracleFilter.from_ibf_file(formats.read_ibf(ibfs[fi + j]), cut[fi + j]) for j in range(nf)]
            fi += nf
            res = O.classify_level(filters, left, relf[li], fprq[li])
            got_all += O.all_lines(res)
            left = [r for r, x in zip(left, res) if not x["matches"]]
        return sorted(got_all), sorted(r[0].decode() for r in left)

    return args, desc, oracle_lines


def read_sorted(path):
    if not os.path.exists(path):
        return []
    with open(path) as f:
        return sorted(l.rstrip("\n") for l in f)


def run_case(seed, tmp):
    """The reference binary against the oracle on one random case: (ok, description)."""
    args, desc, oracle_lines = make_case(seed, tmp)
    out = os.path.join(tmp, "s%d_ref" % seed)
    pr = subprocess.run([REF_BIN] + args + ["-o", out, "-t", "2", "--quiet"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if pr.returncode != 0:
        return False, desc + " reference failed: " + pr.stderr[-200:]
    want_all, want_unc = read_sorted(out + ".all"), read_sorted(out + ".unc")
    got_all, got_unc = oracle_lines()
    ok = got_all == want_all and got_unc == want_unc
    return ok, desc + (" all=%d unc=%d" % (len(want_all), len(want_unc)))


if __name__ == "__main__":
    import tempfile

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(first, first + n):
            ok, desc = run_case(seed, tmp)
            if not ok:
                bad += 1
                print("MISMATCH", desc)
            for f in os.listdir(tmp):
                os.remove(os.path.join(tmp, f))
    print("%d cases, %d mismatches" % (n, bad))
    sys.exit(1 if bad else 0)
