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


def make_hibf(rng, path, k, w):
    """A random HIBF in the raptor layout: up to 3 levels, sub-IBFs of 1..200 bins, merged bins, user bins split over
    1..4 technical bins (also across 64-bin words), names that go through the reference's mangling (GC.cpp:916-928)."""
    hf = rng.randint(1, 5)
    ibfs, nxt, pos, genomes, bin_path = [], [], [], [], []

    def user_bin():
        u = len(genomes)
        genomes.append(_seq(rng, rng.choice((200, 400, 900))))
        name = ["GCF_%06d|||%d" % (u, rng.randrange(9)), "s__Sp---nr---%d" % u, "plain%d" % u][u % 3]
        bin_path.append(["/x/y/%s.minimiser" % name] + (["/x/extra%d.fa" % u] if rng.random() < 0.2 else []))
        return u

    def make_ibf(depth, n_bins, bin_size):
        idx = len(ibfs)
        ibfs.append(None)
        nxt.append(None)
        pos.append(None)
        o = O.OracleIBF(n_bins, bin_size, hf)
        my_nxt, my_pos, contained = [idx] * n_bins, [0] * n_bins, []
        b = 0
        while b < n_bins:
            if depth < 2 and rng.random() < (0.12 if depth == 0 else 0.06) and len(ibfs) < 12:
                child, hs = make_ibf(depth + 1, rng.choice((1, 5, 20, 64, 70, 130)), rng.choice((1009, 4099, 20011)))
                for h in hs:
                    o.emplace(int(h), b)
                my_nxt[b], my_pos[b] = child, -1
                contained.extend(hs)
                b += 1
            else:
                nb = min(rng.choice((1, 1, 1, 2, 3, 4)), n_bins - b)
                u = user_bin()
                hs = sorted(set(int(x) for x in O.minimiser_hash(genomes[u], k, w)))
                for i, h in enumerate(hs):
                    o.emplace(h, b + i % nb)
                for j in range(nb):
                    my_pos[b + j] = u
                contained.extend(hs)
                b += nb
        ibfs[idx] = formats.IBF(n_bins, bin_size, hf, o.data.copy())
        nxt[idx], pos[idx] = my_nxt, my_pos
        return idx, contained

    make_ibf(0, rng.choice((3, 40, 64, 66, 130)), rng.choice((4099, 20011, 65537)))
    db = formats.HIBFFile(w, k, ibfs, nxt, ["f%d" % i for i in range(len(genomes))], pos, bin_path, fpr=rng.choice((0.05, 0.01, 0.3)))
    formats.write_hibf(path, db)
    return db, {i: g for i, g in enumerate(genomes)}


def run_hibf_case(seed, tmp):
    """`ganon-classify --hibf` against the oracle's HIBF traversal on one random case: (ok, description)."""
    rng = random.Random(seed)
    k = rng.choice((12, 19, 19, 21, 31))
    w = k + rng.choice((0, 4, 12))
    path = os.path.join(tmp, "h%d.hibf" % seed)
    db, genomes = make_hibf(rng, path, k, w)
    cut, relf, fprq = rng.choice((0.0, 0.05, 0.3, 0.75)), rng.choice((0.0, 0.2, 1.0)), rng.choice((1.0, 0.5, 1e-3))
    reads1 = make_reads(rng, genomes, 100, w)
    f1 = os.path.join(tmp, "h%d.fq" % seed)
    write_fastq(f1, reads1)
    out = os.path.join(tmp, "h%d_ref" % seed)
    pr = subprocess.run([REF_BIN, "--hibf", "-r", f1, "-i", path, "-c", str(cut), "-d", str(relf), "-f", str(fprq), "-a", "-u", "-o", out, "-t", "2", "--quiet"],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    desc = "hibf seed %d k=%d w=%d ibfs=%d user bins=%d cut=%s relf=%s fprq=%s" % (seed, k, w, len(db.ibfs), len(genomes), cut, relf, fprq)
    if pr.returncode != 0:
        return False, desc + " reference failed: " + pr.stderr[-200:]
    oh = O.OracleHIBF([O.OracleIBF(i.bins, i.bin_size, i.hash_funs, i.data) for i in db.ibfs], db.next_ibf_id, db.bin_to_user, len(db.bin_path))
    tmap = {}
    for u, paths in enumerate(db.bin_path):
        for p_ in paths:
            tmap.setdefault(formats.hibf_target_name(p_), []).append(u)
    targets = list(tmap)
    filt = O.OracleFilter(oh, targets, [tmap[t] for t in targets], [db.fpr] * len(targets), cut, k, w)
    res = O.classify_level([filt], [(a[0], a[1], None) for a in reads1], relf, fprq)
    want_all, want_unc = read_sorted(out + ".all"), read_sorted(out + ".unc")
    ok = O.all_lines(res) == want_all and sorted(r["id"].decode() for r in res if not r["matches"]) == want_unc
    return ok, desc + (" all=%d unc=%d" % (len(want_all), len(want_unc)))


if __name__ == "__main__":
    import tempfile

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(first, first + n):
            ok, desc = (run_hibf_case if os.environ.get("FUZZ_HIBF") else run_case)(seed, tmp)
            if not ok:
                bad += 1
                print("MISMATCH", desc)
            for f in os.listdir(tmp):
                os.remove(os.path.join(tmp, f))
    print("%d cases, %d mismatches" % (n, bad))
    sys.exit(1 if bad else 0)
