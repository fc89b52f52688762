#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Needs /root/reference (genome files of the reference's own test-suite) and the reference binaries
built by `make -C oracle ref` (oracle/_ref/ganon-build, oracle/_ref/ganon-classify).  It is run once,
in the build container; the outputs are committed, so tests never need /root/reference.

What is produced
  real4.ibf.gz     flat IBF built by the reference ganon-build from the 4 genomes of
                   tests/ganon/data/build-custom/files (0.5 MB filter => 64 bins, h=5, multi-bin targets)
  real4b.ibf.gz    same genomes, 0.25 MB filter (different counts for the same targets)
  synth.ibf.gz     hand-written flat IBF (130 bins / 192 technical, non-power-of-two bin_size, h=3,
                   2-3 bin split targets, unsorted bin map) holding planted random genomes
  real4.tax        a small taxonomy over the 4 targets (for LCA)
  reads.{1,2}.fq   paired reads: from the real genomes, from the planted genomes, random, adversarial
  reads.se.fq      single-end reads with lengths 20..300 (some shorter than the window)
  reads.fa         FASTA subset (multi-line)
  expected/<scenario>.*  the reference's .all/.one/.rep/.unc/.sta for each scenario in SCENARIOS
  scenarios.json   scenario -> argv template
"""
import glob
import gzip
import json
import os
import random
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ganon_b200 import formats  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
BUILD = os.path.join(ROOT, "oracle/_ref/ganon-build")
CLASSIFY = os.path.join(ROOT, "oracle/_ref/ganon-classify")
K, W = 19, 31

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def read_fasta_gz(path):
    seqs, cur = [], []
    with gzip.open(path, "rb") as f:
        for ln in f:
            if ln.startswith(b">"):
                if cur:
                    seqs.append(b"".join(cur))
                cur = []
            else:
                cur.append(ln.strip().upper())
    if cur:
        seqs.append(b"".join(cur))
    return seqs


def mutate(rng, s: bytes, rate: float) -> bytes:
    b = bytearray(s)
    for i in range(len(b)):
        if rng.random() < rate:
            b[i] = rng.choice(b"ACGT")
    return bytes(b)


def sample_pair(rng, genome: bytes, rlen=150, insert=300, err=0.01):
    p = rng.randrange(0, len(genome) - insert)
    frag = genome[p : p + insert]
    if rng.random() < 0.5:
        frag = revcomp(frag)
    return mutate(rng, frag[:rlen], err), mutate(rng, revcomp(frag)[:rlen], err)


def write_fastq(path, recs):
    with open(path, "wb") as f:
        for rid, s in recs:
            f.write(b"@" + rid + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


def main():
    rng = random.Random(20261017)
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)

    # ---------------------------------------------------------------- real4 via the reference builder
    files = sorted(glob.glob(REF + "/tests/ganon/data/build-custom/files/*.fna.gz"))
    targets = ["_".join(os.path.basename(f).split("_")[:2]) for f in files]
    with open(os.path.join(tmp, "in.tsv"), "w") as f:
        for p, t in zip(files, targets):
            f.write("%s\t%s\n" % (p, t))
    for name, size in (("real4", "0.5"), ("real4b", "0.25")):
        subprocess.check_call([BUILD, "-i", os.path.join(tmp, "in.tsv"), "-o", os.path.join(tmp, name + ".ibf"), "-k", str(K), "-w", str(W), "-f", size, "-t", "4", "--quiet"])
    genomes = {t: read_fasta_gz(p) for p, t in zip(files, targets)}
    with open(os.path.join(HERE, "real4.tax"), "w") as f:
        # node <tab> parent <tab> rank <tab> name ; two genera under one family under root "1"
        f.write("1\t0\troot\troot\n")
        f.write("F1\t1\tfamily\tfam one\n")
        f.write("G1\tF1\tgenus\tgenus one\n")
        f.write("G2\tF1\tgenus\tgenus two\n")
        for t, g in zip(targets, ("G1", "G1", "G2", "1")):
            f.write("%s\t%s\tassembly\tname of %s\n" % (t, g, t))

    # ---------------------------------------------------------------- synthetic flat IBF with planted genomes
    n_bins, bin_size, hf = 130, 100003, 3
    ibf = O.OracleIBF(n_bins, bin_size, hf)
    nprng = np.random.default_rng(7)
    # background noise: ~12% of all bits
    noise = nprng.integers(0, 1 << 63, size=ibf.data.size, dtype=np.uint64) & nprng.integers(0, 1 << 63, size=ibf.data.size, dtype=np.uint64) & nprng.integers(0, 1 << 63, size=ibf.data.size, dtype=np.uint64)
    # padding bins 130..191 must stay zero (IBF.hpp:238-240): mask the third word of every row
    noise = noise.reshape(bin_size, ibf.bin_words)
    noise[:, 2] &= np.uint64((1 << (n_bins - 128)) - 1)
    ibf.data[:] = noise.reshape(-1)
    synth_genomes, bin_map, hashes_count = {}, [], []
    b = 0
    t = 0
    while b < n_bins:
        nb = min(rng.choice((1, 1, 1, 2, 3)), n_bins - b)
        name = "S%03d" % t
        g = bytes(rng.choice(b"ACGT") for _ in range(6000))
        synth_genomes[name] = [g]
        hs = O.minimiser_hash(g, K, W)
        uniq = sorted(set(int(x) for x in hs))
        for i, h in enumerate(uniq):
            ibf.emplace(h, b + i % nb)
        for j in range(nb):
            bin_map.append((b + j, name))
        hashes_count.append((name, len(uniq)))
        b += nb
        t += 1
    rng.shuffle(bin_map)
    max_hashes_bin = max(-(-c // sum(1 for bb, tt in bin_map if tt == n)) for n, c in hashes_count)
    db = formats.IBFFile(formats.IBF(n_bins, bin_size, hf, ibf.data), K, W, max_hashes_bin, hashes_count, bin_map)
    formats.write_ibf(os.path.join(tmp, "synth.ibf"), db)

    # ---------------------------------------------------------------- reads
    pairs = []
    allg = [(t, s) for t, ss in genomes.items() for s in ss if len(s) > 1000]
    for i in range(300):
        t, s = rng.choice(allg)
        a, b2 = sample_pair(rng, s)
        pairs.append((("real_%s_%d" % (t, i)).encode(), a, b2))
    sg = [(t, s) for t, ss in synth_genomes.items() for s in ss]
    for i in range(300):
        t, s = rng.choice(sg)
        a, b2 = sample_pair(rng, s, err=rng.choice((0.0, 0.01, 0.05)))
        pairs.append((("synth_%s_%d some description" % (t, i)).encode(), a, b2))
    for i in range(100):
        pairs.append((("random_%d" % i).encode(), bytes(rng.choice(b"ACGT") for _ in range(150)), bytes(rng.choice(b"ACGT") for _ in range(150))))
    # adversarial: homopolymers, short-period repeats, AT-only, N / IUPAC / lower case, odd lengths, short mates
    adv = []
    for c in b"ACGTN":
        adv.append(bytes([c]) * 150)
    for period in (2, 3, 4, 5, 7, 12, 13, 19, 20, 31, 39):
        unit = bytes(rng.choice(b"ACGT") for _ in range(period))
        adv.append((unit * 200)[:150])
    for _ in range(20):
        adv.append(bytes(rng.choice(b"AT") for _ in range(150)))
    for _ in range(30):
        t, s = rng.choice(allg + sg)
        p = rng.randrange(0, len(s) - 200)
        r = bytearray(s[p : p + rng.randrange(31, 200)])
        for _ in range(rng.randrange(1, 8)):
            r[rng.randrange(len(r))] = rng.choice(b"NRYSWKMBDHVUnacgtryswkmbdhvu")
        adv.append(bytes(r))
    for ln in (10, 18, 19, 30, 31, 32, 43, 44, 200, 300):
        t, s = rng.choice(allg + sg)
        p = rng.randrange(0, len(s) - 400)
        adv.append(s[p : p + ln])
    for i, a in enumerate(adv):
        mate = adv[(i * 7 + 3) % len(adv)]
        pairs.append((("adv_%d" % i).encode(), a, mate))
    rng.shuffle(pairs)
    write_fastq(os.path.join(HERE, "reads.1.fq"), [(i + b"/1", a) for i, a, _ in pairs])
    write_fastq(os.path.join(HERE, "reads.2.fq"), [(i + b"/2", b2) for i, _, b2 in pairs])
    se = []
    for i in range(400):
        t, s = rng.choice(allg + sg)
        ln = rng.choice((20, 30, 31, 35, 50, 75, 100, 150, 151, 250, 300))
        p = rng.randrange(0, len(s) - 400)
        r = s[p : p + ln]
        if rng.random() < 0.5:
            r = revcomp(r)
        se.append((("se_%s_%d" % (t, i)).encode(), mutate(rng, r, 0.02)))
    write_fastq(os.path.join(HERE, "reads.se.fq"), se)
    with open(os.path.join(HERE, "reads.fa"), "wb") as f:
        for rid, s in se[:150]:
            f.write(b">" + rid + b" fasta\n")
            for o in range(0, len(s), 60):
                f.write(s[o : o + 60] + b"\n")

    # ---------------------------------------------------------------- scenarios run through the reference
    P = "{golden}/"
    scenarios = {
        # every non-zero per-target count exposed
        "pe_real4_all": ["-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/real4.ibf", "-c", "0", "-d", "1", "-a", "-u", "-z"],
        # ganon (python CLI) defaults incl. fpr-query, with taxonomy / LCA
        "pe_real4_defaults_lca": ["-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/real4.ibf", "-x", P + "real4.tax", "-c", "0.75", "-d", "0.1", "-f", "1e-5", "-a", "-l", "-u", "-z"],
        "pe_real4_lowcut_lca": ["-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/real4.ibf", "-x", P + "real4.tax", "-c", "0.1", "-d", "0.5", "-f", "0.01", "-a", "-l", "-u", "-z"],
        "se_synth": ["-r", P + "reads.se.fq", "-i", "{tmp}/synth.ibf", "-c", "0.25", "-d", "0.5", "-f", "0.001", "-a", "-u", "-z"],
        "se_synth_all": ["-r", P + "reads.se.fq", "-i", "{tmp}/synth.ibf", "-c", "0", "-d", "1", "-a", "-u", "-z"],
        "pe_synth_all": ["-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/synth.ibf", "-c", "0", "-d", "1", "-a", "-u", "-z"],
        # binary defaults (rel-cutoff 0.2, rel-filter 0, fpr-query 1), FASTA input
        "fa_real4_bindefaults": ["-r", P + "reads.fa", "-i", "{tmp}/real4.ibf", "-a", "-u", "-z"],
        # two filters on one level sharing target names (stale-min quirk), per-filter cutoffs
        "pe_two_filters": ["-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/real4.ibf,{tmp}/real4b.ibf", "-c", "0.1,0.3", "-d", "0.2", "-a", "-u", "-z"],
        "pe_three_filters": ["-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/real4b.ibf,{tmp}/synth.ibf,{tmp}/real4.ibf", "-c", "0.05", "-d", "0.3", "-f", "0.5", "-a", "-u", "-z"],
        # two hierarchy levels, split and single outputs; mixed single + paired input
        "hier_two_levels": ["-r", P + "reads.se.fq", "-p", P + "reads.1.fq," + P + "reads.2.fq", "-i", "{tmp}/synth.ibf,{tmp}/real4.ibf", "-y", "1_first,2_second", "-c", "0.3,0.1", "-d", "0.1,0.5", "-a", "-u", "-z"],
        "hier_two_levels_single": ["-r", P + "reads.se.fq", "-i", "{tmp}/real4.ibf,{tmp}/synth.ibf", "-y", "B,A", "-c", "0.5", "-d", "0", "-f", "1,0.1", "-a", "-u", "-z", "-s"],
    }
    with open(os.path.join(HERE, "scenarios.json"), "w") as f:
        json.dump(scenarios, f, indent=1)
    for name, args in scenarios.items():
        argv = [a.format(golden=HERE, tmp=tmp) for a in args]
        pre = os.path.join(HERE, "expected", name)
        subprocess.check_call([CLASSIFY] + argv + ["-o", pre, "-t", "4", "--quiet"])
        for fn in glob.glob(pre + ".*"):
            # the reference's line order is not deterministic -> store sorted (the .sta/.rep have '#'/header lines that sort fine too)
            with open(fn) as fh:
                lines = sorted(fh.readlines())
            with open(fn, "w") as fh:
                fh.writelines(lines)
    for n in ("real4", "real4b", "synth"):
        with open(os.path.join(tmp, n + ".ibf"), "rb") as fi, gzip.GzipFile(os.path.join(HERE, n + ".ibf.gz"), "wb", 9, mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    shutil.rmtree(tmp)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
