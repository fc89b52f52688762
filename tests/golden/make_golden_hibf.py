#!/usr/bin/env python
"""Golden fixtures for the HIBF path, made by RUNNING THE UNMODIFIED REFERENCE (`ganon-classify --hibf`).

raptor (the HIBF builder) is not available offline, so the .hibf is synthesised in the raptor 3.0.1 index layout
(ganon_b200/formats.py:write_hibf, layout verified against the reference's reader in SURVEY.md §8c): three levels,
split user bins, merged bins, name mangling of GC.cpp:916-928.  Outputs (committed):
  synth.hibf.gz, expected/hibf_*.{all,rep,sta,unc}, scenarios_hibf.json
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

CLASSIFY = os.path.join(ROOT, "oracle/_ref/ganon-classify")
K, W, H = 19, 31, 3


def build(rng):
    """Returns (HIBFFile, genomes per user bin)."""
    ibfs, nxt, pos = [], [], []
    user_genomes, bin_path = [], []

    def new_user_bin():
        u = len(user_genomes)
        user_genomes.append(bytes(rng.choice(b"ACGT") for _ in range(2500)))
        name = ["GCF_%06d|||1" % u, "s__Species---number---%d" % u, "plain%d" % u][u % 3]
        bin_path.append(["/data/build/%s.minimiser" % name])
        return u

    def make_ibf(depth, n_bins, bin_size):
        idx = len(ibfs)
        ibfs.append(None)
        nxt.append(None)
        pos.append(None)
        o = O.OracleIBF(n_bins, bin_size, H)
        my_nxt, my_pos = [idx] * n_bins, [0] * n_bins
        contained = []  # hashes of everything below this ibf
        b = 0
        while b < n_bins:
            kind = rng.random()
            if depth < 2 and kind < 0.12:
                # merged bin -> child IBF
                child_bins = rng.choice((40, 64, 70, 130))
                child, child_hashes = make_ibf(depth + 1, child_bins, rng.choice((20011, 30011)))
                for h in child_hashes:
                    o.emplace(int(h), b)
                my_nxt[b], my_pos[b] = child, -1
                contained.extend(child_hashes)
                b += 1
            else:
                nb = min(rng.choice((1, 1, 1, 1, 2, 3)), n_bins - b)
                u = new_user_bin()
                hs = sorted(set(int(x) for x in O.minimiser_hash(user_genomes[u], K, W)))
                for i, h in enumerate(hs):
                    o.emplace(h, b + i % nb)
                for j in range(nb):
                    my_pos[b + j] = u
                contained.extend(hs)
                b += nb
        ibfs[idx] = formats.IBF(n_bins, bin_size, H, o.data.copy())
        nxt[idx], pos[idx] = my_nxt, my_pos
        return idx, contained

    make_ibf(0, 70, 60013)
    db = formats.HIBFFile(W, K, ibfs, nxt, ["f%d" % i for i in range(len(user_genomes))], pos, bin_path, fpr=0.05)
    return db, user_genomes


def main():
    rng = random.Random(4242)
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    db, genomes = build(rng)
    path = os.path.join(tmp, "synth.hibf")
    formats.write_hibf(path, db)
    print("ibfs:", len(db.ibfs), "user bins:", len(genomes), "merged:", sum(1 for p in db.bin_to_user for x in p if x < 0))
    # reads: from the user bins' genomes (with errors), random, and the committed adversarial single-end set
    recs = []
    for i in range(500):
        u = rng.randrange(len(genomes))
        g = genomes[u]
        L = rng.choice((60, 100, 150, 250))
        p = rng.randrange(0, len(g) - L)
        s = bytearray(g[p : p + L])
        for _ in range(rng.choice((0, 0, 1, 3, 8))):
            s[rng.randrange(L)] = rng.choice(b"ACGTN")
        recs.append((b"hr_u%d_%d" % (u, i), bytes(s)))
    for i in range(60):
        recs.append((b"hr_random_%d" % i, bytes(rng.choice(b"ACGT") for _ in range(150))))
    with open(os.path.join(HERE, "reads.hibf.fq"), "wb") as f:
        for rid, s in recs:
            f.write(b"@" + rid + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    P = "{golden}/"
    scenarios = {
        "hibf_all": ["--hibf", "-r", P + "reads.hibf.fq", "-i", "{tmp}/synth.hibf", "-c", "0", "-d", "1", "-a", "-u", "-z"],
        "hibf_cut": ["--hibf", "-r", P + "reads.hibf.fq", "-i", "{tmp}/synth.hibf", "-c", "0.3", "-d", "0.2", "-f", "0.001", "-a", "-u", "-z"],
        "hibf_pe_adv": ["--hibf", "-p", P + "reads.1.fq," + P + "reads.2.fq", "-r", P + "reads.se.fq", "-i", "{tmp}/synth.hibf", "-c", "0.05", "-d", "1", "-a", "-u", "-z"],
    }
    with open(os.path.join(HERE, "scenarios_hibf.json"), "w") as f:
        json.dump(scenarios, f, indent=1)
    for name, args in scenarios.items():
        argv = [a.format(golden=HERE, tmp=tmp) for a in args]
        pre = os.path.join(HERE, "expected", name)
        subprocess.check_call([CLASSIFY] + argv + ["-o", pre, "-t", "4", "--quiet"])
        for fn in glob.glob(pre + ".*"):
            with open(fn) as fh:
                lines = sorted(fh.readlines())
            with open(fn, "w") as fh:
                fh.writelines(lines)
            print(fn, len(lines))
    with open(path, "rb") as fi, gzip.GzipFile(os.path.join(HERE, "synth.hibf.gz"), "wb", 9, mtime=0) as fo:
        shutil.copyfileobj(fi, fo)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
