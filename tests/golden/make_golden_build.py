#!/usr/bin/env python
"""Golden fixtures for the build-side arithmetic, made by RUNNING THE UNMODIFIED REFERENCE `ganon-build`
(oracle/_ref/ganon-build, compiled from /root/reference by oracle/Makefile) on random genome sets with different
parameters.  From every `.ibf` it writes only the header is kept: the IBFConfig the reference chose, the per-target
minimiser counts and the bin map.  Output (committed): tests/golden/build_cases.json."""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ganon_b200 import formats  # noqa: E402

BUILD = os.path.join(ROOT, "oracle/_ref/ganon-build")


def main():
    rng = random.Random(77)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for ci in range(48):
            n_targets = rng.choice((1, 2, 5, 12, 40))
            k = rng.choice((19, 19, 21, 27))
            w = k + rng.choice((0, 4, 12))
            genomes = {}
            with open(os.path.join(tmp, "in.tsv"), "w") as tsv:
                for t in range(n_targets):
                    name = "tg%d_%d" % (ci, t)
                    L = rng.choice((150, 800, 3000, 12000, 40000))
                    seq = "".join(rng.choice("ACGT") for _ in range(L))
                    genomes[name] = seq
                    p = os.path.join(tmp, name + ".fa")
                    with open(p, "w") as f:
                        f.write(">%s\n%s\n" % (name, seq))
                    tsv.write("%s\t%s\n" % (p, name))
            args = ["-k", str(k), "-w", str(w)]
            params = dict(k=k, w=w, max_fp=0.05, filter_size=0.0, hash_functions=0, mode="avg")
            kind = rng.random()
            if kind < 0.45:
                params["max_fp"] = rng.choice((0.05, 0.01, 0.001, 0.2))
                args += ["-p", str(params["max_fp"])]
            elif kind < 0.8:
                params["filter_size"] = rng.choice((0.05, 0.2, 1.0, 3.0))
                args += ["-f", str(params["filter_size"])]
            if rng.random() < 0.5:
                params["hash_functions"] = rng.choice((1, 2, 3, 4, 5))
                args += ["-s", str(params["hash_functions"])]
            if rng.random() < 0.5:
                params["mode"] = rng.choice(("smaller", "smallest", "faster", "fastest", "avg"))
                args += ["-j", params["mode"]]
            out = os.path.join(tmp, "o.ibf")
            pr = subprocess.run([BUILD, "-i", os.path.join(tmp, "in.tsv"), "-o", out, "-t", "2", "--quiet"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            if pr.returncode != 0 or not os.path.exists(out):
                print("case", ci, "failed:", pr.stderr[-200:])
                continue
            db = formats.read_ibf(out, load_data=False)
            cases.append(dict(params=params, genomes=genomes if sum(map(len, genomes.values())) < 20000 else None, hashes_count=db.hashes_count,
                              bin_map=sorted(db.bin_map), n_bins=db.ibf.bins, bin_size_bits=db.ibf.bin_size, hash_functions=db.ibf.hash_funs,
                              max_hashes_bin=db.max_hashes_bin, max_fp=db.max_fp, true_max_fp=db.true_max_fp, true_avg_fp=db.true_avg_fp))
            os.remove(out)
    with open(os.path.join(HERE, "build_cases.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
