"""`ganon-build` drop-in (ganon_b200/build.py:run_build) with the oracle standing in for the device: the file it writes has
the parameters the reference builder chose for the same input, holds every minimiser of every target in one of the target's
bins, loads in the unmodified reference `ganon-classify` and classifies reads to the genomes they come from."""
import json
import os
import subprocess

import numpy as np
import pytest

from ganon_b200 import build as B
from ganon_b200 import formats
from oracle import oracle as O
from tests.build_util import OracleBackend

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [c for c in json.load(open(os.path.join(HERE, "golden", "build_cases.json"))) if c["genomes"]]
REF_CLASSIFY = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ganon-classify")


def _write_inputs(c, tmp):
    tsv = os.path.join(tmp, "in.tsv")
    with open(tsv, "w") as t:
        for name, seq in c["genomes"].items():
            p = os.path.join(tmp, name + ".fa")
            with open(p, "w") as f:
                f.write(">%s some description\n" % name)
                for o in range(0, len(seq), 70):
                    f.write(seq[o : o + 70] + "\n")
            t.write("%s\t%s\n" % (p, name))
    return tsv


def _cfg(c, tsv, out):
    p = c["params"]
    return B.GanonBuildConfig(input_file=tsv, output_file=out, kmer_size=p["k"], window_size=p["w"], max_fp=p["max_fp"], filter_size=p["filter_size"],
                              hash_functions=p["hash_functions"], mode=p["mode"], quiet=True)


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_built_file_has_reference_parameters_and_content(ci, tmp_path):
    c = CASES[ci]
    tsv = _write_inputs(c, str(tmp_path))
    out = str(tmp_path / "mine.ibf")
    assert B.run_build(_cfg(c, tsv, out), backend=OracleBackend())
    db = formats.read_ibf(out)
    assert (db.ibf.bins, db.ibf.bin_size, db.ibf.hash_funs, db.max_hashes_bin, db.kmer_size, db.window_size) == (c["n_bins"], c["bin_size_bits"], c["hash_functions"], c["max_hashes_bin"], c["params"]["k"], c["params"]["w"])
    assert (db.max_fp, db.true_max_fp) == (c["max_fp"], c["true_max_fp"]) and db.true_avg_fp == pytest.approx(c["true_avg_fp"], rel=1e-12)
    assert sorted(db.hashes_count) == sorted((t, n) for t, n in c["hashes_count"])
    bins_of = {}
    for b, t in db.bin_map:
        bins_of.setdefault(t, []).append(b)
    ref_bins = {}
    for b, t in c["bin_map"]:
        ref_bins[t] = ref_bins.get(t, 0) + 1
    assert {t: len(v) for t, v in bins_of.items()} == ref_bins
    # every distinct minimiser of a target sits in exactly the target's bins (and the filter holds nothing else:
    # the number of set bits is at most hashes x hash functions)
    o = O.OracleIBF(db.ibf.bins, db.ibf.bin_size, db.ibf.hash_funs, db.ibf.data)
    total = 0
    for t, seq in c["genomes"].items():
        hs = np.unique(O.minimiser_hash(seq.encode(), db.kmer_size, db.window_size))
        total += hs.size
        cnt = o.bulk_count(hs)
        assert int(cnt[bins_of[t]].sum()) >= hs.size
        for h in hs[:: max(1, hs.size // 40)]:
            assert o.bulk_count(np.array([h], dtype=np.uint64))[bins_of[t]].max() == 1
    bits = int(np.unpackbits(db.ibf.data.view(np.uint8)).sum())
    assert 0 < bits <= total * db.ibf.hash_funs
    # equal inputs give equal files
    out2 = str(tmp_path / "again.ibf")
    assert B.run_build(_cfg(c, tsv, out2), backend=OracleBackend())
    assert open(out, "rb").read() == open(out2, "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_CLASSIFY), reason="oracle/_ref/ganon-classify is not built")
def test_reference_classifier_accepts_the_built_file(tmp_path):
    c = max(CASES, key=lambda x: len(x["genomes"]))
    tsv = _write_inputs(c, str(tmp_path))
    out = str(tmp_path / "mine.ibf")
    assert B.run_build(_cfg(c, tsv, out), backend=OracleBackend())
    rng = np.random.default_rng(3)
    w = c["params"]["w"]
    recs, truth = [], {}
    for name, seq in c["genomes"].items():
        if len(seq) < 200:
            continue
        for j in range(4):
            p = int(rng.integers(0, len(seq) - 150))
            rid = "%s_r%d" % (name, j)
            recs.append("@%s\n%s\n+\n%s\n" % (rid, seq[p : p + 150], "I" * 150))
            truth[rid] = name
    fq = str(tmp_path / "r.fq")
    open(fq, "w").write("".join(recs))
    pre = str(tmp_path / "ref")
    pr = subprocess.run([REF_CLASSIFY, "-r", fq, "-i", out, "-c", "1", "-d", "0", "-a", "-o", pre, "-t", "2", "--quiet"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert pr.returncode == 0, pr.stderr
    hits = {}
    for line in open(pre + ".all"):
        rid, target, _k = line.rstrip("\n").split("\t")
        hits.setdefault(rid, set()).add(target)
    assert truth and all(truth[r] in hits.get(r, ()) for r in truth)


def test_command_line_validation(tmp_path, capsys):
    assert B.build_main([]) == 1
    assert B.build_main(["-o", str(tmp_path / "x.ibf")]) == 1
    assert "--input-file is mandatory" in capsys.readouterr().err
    tsv = tmp_path / "in.tsv"
    tsv.write_text("nofile.fa\tT\n")
    assert B.build_main(["-i", str(tsv), "-o", str(tmp_path / "x.ibf"), "-k", "40", "-w", "50"]) == 1
    assert "--kmer-size has to be <= 32" in capsys.readouterr().err
    assert B.build_main(["-i", str(tsv), "-o", str(tmp_path / "x.ibf"), "-j", "tiny"]) == 1
    assert "Invalid --mode" in capsys.readouterr().err
    assert B.build_main(["-i", str(tsv), "-o", str(tmp_path / "x.ibf"), "-w", "10", "-k", "19"]) == 1
    assert "--window-size has to be >= --kmer-size" in capsys.readouterr().err
