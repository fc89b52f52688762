"""`ganon-build` drop-in against the UNMODIFIED reference builder on random option sets (GanonBuild::Config::validate,
Config.hpp:33-107, and the parameter choice of GanonBuild.cpp:290-618 in its degenerate corners: --max-fp 1, --filter-size with
a handful of hashes, --min-length above every sequence ...): same accept / reject decision, same first message on stderr, and
for accepted runs the same IBF parameters.  The device calls are answered by the oracle.  The reference computes in IEEE
doubles without traps (log 0, x / 0, nan scores); the first version of the restatement raised Python exceptions there."""
import contextlib
import io
import os
import random
import subprocess

import pytest

from ganon_b200 import build as B
from ganon_b200 import formats
from tests.build_util import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BUILD = os.path.join(ROOT, "oracle", "_ref", "ganon-build")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_BUILD), reason="oracle/_ref not built (only in the build container)")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("buildcli"))
    rng = random.Random(3)
    p = {}
    lines = []
    for i in range(3):
        fa = os.path.join(d, "g%d.fa" % i)
        with open(fa, "w") as f:
            f.write(">g%d\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(rng.choice((189, 700, 2500))))))
        lines.append("%s\tT%d\n" % (fa, i % 2))
    p["tsv"] = os.path.join(d, "in.tsv")
    open(p["tsv"], "w").write("".join(lines))
    p["tsv_missing_files"] = os.path.join(d, "bad.tsv")
    open(p["tsv_missing_files"], "w").write("%s\tA\n" % os.path.join(d, "nofile.fa"))
    p["empty"] = os.path.join(d, "empty.tsv")
    open(p["empty"], "w").close()
    p["missing"] = os.path.join(d, "missing.tsv")
    p["tmpdir"] = os.path.join(d, "tmpdir")
    os.makedirs(p["tmpdir"])
    p["nodir"] = os.path.join(d, "nodir")
    p["dir"] = d
    return p


def _argv(rng, p):
    a = []
    if rng.random() < 0.9:
        a += ["-i", rng.choice([p["tsv"], p["tsv"], p["tsv"], p["tsv_missing_files"], p["empty"], p["missing"]])]
    if rng.random() < 0.9:
        a += ["-o", os.path.join(p["dir"], "out.ibf")]
    if rng.random() < 0.6:
        a += ["-k", rng.choice(["4", "19", "32", "33", "10"])]
    if rng.random() < 0.6:
        a += ["-w", rng.choice(["4", "19", "31", "40", "9"])]
    if rng.random() < 0.4:
        a += ["-s", rng.choice(["0", "1", "5", "6"])]
    if rng.random() < 0.5:
        a += ["-p", rng.choice(["0", "0.05", "0.5", "1", "0.001"])]
    if rng.random() < 0.3:
        a += ["-f", rng.choice(["0", "0.1", "1"])]
    if rng.random() < 0.3:
        a += ["-j", rng.choice(["avg", "smaller", "smallest", "faster", "fastest", "bogus"])]
    if rng.random() < 0.2:
        a += ["-y", rng.choice(["0", "50", "300"])]
    if rng.random() < 0.2:
        a += ["-m", rng.choice([p["tmpdir"], p["nodir"]])]
    return a


def _first_message(text):
    keep = [l for l in text.strip().splitlines() if not l.startswith(("ganon-build processed", " - ", "---"))]
    return keep[:1]


@pytest.mark.parametrize("first", range(0, 300, 50))
def test_builder_options_match_the_reference(files, first):
    accepted = 0
    short = {v[0]: k for k, v in B._BUILD_OPTS.items() if v[0]}
    for seed in range(first, first + 50):
        a = _argv(random.Random(seed), files)
        ref_out = os.path.join(files["dir"], "ref.ibf")
        ra = [ref_out if a[i - 1] == "-o" else x for i, x in enumerate(a)]
        pr = subprocess.run([REF_BUILD] + ra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if pr.returncode < 0:
            continue  # the reference itself dies (std::bad_alloc when --filter-size meets a target set without hashes)
        cfg = B.GanonBuildConfig()
        for i in range(0, len(a), 2):
            _s, kind, attr = B._BUILD_OPTS[short[a[i][1:]]]
            setattr(cfg, attr, kind(a[i + 1]))
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            ok = B.run_build(cfg, backend=OracleBackend())
        assert bool(ok) == (pr.returncode == 0), (seed, a, pr.stderr, err.getvalue())
        if not ok:
            assert _first_message(err.getvalue()) == _first_message(pr.stderr), (seed, a)
            continue
        accepted += 1
        x, y = formats.read_ibf(cfg.output_file), formats.read_ibf(ref_out)
        assert (x.ibf.bins, x.ibf.bin_size, x.ibf.hash_funs, x.max_hashes_bin, x.kmer_size, x.window_size) == (y.ibf.bins, y.ibf.bin_size, y.ibf.hash_funs, y.max_hashes_bin, y.kmer_size, y.window_size), (seed, a)
        assert (x.max_fp, x.true_max_fp) == (y.max_fp, y.true_max_fp) and x.true_avg_fp == pytest.approx(y.true_avg_fp, rel=1e-12, nan_ok=True), (seed, a)
        assert sorted(x.hashes_count) == sorted(y.hashes_count), (seed, a)
    assert accepted >= 3


@pytest.mark.parametrize("table", ["{fa}\tA\t\n", "{fa}\t\n", "{fa}\tA\tseqid\n{fa2}\tB\n", "{fa}\n{fa2}\tB", "{missing}\tA\n{fa}\tB\n"])
def test_input_tables_are_read_like_the_reference(files, table, tmp_path):
    """parse_input_file (GanonBuild.cpp:86-137): a trailing tab adds no field, a line with three fields names no target, a file
    that does not exist is skipped; same targets and hash counts as the reference builder."""
    fa, fa2 = os.path.join(files["dir"], "g0.fa"), os.path.join(files["dir"], "g1.fa")
    tsv = str(tmp_path / "in.tsv")
    open(tsv, "w").write(table.format(fa=fa, fa2=fa2, missing=files["missing"]))
    ref_out, out = str(tmp_path / "ref.ibf"), str(tmp_path / "mine.ibf")
    pr = subprocess.run([REF_BUILD, "-i", tsv, "-o", ref_out, "-k", "19", "-w", "31", "-p", "0.05", "--quiet"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    ok = B.run_build(B.GanonBuildConfig(input_file=tsv, output_file=out, quiet=True), backend=OracleBackend())
    assert bool(ok) == (pr.returncode == 0), pr.stderr
    if ok:
        assert sorted(formats.read_ibf(out).hashes_count) == sorted(formats.read_ibf(ref_out).hashes_count)
