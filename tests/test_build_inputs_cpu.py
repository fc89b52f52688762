"""Input side of the `ganon-build` drop-in (ganon_b200/build.py: parse_input_table, read_sequences, --min-length, files
dropped on a parse error) differentially against the UNMODIFIED reference builder: random input tables over randomly
formatted FASTA / FASTQ / gzip files (wrapped lines, blanks, digits, CRLF, `;` headers, illegal letters ...) must give the
same per-target hash counts and the same IBF parameters.  The device calls are answered by the oracle (tests/build_util.py).
750 further seeds were run in round 1 without a mismatch (the first version of read_sequences failed 113 of 150)."""
import gzip
import os
import random
import subprocess

import pytest

from ganon_b200 import build as B
from ganon_b200 import formats
from tests import fuzz_util as F
from tests import reader_util as R
from tests.build_util import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BUILD = os.path.join(ROOT, "oracle", "_ref", "ganon-build")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_BUILD), reason="oracle/_ref not built (only in the build container)")


def _case(seed, tmp):
    rng = random.Random(seed)
    k = rng.choice((8, 12, 19))
    w = k + rng.choice((0, 4, 12))
    tsv = os.path.join(tmp, "b%d.tsv" % seed)
    desc = []
    with open(tsv, "w") as t:
        for fi in range(rng.randint(1, 5)):
            recs = [(("s%d_%d" % (fi, j)).encode(), F._seq(rng, rng.choice((5, 15, 40, 200, 1000)))) for j in range(rng.randint(1, 4))]
            fmt = rng.choice(("fasta", "fasta", "fastq"))
            style = {s: rng.random() < 0.3 for s in R.STYLES}
            if rng.random() < 0.2:
                j = rng.randrange(len(recs))
                rid, s = recs[j]
                p = rng.randrange(len(s))
                recs[j] = (rid, s[:p] + rng.choice((b"X", b"E", b"*", b"-")) + s[p + 1 :])
                style["illegal letter"] = True
            data = R.dress(rng, recs, fmt, style)
            path = os.path.join(tmp, "b%d_%d.%s" % (seed, fi, "fa" if fmt == "fasta" else "fq"))
            if rng.random() < 0.2:
                path += ".gz"
                data = gzip.compress(data)
            with open(path, "wb") as f:
                f.write(data)
            t.write("%s\t%s\n" % (path, "T%d" % rng.randrange(3)))
            desc.append((fmt, [s for s in style if style[s]]))
    min_len = rng.choice((0, 0, 30))
    out_ref, out_mine = os.path.join(tmp, "b%d.ref.ibf" % seed), os.path.join(tmp, "b%d.mine.ibf" % seed)
    pr = subprocess.run([REF_BUILD, "-i", tsv, "-o", out_ref, "-k", str(k), "-w", str(w), "-p", "0.05", "-y", str(min_len), "--quiet"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    cfg = B.GanonBuildConfig(input_file=tsv, output_file=out_mine, kmer_size=k, window_size=w, max_fp=0.05, min_length=min_len, quiet=True)
    ok = B.run_build(cfg, backend=OracleBackend())
    assert bool(ok) == (pr.returncode == 0), (seed, desc, pr.stderr[-300:])
    if not ok:
        return
    a, b = formats.read_ibf(out_mine), formats.read_ibf(out_ref)
    assert sorted(a.hashes_count) == sorted(b.hashes_count), (seed, k, w, min_len, desc)
    assert (a.ibf.bins, a.ibf.bin_size, a.ibf.hash_funs, a.max_hashes_bin, a.max_fp, a.true_max_fp) == (b.ibf.bins, b.ibf.bin_size, b.ibf.hash_funs, b.max_hashes_bin, b.max_fp, b.true_max_fp), (seed, desc)


@pytest.mark.parametrize("first", range(0, 60, 10))
def test_builder_reads_inputs_like_the_reference(first, tmp_path, capsys):
    for seed in range(first, first + 10):
        _case(seed, str(tmp_path))
    capsys.readouterr()  # "Error parsing file [...]" lines of the dropped files
