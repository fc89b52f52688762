"""The reference's own tests of the builder (tests/ganon-build/GanonBuild.test.cpp:131-577), restated for the
`ganon-build` drop-in (ganon_b200/build.py): every section's configuration on the same literal sequences
(tests/golden/build_kats.json, extracted by tests/golden/make_golden_refkats.py), with the reference's two validators

  * validate_filter   (:20-52)  bins == bin map == config, hash functions as configured, the achieved false-positive rate within
                                the requested one to two decimals;
  * validate_elements (:54-98)  every minimiser of every input sequence is found in the bins of its target;

on the CPU with the oracle as the device, next to the unmodified `ganon-build` where oracle/_ref exists (same IBF parameters),
and on the GPU with K2 + insertion on the device (`-m gpu`).
"""
import json
import math
import os
import subprocess

import numpy as np
import pytest

from ganon_b200 import build as B
from ganon_b200 import formats
from oracle import oracle as O
from tests import scenario_util as SU
from tests.build_util import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BUILD = os.path.join(ROOT, "oracle", "_ref", "ganon-build")
REF_DATA = "/root/reference/tests/ganon-build/data"
SEQS = json.load(open(os.path.join(SU.GOLDEN, "build_kats.json")))


def default_config(prefix, **kw):
    """config_build::defaultConfig (:100-112)."""
    base = dict(input_file=prefix + "_input.tsv", output_file=prefix + ".ibf", quiet=True, kmer_size=19, window_size=32, hash_functions=4, max_fp=0.05)
    base.update(kw)
    return B.GanonBuildConfig(**base)


def write_seqtarget(prefix, seqs, targets=None):
    """aux::SeqTarget (tests/aux/Aux.hpp:148-176): one FASTA file per sequence, `<prefix>.SEQ<i>.fasta`, header SEQ<i>; the
    input table has one column (target = file name) or two.  Returns [(target, sequence)]."""
    out = []
    with open(prefix + "_input.tsv", "w") as t:
        for i, s in enumerate(seqs):
            p = os.path.abspath("%s.SEQ%d.fasta" % (prefix, i))
            with open(p, "w") as f:
                f.write(">SEQ%d\n%s\n" % (i, s))
            if targets is None:
                t.write(p + "\n")
                out.append((os.path.basename(p), s))
            else:
                t.write("%s\t%s\n" % (p, targets[i]))
                out.append((targets[i], s))
    return out


def validate_filter(cfg):
    assert os.path.getsize(cfg.output_file) > 0
    db = formats.read_ibf(cfg.output_file)
    assert db.ibf.bins == len(db.bin_map)
    assert sorted(b for b, _t in db.bin_map) == list(range(db.ibf.bins))
    if cfg.hash_functions > 0:
        assert db.ibf.hash_funs == cfg.hash_functions
    if not cfg.filter_size:
        assert math.floor(db.true_max_fp * 100.0) / 100.0 <= math.floor(cfg.max_fp * 100.0) / 100.0
        assert math.floor(db.true_avg_fp * 100.0) / 100.0 <= math.floor(cfg.max_fp * 100.0) / 100.0
    return db


def validate_elements(cfg, seqtarget):
    db = formats.read_ibf(cfg.output_file)
    ibf = O.OracleIBF(db.ibf.bins, db.ibf.bin_size, db.ibf.hash_funs, db.ibf.data)
    bins = {}
    for b, t in db.bin_map:
        bins.setdefault(t, []).append(b)
    for target, seq in seqtarget:
        h = O.minimiser_hash(seq.encode(), cfg.kmer_size, cfg.window_size)
        counts = ibf.bulk_count(h)
        assert int(sum(int(counts[b]) for b in bins[target])) == h.size, target


def same_parameters_as_reference(cfg, tmp):
    """The unmodified ganon-build on the same input picks the same IBF parameters (bins, bin size, hash functions, maximum
    hashes per bin, false-positive figures); bin numbering and bit content are not comparable."""
    if not os.path.exists(REF_BUILD):
        return
    out = os.path.join(tmp, os.path.basename(cfg.output_file) + ".ref")
    argv = [REF_BUILD, "-i", cfg.input_file, "-o", out, "-k", str(cfg.kmer_size), "-w", str(cfg.window_size), "-s", str(cfg.hash_functions), "-j", cfg.mode,
            "-y", str(cfg.min_length), "--quiet"] + (["-f", repr(cfg.filter_size)] if cfg.filter_size else ["-p", repr(cfg.max_fp)])
    pr = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert pr.returncode == 0, pr.stderr
    a, b = formats.read_ibf(cfg.output_file), formats.read_ibf(out)
    assert (a.ibf.bins, a.ibf.bin_size, a.ibf.hash_funs, a.max_hashes_bin) == (b.ibf.bins, b.ibf.bin_size, b.ibf.hash_funs, b.max_hashes_bin)
    assert (a.max_fp, a.true_max_fp) == (b.max_fp, b.true_max_fp)
    assert a.true_avg_fp == pytest.approx(b.true_avg_fp, rel=1e-12)  # summed in hash-map order by the reference
    assert sorted(a.hashes_count) == sorted(b.hashes_count)
    assert os.path.getsize(cfg.output_file) == os.path.getsize(out)


# (name, GanonBuild.test.cpp lines, config overrides, sequence set, custom targets, expected run() result)
SECTIONS = [
    ("input_file_one_col", "171-183", {}, "seqs", None, True),
    ("input_file_two_cols", "185-199", {}, "seqs", ["T1", "T9", "T1", "T8", "T1", "T1", "T1", "T1", "T4", "T1"], True),
    ("max_fp_0.01", "202-215", dict(max_fp=0.01), "seqs", None, True),
    ("max_fp_0.5", "217-228", dict(max_fp=0.5), "seqs", None, True),
    ("filter_size_0.1", "233-247", dict(filter_size=0.1), "seqs", None, True),
    ("filter_size_1", "249-260", dict(filter_size=1.0), "seqs", None, True),
    ("hash_functions_0", "342-355", dict(hash_functions=0), "seqs", None, True),
    ("hash_functions_2", "357-370", dict(hash_functions=2), "seqs", None, True),
    ("hash_functions_6", "373-384", dict(hash_functions=6), "seqs", None, False),
    ("w32_k19", "390-404", dict(window_size=32, kmer_size=19), "seqs", None, True),
    ("w23_k21", "406-420", dict(window_size=23, kmer_size=21), "seqs", None, True),
    ("w27_k27", "422-436", dict(window_size=27, kmer_size=27), "seqs", None, True),
    ("w42_k35", "438-450", dict(window_size=42, kmer_size=35), "seqs", None, False),
    ("w12_k32", "452-464", dict(window_size=12, kmer_size=32), "seqs", None, False),
    ("tmp_output_folder_empty", "470-483", dict(tmp_output_folder=""), "seqs", None, True),
    ("tmp_output_folder_existing", "485-501", dict(tmp_output_folder="{prefix}"), "seqs", None, True),
    ("tmp_output_folder_non_existing", "503-515", dict(tmp_output_folder="{prefix}_missing"), "seqs", None, False),
    ("min_length_0", "535-548", dict(min_length=0), "seqs2", None, True),
    ("min_length_50", "550-573", dict(min_length=50), "seqs2", None, True),
]


def _run_section(sec, tmp, backend_factory, compare_reference):
    name, _lines, over, seqset, targets, ok = sec
    prefix = os.path.join(tmp, name)
    over = {k: (v.replace("{prefix}", prefix) if isinstance(v, str) else v) for k, v in over.items()}
    if name == "tmp_output_folder_existing":
        os.makedirs(prefix, exist_ok=True)
    cfg = default_config(prefix, **over)
    seqtarget = write_seqtarget(prefix, SEQS[seqset], targets)
    backend = backend_factory()
    try:
        assert B.run_build(cfg, backend=backend) == ok, name
    finally:
        backend.close()
    if not ok:
        return None
    validate_filter(cfg)
    if name == "min_length_50":  # :562-572: only the sequences of at least 50 bp are in the filter
        seqtarget = [(t, s) for t, s in seqtarget if len(s) >= 50]
        assert [s for _t, s in seqtarget] == SEQS["seqs3"]
    validate_elements(cfg, seqtarget)
    if compare_reference:
        same_parameters_as_reference(cfg, tmp)
    return cfg


@pytest.mark.parametrize("sec", SECTIONS, ids=[s[0] for s in SECTIONS])
def test_build_sections_oracle_backend(sec, tmp_path):
    _run_section(sec, str(tmp_path), OracleBackend, True)


def test_build_file_sizes_follow_max_fp_and_filter_size(tmp_path):
    """:230 a smaller --max-fp gives a larger file; :262 a larger --filter-size gives a larger file."""
    by_name = {s[0]: s for s in SECTIONS}
    size = lambda n: os.path.getsize(_run_section(by_name[n], str(tmp_path), OracleBackend, False).output_file)
    assert size("max_fp_0.01") > size("max_fp_0.5")
    assert size("filter_size_0.1") < size("filter_size_1")


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="needs the reference's test genomes (/root/reference, build container only)")
def test_build_modes_on_the_reference_test_genomes(tmp_path, monkeypatch):
    """--mode (:265-337) on tests/ganon-build/data/mode_input.tsv (25 small genomes, paths relative to that directory)."""
    monkeypatch.chdir(REF_DATA)
    run = {}
    for name, over in [("avg_fp", dict(max_fp=0.001, mode="avg")), ("smallest_fp", dict(mode="smallest")),  # :283-286: smallest runs at the default 0.05
                       ("avg_fs", dict(filter_size=1.0, mode="avg")), ("smallest_fs", dict(filter_size=1.0, mode="smallest")), ("fastest_fs", dict(filter_size=1.0, mode="fastest"))]:
        cfg = default_config(str(tmp_path / name), input_file="mode_input.tsv", **over)
        assert B.run_build(cfg, backend=OracleBackend())
        run[name] = (cfg, validate_filter(cfg))
        same_parameters_as_reference(cfg, str(tmp_path))
    assert os.path.getsize(run["smallest_fp"][0].output_file) < os.path.getsize(run["avg_fp"][0].output_file)  # :291
    assert run["smallest_fs"][1].max_fp < run["avg_fs"][1].max_fp  # :331
    assert run["fastest_fs"][1].ibf.bins < run["avg_fs"][1].ibf.bins  # :334


@pytest.mark.gpu
@pytest.mark.parametrize("sec", SECTIONS, ids=[s[0] for s in SECTIONS])
def test_build_sections_on_gpu(sec, tmp_path):
    """The same sections with K2 and the insertion on the device; the file must equal the oracle-backend one byte for byte
    (the layout is deterministic in the drop-in)."""
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    cfg = _run_section(sec, str(tmp_path / "gpu"), lambda: B.GpuBackend(0), False)
    if cfg is None:
        return
    ref = _run_section(sec, str(tmp_path / "cpu"), OracleBackend, False)
    a, b = formats.read_ibf(cfg.output_file), formats.read_ibf(ref.output_file)
    assert (a.ibf.bins, a.ibf.bin_size, a.ibf.hash_funs, a.max_hashes_bin) == (b.ibf.bins, b.ibf.bin_size, b.ibf.hash_funs, b.max_hashes_bin)
    assert np.array_equal(np.asarray(a.ibf.data), np.asarray(b.ibf.data))


# ---- not in the reference's test file: sequences shorter than the window -------------------------------------------------
SHORT = ["ACGTTGCAAGCTTGCAATGCATGCA", "ACGTTGCAAGCTTGCAATGCATGCAACGTTGCAAGCTTGCAATGCATGCATTTT", "ACGTACGTACGTACGTAC", "TTGCAAGCTTGCAATGCAT"]  # 25, 54, 18 (< k), 19 bp


def _short_case(tmp, backend_factory):
    prefix = os.path.join(tmp, "short")
    cfg = default_config(prefix, kmer_size=19, window_size=31, hash_functions=0)
    write_seqtarget(prefix, SHORT, ["A", "B", "C", "D"])
    backend = backend_factory()
    try:
        assert B.run_build(cfg, backend=backend)
    finally:
        backend.close()
    return cfg


def test_sequences_shorter_than_the_window_count_like_the_reference(tmp_path):
    """seqan3's minimiser view shrinks the window to a sequence shorter than it (minimiser.hpp:298-299): the reference builder
    counts one minimiser for 19..30 bp at k=19, w=31 and none below k -- unlike ganon-classify, which skips such reads."""
    cfg = _short_case(str(tmp_path), OracleBackend)
    assert dict(formats.read_ibf(cfg.output_file).hashes_count) == {"A": 1, "B": 4, "C": 0, "D": 1}
    same_parameters_as_reference(cfg, str(tmp_path))


@pytest.mark.gpu
def test_sequences_shorter_than_the_window_on_gpu(tmp_path):
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    a = formats.read_ibf(_short_case(str(tmp_path / "gpu"), lambda: B.GpuBackend(0)).output_file)
    b = formats.read_ibf(_short_case(str(tmp_path / "cpu"), OracleBackend).output_file)
    assert a.hashes_count == b.hashes_count and np.array_equal(np.asarray(a.ibf.data), np.asarray(b.ibf.data))
