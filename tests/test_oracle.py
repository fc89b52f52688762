"""CPU tests: pin the oracle (oracle/ganon_oracle.c) against
  * seqan3's own known-answer tests for the arithmetic of the path, and
  * the committed outputs of the unmodified reference binary (tests/golden/expected, made by make_golden.py).
"""
import numpy as np
import pytest

from ganon_b200 import formats
from oracle import oracle as O
from tests import scenario_util as SU


# ---- seqan3 KATs: libs/seqan3/test/unit/search/views/minimiser_hash_test.cpp:41,62-79,103-136
def test_minimiser_hash_seqan3_kats():
    assert O.minimiser_hash(b"ACGGCGACGTTTAG", 4, 8, seed=0).tolist() == [26, 97, 27, 6, 1]
    assert O.minimiser_hash(b"ACGTCGACGTTTAG", 4, 8, seed=0).tolist() == [27, 97, 27, 6, 1]
    assert O.minimiser_hash(b"A" * 19, 4, 8, seed=0).tolist() == [0, 0, 0]
    assert O.minimiser_hash(b"A" * 19, 4, 8, seed=0x8F3F73B5CF1C9ADE).tolist() == [0x8F3F73B5CF1C9A21] * 3
    assert O.minimiser_hash(b"AC", 4, 8, seed=0).tolist() == []
    # stop_at_t: "ACGGCGACG" -> {26, 97}
    assert O.minimiser_hash(b"ACGGCGACG", 4, 8, seed=0).tolist() == [26, 97]


def test_adjust_seed():
    # src/utils/include/utils/adjust_seed.hpp:33-37
    assert O.adjust_seed(19) == 0x8F3F73B5CF1C9ADE >> 26 == 0x23CFDCED73
    assert O.adjust_seed(32) == 0x8F3F73B5CF1C9ADE


def test_dna4_rank_table():
    L = O.lib()
    table = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3, "R": 0, "W": 0, "M": 0, "D": 0, "H": 0, "V": 0, "N": 0, "Y": 1, "S": 1, "B": 1, "K": 2}
    for ch, r in table.items():
        assert L.go_dna4_rank(ord(ch)) == r and L.go_dna4_rank(ord(ch.lower())) == r
        assert L.go_dna15_valid(ord(ch)) and L.go_dna15_valid(ord(ch.lower()))
    for ch in "EFIJLOPQXZ*-. 0":
        assert not L.go_dna15_valid(ord(ch))


def test_ibf_emplace_bulk_count_consistency():
    # the property checked by tests/ganon-build/GanonBuild.test.cpp:54-92 and seqan3's IBF tests:
    # every emplaced value is found in its bin
    rng = np.random.default_rng(1)
    ibf = O.OracleIBF(bins=70, bin_size=1009, hash_funs=4)
    vals = rng.integers(0, 1 << 38, size=200, dtype=np.uint64)
    bins = rng.integers(0, 70, size=200)
    for v, b in zip(vals, bins):
        ibf.emplace(int(v), int(b))
    for v, b in zip(vals, bins):
        c = ibf.bulk_count(np.array([v], dtype=np.uint64))
        assert c[b] == 1
        assert c[70:].sum() == 0
    c = ibf.bulk_count(vals)
    for b in range(70):
        assert c[b] >= (bins == b).sum()
    # rows are < bin_size and depend on the hash function
    r = ibf.rows(12345)
    assert len(set(r)) > 1 and all(0 <= x < 1009 for x in r)


def test_threshold_helpers():
    L = O.lib()
    assert L.go_threshold_cutoff(17, 0.75) == 13
    assert L.go_threshold_cutoff(17, 0.0) == 1
    assert L.go_threshold_cutoff(0, 0.5) == 1
    assert L.go_threshold_filter(20, 10, 0.1) == 19
    assert L.go_threshold_filter(20, 10, 1.0) == 10
    assert L.go_threshold_filter(20, 20, 0.5) == 20


def test_ibf_file_roundtrip(golden_dbs, tmp_path):
    db = formats.read_ibf(golden_dbs["real4"])
    assert (db.ibf.bins, db.ibf.bin_words, db.ibf.hash_funs, db.kmer_size, db.window_size) == (64, 1, 5, 19, 31)
    assert db.ibf.hash_shift == 64 - int(db.ibf.bin_size).bit_length()
    p = str(tmp_path / "copy.ibf")
    formats.write_ibf(p, db)
    assert open(p, "rb").read() == open(golden_dbs["real4"], "rb").read()


def _run_oracle_scenario(name, dbs):
    args = SU.expand(SU.load_scenarios()[name], dbs)
    cfg = SU.parse_args(args)
    reads = []
    for f in cfg["single"]:
        reads += [(i, s, None) for i, s in O.parse_reads(f)]
    for a, b in zip(cfg["paired"][0::2], cfg["paired"][1::2]):
        reads += [(x[0], x[1], y[1]) for x, y in zip(O.parse_reads(a), O.parse_reads(b))]
    per_level = {}
    for lab, lev in cfg["levels"]:
        filters = [O.OracleFilter.from_ibf_file(formats.read_ibf(p), c) for p, c in lev["filters"]]
        res = O.classify_level(filters, reads, lev["rel_filter"], lev["fpr_query"])
        per_level[lab] = res
        reads = [r for r, x in zip(reads, res) if not x["matches"]]
    unc = sorted(r[0].decode() for r in reads)
    return cfg, per_level, unc


@pytest.mark.parametrize("name", sorted(SU.load_scenarios()))
def test_oracle_matches_reference_outputs(name, golden_dbs):
    cfg, per_level, unc = _run_oracle_scenario(name, golden_dbs)
    labels = [lab for lab, _ in cfg["levels"]]
    if len(labels) > 1 and "-s" not in cfg["flags"]:
        for lab in labels:
            assert O.all_lines(per_level[lab]) == SU.expected_lines(name, lab + ".all"), lab
    else:
        mine = sorted(sum((O.all_lines(per_level[lab]) for lab in labels), []))
        assert mine == SU.expected_lines(name, "all")
    assert unc == SU.expected_lines(name, "unc")


# ---- HIBF: the oracle's traversal (go_hibf_bulk_count) against `ganon-classify --hibf` outputs (make_golden_hibf.py)
def _oracle_hibf_filter(path, rel_cutoff):
    h = formats.read_hibf(path)
    ibfs = [O.OracleIBF(i.bins, i.bin_size, i.hash_funs, i.data) for i in h.ibfs]
    oh = O.OracleHIBF(ibfs, h.next_ibf_id, h.bin_to_user, len(h.bin_path))
    tmap = {}
    for u, paths in enumerate(h.bin_path):
        for p in paths:
            tmap.setdefault(formats.hibf_target_name(p), []).append(u)
    targets = list(tmap)
    return O.OracleFilter(oh, targets, [tmap[t] for t in targets], [h.fpr] * len(targets), rel_cutoff, h.kmer_size, h.window_size)


@pytest.mark.parametrize("name", sorted(SU.load_hibf_scenarios()))
def test_oracle_hibf_matches_reference_outputs(name, golden_dbs):
    args = SU.expand(SU.load_hibf_scenarios()[name], golden_dbs)
    cfg = SU.parse_args(args)
    reads = []
    for f in cfg["single"]:
        reads += [(i, s, None) for i, s in O.parse_reads(f)]
    for a, b in zip(cfg["paired"][0::2], cfg["paired"][1::2]):
        reads += [(x[0], x[1], y[1]) for x, y in zip(O.parse_reads(a), O.parse_reads(b))]
    (lab, lev), = cfg["levels"]
    filt = _oracle_hibf_filter(lev["filters"][0][0], lev["filters"][0][1])
    res = O.classify_level([filt], reads, lev["rel_filter"], lev["fpr_query"])
    assert O.all_lines(res) == SU.expected_lines(name, "all")
    assert sorted(r["id"].decode() for r in res if not r["matches"]) == SU.expected_lines(name, "unc")
