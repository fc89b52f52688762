"""Build-side host arithmetic (ganon_b200/build.py) against the IBFConfig the unmodified reference `ganon-build` chose
(tests/golden/build_cases.json, written by tests/golden/make_golden_build.py): 48 random genome sets x parameter
combinations (--max-fp / --filter-size / --hash-functions / --mode)."""
import json
import os

import numpy as np
import pytest

from ganon_b200 import build as B
from oracle import oracle as O

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "build_cases.json")))


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_chosen_parameters_match_reference_build(ci):
    c = CASES[ci]
    p = c["params"]
    counts = {t: n for t, n in c["hashes_count"]}
    got = B.choose_ibf_params(counts, max_fp=p["max_fp"], filter_size=p["filter_size"], hash_functions=p["hash_functions"], mode=p["mode"])
    assert (got.n_bins, got.bin_size_bits, got.hash_functions, got.max_hashes_bin) == (c["n_bins"], c["bin_size_bits"], c["hash_functions"], c["max_hashes_bin"]), p
    assert got.max_fp == c["max_fp"]
    assert got.true_max_fp == c["true_max_fp"]
    assert got.true_avg_fp == pytest.approx(c["true_avg_fp"], rel=1e-12)  # summed in hash-map order by the reference
    # bins per target as in the reference's bin map
    layout = B.bin_layout(counts, got)
    assert len(layout) == got.n_bins
    per_target = {}
    for t, first, last in layout:
        per_target[t] = per_target.get(t, 0) + 1
        assert 0 <= first <= last < counts[t] and last - first + 1 <= got.max_hashes_bin
    ref = {}
    for _b, t in c["bin_map"]:
        ref[t] = ref.get(t, 0) + 1
    assert per_target == ref


def test_hash_counts_are_distinct_minimisers():
    """hashes_count of the reference = distinct minimisers of the target's sequences (count_hashes, GanonBuild.cpp:184-249)."""
    n = 0
    for c in CASES:
        if not c["genomes"]:
            continue
        counts = dict((t, k) for t, k in c["hashes_count"])
        for t, seq in c["genomes"].items():
            hs = O.minimiser_hash(seq.encode(), c["params"]["k"], c["params"]["w"])
            assert np.unique(hs).size == counts[t], (t, c["params"])
            n += 1
    assert n >= 20
