"""Randomised differential cases (tests/fuzz_util.py).
CPU: the oracle against the unmodified reference binary (skipped where oracle/_ref is not built).
GPU: the `ganon-classify` drop-in (K1-K4 / host stage) against the oracle on the same kind of cases -- random k / w
(every K2 code path), bin counts, hash functions, multi-bin targets, several filters and hierarchy levels."""
import os

import pytest

from tests import fuzz_util as F


@pytest.mark.skipif(not os.path.exists(F.REF_BIN), reason="oracle/_ref/ganon-classify is not built")
@pytest.mark.parametrize("seed", range(100, 112))
def test_oracle_matches_reference_on_random_cases(seed, tmp_path):
    ok, desc = F.run_case(seed, str(tmp_path))
    assert ok, desc


@pytest.mark.skipif(not os.path.exists(F.REF_BIN), reason="oracle/_ref/ganon-classify is not built")
@pytest.mark.parametrize("seed", range(500, 505))
def test_oracle_hibf_matches_reference_on_random_cases(seed, tmp_path):
    ok, desc = F.run_hibf_case(seed, str(tmp_path))
    assert ok, desc


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(200, 216))
def test_dropin_matches_oracle_on_random_cases(seed, tmp_path):
    from ganon_b200 import cli

    args, desc, oracle_lines = F.make_case(seed, str(tmp_path))
    out = str(tmp_path / "mine")
    assert cli.main(args + ["-o", out, "-t", "2", "--quiet"]) == 0, desc
    want_all, want_unc = oracle_lines()
    assert F.read_sorted(out + ".all") == want_all, desc
    assert F.read_sorted(out + ".unc") == want_unc, desc
