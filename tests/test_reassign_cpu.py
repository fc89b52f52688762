"""The EM reassignment restatement (oracle/reassign_oracle.py) against the fixtures the unmodified reference module
wrote (tests/golden/make_golden_em.py -> tests/golden/expected_em/)."""
import glob
import os

import pytest

from oracle import reassign_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETTINGS = {"default": dict(threshold=0, max_iter=10), "one_iter": dict(threshold=0, max_iter=1), "thr": dict(threshold=0.05, max_iter=0)}
SCENARIOS = sorted({os.path.basename(p).split("__")[0] for p in glob.glob(os.path.join(GOLDEN, "expected_em", "*.rep"))})


def _read(p):
    with open(p) as f:
        return f.read()


@pytest.mark.parametrize("setting", sorted(SETTINGS))
@pytest.mark.parametrize("scenario", SCENARIOS)
def test_restatement_matches_reference_reassign(scenario, setting):
    rep = _read(os.path.join(GOLDEN, "expected", scenario + ".rep"))
    have = [os.path.basename(p)[len(scenario) + 1 :] for p in glob.glob(os.path.join(GOLDEN, "expected", scenario + ".*all"))]
    labels = RO.all_files_of(rep, have)
    assert labels
    texts = {h: _read(os.path.join(GOLDEN, "expected", scenario + ("." + h if h else "") + ".all")) for h in labels}
    ones, new_rep = RO.reassign_texts(rep, texts, **SETTINGS[setting])
    pre = os.path.join(GOLDEN, "expected_em", "%s__%s" % (scenario, setting))
    assert new_rep == _read(pre + ".rep")
    for h, one in ones.items():
        name = pre + (".%s.one" % h if len(ones) > 1 else ".one")
        assert one == _read(name), name


def test_scenarios_cover_multi_matching_reads():
    assert len(SCENARIOS) >= 8
    # the fixtures exercise the EM itself: some read changes its target between the one-iteration and the converged run
    changed = 0
    for sc in SCENARIOS:
        a = glob.glob(os.path.join(GOLDEN, "expected_em", sc + "__one_iter*.one"))
        for p in a:
            q = p.replace("__one_iter", "__default")
            changed += _read(p) != _read(q)
    assert changed >= 1
