"""EM reassignment: the restatement (oracle/reassign_oracle.py, the checker of the device EM in tests/test_gpu_parity.py) against
the UNMODIFIED reference module `src/ganon/reassign.py`, loaded from /root/reference the way tests/golden/make_golden_em.py
does, on random `.all` / `.rep` inputs: many ties in the probabilities, targets without unique reads, reads repeated in the file,
one or several hierarchy levels with one `.all` each or a single one (--output-single), thresholds / iteration limits."""
import importlib.util
import os
import random
import sys
import types

import pytest

from oracle import reassign_oracle as RO

REF = "/root/reference/src/ganon"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "reassign.py")), reason="needs /root/reference (build container only)")


@pytest.fixture(scope="module")
def reference_reassign():
    saved = {k: sys.modules.get(k) for k in ("ganon", "ganon.util", "ganon.reassign")}
    pkg = types.ModuleType("ganon")
    pkg.__path__ = [REF]
    sys.modules["ganon"] = pkg
    try:
        for name in ("util", "reassign"):
            spec = importlib.util.spec_from_file_location("ganon." + name, os.path.join(REF, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules["ganon." + name] = mod
            spec.loader.exec_module(mod)
        yield sys.modules["ganon.reassign"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _make(rng):
    """-> (rep text, {label or "": all text}) the way ganon-classify writes them."""
    labels = rng.choice((["H1"], ["1_a", "2_b"], ["x", "y", "z"]))
    single = len(labels) > 1 and rng.random() < 0.4  # --output-single: one .all for all levels
    n_targets = rng.choice((1, 2, 5, 12))
    rep_lines, per_label = [], {}
    rid = 0
    for lab in labels:
        targets = ["%s.T%d" % (lab, t) if rng.random() < 0.7 else "T%d" % t for t in range(n_targets)]
        lines, matches_of, unique_of, lca = [], {}, {}, 0
        for _ in range(rng.choice((1, 8, 60))):
            name = "r%d" % rid
            rid += 1
            picks = rng.sample(targets, min(len(targets), rng.choice((1, 1, 2, 3, 4))))
            for t in picks:
                lines.append("%s\t%s\t%d" % (name, t, rng.choice((1, 5, 5, 17, 30))))
                matches_of[t] = matches_of.get(t, 0) + 1
            if len(picks) == 1:
                unique_of[picks[0]] = unique_of.get(picks[0], 0) + 1
            else:
                lca += 1
        if rng.random() < 0.5:
            rng.shuffle(lines)  # the reference's output order is arbitrary
        per_label[lab] = "".join(l + "\n" for l in lines)
        for t in targets:
            if t in matches_of:
                rep_lines.append("%s\t%s\t%d\t%d\t0" % (lab, t, matches_of[t], unique_of.get(t, 0)) + ("\tspecies\tname of %s" % t if rng.random() < 0.5 else ""))
        if lca:
            rep_lines.append("%s\t1\t0\t0\t%d" % (lab, lca))  # multi-matching reads counted on the root node
    rep = "".join(l + "\n" for l in rep_lines) + "#total_classified\t%d\n#total_unclassified\t%d\n" % (rid, rng.randrange(50))
    texts = {"": "".join(per_label[l] for l in labels)} if single or len(labels) == 1 else per_label
    return rep, texts


@pytest.mark.parametrize("first", range(0, 200, 25))
def test_restatement_matches_reference_on_random_inputs(reference_reassign, first, tmp_path):
    for seed in range(first, first + 25):
        rng = random.Random(seed)
        rep, texts = _make(rng)
        d = tmp_path / ("s%d" % seed)
        d.mkdir()
        (d / "in.rep").write_text(rep)
        for lab, t in texts.items():
            (d / ("in." + lab + ".all" if lab else "in.all")).write_text(t)
        setting = dict(threshold=rng.choice((0, 0, 0.01, 0.2)), max_iter=rng.choice((0, 1, 3, 10)))
        cfg = types.SimpleNamespace(input_prefix=[str(d / "in")], output_prefix=str(d / "out"), skip_rep=False, skip_one=False, remove_all=False, quiet=True, verbose=False, **setting)
        assert reference_reassign.reassign(cfg), seed
        have = [(lab + ".all" if lab else "all") for lab in texts]
        labels = RO.all_files_of(rep, have)
        assert sorted(labels) == sorted(texts), (seed, labels)
        ones, new_rep = RO.reassign_texts(rep, {h: texts[h] for h in labels}, **setting)
        assert new_rep == (d / "out.rep").read_text(), (seed, setting)
        for h, one in ones.items():
            name = "out.%s.one" % h if len(ones) > 1 else "out.one"
            assert one == (d / name).read_text(), (seed, setting, h)
