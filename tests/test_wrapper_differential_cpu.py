"""`ganon_b200.classify.classify(cfg)` against the reference's own Python wrapper (`src/ganon/classify.py:7-64`), imported
from /root/reference in the build container (its `multitax` dependency and package metadata are stubbed, nothing of it is
copied): for random `ganon classify` parameter sets, the command line the reference assembles for `ganon-classify`, parsed
by the drop-in's flag grammar, must be the configuration the drop-in's wrapper hands to its own run()."""
import dataclasses
import os
import random
import shlex
import sys
import types

import pytest

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_SRC, "ganon")), reason="needs /root/reference (build container only)")


@pytest.fixture(scope="module")
def reference_modules():
    import importlib.metadata as md

    saved_version, saved_path, saved_mods = md.version, list(sys.path), {k: sys.modules.get(k) for k in ("multitax", "ganon")}
    md.version = lambda n: "0.0.0" if n == "ganon" else saved_version(n)
    stub = types.ModuleType("multitax")
    for cls in ("CustomTx", "NcbiTx", "GtdbTx", "DummyTx"):
        setattr(stub, cls, type(cls, (), {"_supported_versions": []}))
    sys.modules["multitax"] = stub
    sys.path.insert(0, REF_SRC)
    try:
        import ganon.classify as RC
        from ganon.config import Config

        yield RC, Config
    finally:
        md.version = saved_version
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "ganon" or k.startswith("ganon.")] + ["multitax"]:
            sys.modules.pop(k, None)
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v


def _params(rng, d):
    n_db = rng.randint(1, 3)
    dbs = []
    for i in range(n_db):
        pre = os.path.join(d, "db%d_%d" % (rng.randrange(1 << 30), i))
        open(pre + (".hibf" if rng.random() < 0.2 else ".ibf"), "wb").write(b"x")
        if rng.random() < 0.7:
            open(pre + ".tax", "wb").write(b"x")
        dbs.append(pre)
    p = dict(db_prefix=dbs, output_prefix=os.path.join(d, "out"), quiet=True)
    if rng.random() < 0.8:
        p["single_reads"] = [os.path.join(d, "r%d.fq" % i) for i in range(rng.randint(1, 2))]
    else:
        p["paired_reads"] = [os.path.join(d, "p.1.fq"), os.path.join(d, "p.2.fq")]
    if rng.random() < 0.5:
        p["rel_cutoff"] = [rng.choice(("0", "0.25", "0.75", "1")) for _ in range(rng.choice((1, n_db)))]  # lists are handed to argparse as strings
    if rng.random() < 0.5:
        p["rel_filter"] = [rng.choice(("0", "0.1", "1"))]
    if rng.random() < 0.5:
        p["fpr_query"] = [rng.choice(("1e-5", "0.001", "1"))]
    if rng.random() < 0.4:
        p["hierarchy_labels"] = [rng.choice(("1_a", "2_b")) for _ in range(n_db)]
    p["multiple_matches"] = rng.choice(("em", "lca", "skip"))
    for flag in ("output_one", "output_all", "output_unclassified", "output_single", "verbose"):
        if rng.random() < 0.3:
            p[flag] = True
    if rng.random() < 0.5:
        p["threads"] = rng.choice((1, 4, 16))
    if rng.random() < 0.2:
        p["n_reads"] = rng.choice((1, 400))
    if rng.random() < 0.2:
        p["n_batches"] = rng.choice((5, 1000))
    return p


def test_wrapper_builds_the_reference_command_line(reference_modules, tmp_path, monkeypatch):
    RC, Config = reference_modules
    from ganon_b200 import classify as K
    from ganon_b200 import cli

    seen = {}
    monkeypatch.setattr(RC, "run", lambda cmd, **kw: seen.__setitem__("cmd", cmd))
    monkeypatch.setattr(RC, "reassign", lambda c: True)
    monkeypatch.setattr(RC, "report", lambda c: True)
    monkeypatch.setattr(K, "run", lambda c: seen.__setitem__("mine", c) is None)
    for f in ("r0.fq", "r1.fq", "p.1.fq", "p.2.fq"):
        (tmp_path / f).write_text("@r\nACGT\n+\nIIII\n")
    compared = 0
    for seed in range(300):
        rng = random.Random(seed)
        cfg = Config("classify", **_params(rng, str(tmp_path)))
        cfg.set_paths = lambda: True
        cfg.path_exec = {"classify": "ganon-classify"}
        RC.classify(cfg)
        want = cli.parse(shlex.split(seen.pop("cmd"))[1:])
        assert K.classify(cfg)
        got = seen.pop("mine")
        for fld in dataclasses.fields(want):
            if fld.name in ("reassign_em", "em_write_one", "em_max_iter", "em_threshold", "device"):  # extensions of the drop-in
                continue
            assert getattr(got, fld.name) == getattr(want, fld.name), (seed, fld.name, getattr(got, fld.name), getattr(want, fld.name))
        compared += 1
    assert compared == 300
