#!/usr/bin/env python
"""Golden fixtures for the EM reassignment, made by RUNNING THE UNMODIFIED REFERENCE
(`/root/reference/src/ganon/reassign.py`, the step `ganon classify --multiple-matches em` runs after the binary).

The reference package cannot be imported as a whole here (`ganon/__init__.py` asks importlib.metadata for an installed
distribution), so `ganon.util` and `ganon.reassign` are loaded from their files under a stub package.  Inputs are the
`.all` / `.rep` files the reference *binary* wrote for the golden scenarios (tests/golden/expected/), copied to a scratch
directory; outputs (committed): tests/golden/expected_em/<scenario>__<setting>.{one,rep} (+ `.<hierarchy>.one`).
"""
import glob
import importlib.util
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/ganon"
SETTINGS = {"default": dict(threshold=0, max_iter=10), "one_iter": dict(threshold=0, max_iter=1), "thr": dict(threshold=0.05, max_iter=0)}
SCENARIOS = ["pe_real4_all", "se_synth_all", "pe_synth_all", "pe_two_filters", "pe_three_filters", "hier_two_levels", "hier_two_levels_single", "hibf_all", "se_synth"]


def load_reference():
    pkg = types.ModuleType("ganon")
    pkg.__path__ = [REF]
    sys.modules["ganon"] = pkg
    for name in ("util", "reassign"):
        spec = importlib.util.spec_from_file_location("ganon." + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ganon." + name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["ganon.reassign"]


def main():
    R = load_reference()
    out_dir = os.path.join(HERE, "expected_em")
    os.makedirs(out_dir, exist_ok=True)
    for sc in SCENARIOS:
        for tag, st in SETTINGS.items():
            with tempfile.TemporaryDirectory() as tmp:
                for f in glob.glob(os.path.join(HERE, "expected", sc + ".*")):
                    shutil.copy(f, tmp)
                cfg = types.SimpleNamespace(input_prefix=[os.path.join(tmp, sc)], output_prefix=os.path.join(tmp, "out"), skip_rep=False, skip_one=False, remove_all=False,
                                            quiet=True, verbose=False, **st)
                assert R.reassign(cfg), (sc, tag)
                for f in glob.glob(os.path.join(tmp, "out*")):
                    shutil.copy(f, os.path.join(out_dir, "%s__%s%s" % (sc, tag, os.path.basename(f)[3:])))
    print("wrote", len(os.listdir(out_dir)), "files to", out_dir)


if __name__ == "__main__":
    main()
