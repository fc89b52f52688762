"""CPU-side checks of the product boundary: the shared library builds/loads, exports exactly what include/ganon_b200.h
declares, fails loudly without a GPU (no fallback), and the host logic (CLI grammar, config validation) matches the
reference's Config.hpp / CommandLineParser.cpp behaviour."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from ganon_b200 import _lib, cli
from ganon_b200.classify import GanonClassifyConfig


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ganon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 25
    assert sorted(_lib.SYMBOLS) == declared  # the ctypes table and the header agree
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (gnb_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert L.gnb_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    n = C.c_int(-1)
    assert L.gnb_device_count(C.byref(n)) == -4 and n.value == 0  # GNB_ERR_CUDA
    h = C.c_void_p()
    rc = L.gnb_db_create(64, 1000, 3, 19, 31, 0, C.byref(h))
    assert rc == -4 and not h.value
    assert b"cuda" in L.gnb_last_error().lower()
    out = (C.c_uint64 * 8)()
    nout = C.c_uint64()
    assert L.gnb_minimisers(0, 19, 31, b"ACGT" * 20, 80, out, 8, C.byref(nout)) == -4


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_cli_grammar():
    cfg = cli.parse(["-r", "a.fq,b.fq", "--paired-reads=c.1.fq,c.2.fq", "-i", "x.ibf", "-i", "y.ibf", "-c", "0.1,0.3", "-d0.5", "-o", "out", "-au", "--output-stats", "-t", "8", "--hibf"])
    assert cfg.single_reads == ["a.fq", "b.fq"] and cfg.paired_reads == ["c.1.fq", "c.2.fq"]
    assert cfg.ibf == ["x.ibf", "y.ibf"] and cfg.rel_cutoff == [0.1, 0.3] and cfg.rel_filter == [0.5]
    assert cfg.output_all and cfg.output_unclassified and cfg.output_stats and not cfg.output_lca
    assert cfg.threads == 8 and cfg.hibf and cfg.output_prefix == "out"
    # defaults of Config.hpp:30-49
    d = GanonClassifyConfig()
    assert (d.rel_cutoff, d.rel_filter, d.fpr_query, d.hierarchy_labels, d.tax_root_node, d.n_reads, d.n_batches) == ([0.2], [0.0], [1.0], ["H1"], "1", 400, 1000)
    with pytest.raises(cli.CliError):
        cli.parse(["--no-such-flag"])
    # exit codes of main.cpp:9-16
    assert cli.main([]) == 1 and cli.main(["-h"]) == 0 and cli.main(["-v"]) == 0


def test_config_validation_messages(tmp_path, capsys):
    f = tmp_path / "r.fq"
    f.write_text("@r\nACGT\n+\nIIII\n")
    db = tmp_path / "d.ibf"
    db.write_bytes(b"x")

    def err(**kw):
        c = GanonClassifyConfig(**kw)
        assert not c.validate()
        return capsys.readouterr().err

    assert "--output-prefix is mandatory" in err()
    assert "At least one of --[single|paired|batch]-reads is mandatory" in err(output_prefix="o")
    assert "--ibf is mandatory" in err(output_prefix="o", single_reads=[str(f)])
    assert "even number" in err(output_prefix="o", paired_reads=[str(f)], ibf=[str(db)])
    assert "file not found" in err(output_prefix="o", single_reads=["/nonexistent"], ibf=[str(db)])
    assert "--rel-cutoff values" in err(output_prefix="o", single_reads=[str(f)], ibf=[str(db)], rel_cutoff=[1.5])
    assert "one-per-hierarchy --rel-filter" in err(output_prefix="o", single_reads=[str(f)], ibf=[str(db), str(db)], hierarchy_labels=["a", "b"], rel_filter=[0.1, 0.2, 0.3])
    assert "--hierarchy does not match" in err(output_prefix="o", single_reads=[str(f)], ibf=[str(db), str(db)], hierarchy_labels=["a", "b", "c"], rel_filter=[0.1, 0.2, 0.3])
    # broadcast rules (Config.hpp:175-245)
    c = GanonClassifyConfig(output_prefix="o", single_reads=[str(f)], ibf=[str(db)] * 3, hierarchy_labels=["b", "a", "b"], rel_cutoff=[0.3])
    assert c.validate()
    assert c.rel_cutoff == [0.3] * 3 and c.rel_filter == [0.0, 0.0] and c.fpr_query == [1.0, 1.0] and c.skip_lca


def test_graft_entry_build_imports():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G

    assert callable(G.build) and callable(G.smoke)


def test_classify_wrapper_argument_mapping(tmp_path, monkeypatch):
    """ganon_b200.classify.classify(cfg) maps ganon's Config like src/ganon/classify.py:29-64; with reassign_in_memory the EM
    step of classify.py:76-88 is folded into the run."""
    import types

    from ganon_b200 import classify as K

    (tmp_path / "db.ibf").write_bytes(b"x")
    (tmp_path / "db.tax").write_bytes(b"x")
    seen = {}
    monkeypatch.setattr(K, "run", lambda c: seen.setdefault("cfg", c) is not None)
    base = dict(db_prefix=[str(tmp_path / "db")], single_reads=["r.fq"], paired_reads=[], batch_reads=[], output_prefix="out", hierarchy_labels=None, rel_cutoff=[0.5],
                rel_filter=[0.2], fpr_query=[1e-3], output_one=True, output_all=False, output_unclassified=True, output_stats=False, output_single=False, threads=3,
                verbose=False, quiet=True, hibf=False)
    assert K.classify(types.SimpleNamespace(multiple_matches="em", **base))
    c = seen.pop("cfg")
    assert c.ibf == [str(tmp_path / "db.ibf")] and c.tax == [str(tmp_path / "db.tax")] and c.skip_lca and c.output_all and not c.output_lca and not c.reassign_em
    assert (c.rel_cutoff, c.rel_filter, c.fpr_query, c.threads) == ([0.5], [0.2], [1e-3], 3)
    assert K.classify(types.SimpleNamespace(multiple_matches="em", reassign_in_memory=True, max_iter=0, threshold=0.01, **base))
    c = seen.pop("cfg")
    assert c.reassign_em and not c.output_all and c.em_write_one and c.em_max_iter == 0 and c.em_threshold == [0.01]
    assert K.classify(types.SimpleNamespace(multiple_matches="lca", **base))
    c = seen.pop("cfg")
    assert not c.skip_lca and c.output_lca and not c.output_all and not c.reassign_em


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under ganon_b200/, bin/ or include/ may import, include, link or run it."""
    import glob

    hits = []
    for pat in ("ganon_b200/*.py", "ganon_b200/csrc/*.cpp", "ganon_b200/csrc/*.cu", "ganon_b200/csrc/*.h", "ganon_b200/csrc/*.cuh", "ganon_b200/csrc/Makefile", "bin/*", "include/*.h"):
        for p in glob.glob(os.path.join(ROOT, pat)):
            with open(p, errors="replace") as f:
                for i, line in enumerate(f, 1):
                    if re.search(r"\boracle\b", line) and not re.search(r"oracle's|the oracle|against the oracle|vs the oracle|oracle/", line):
                        hits.append((p, i, line.strip()))
                    if re.search(r"(import|from|include|dlopen|CDLL).*oracle", line):
                        hits.append((p, i, line.strip()))
    assert not hits, hits
