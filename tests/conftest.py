import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_dbs(tmp_path_factory):
    """Decompress the golden .ibf.gz fixtures once per session; returns {name: path}."""
    import gzip
    import shutil

    d = tmp_path_factory.mktemp("golden_dbs")
    out = {}
    for n in ("real4", "real4b", "synth"):
        p = str(d / (n + ".ibf"))
        with gzip.open(os.path.join(GOLDEN, n + ".ibf.gz"), "rb") as fi, open(p, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        out[n] = p
    p = str(d / "synth.hibf")
    with gzip.open(os.path.join(GOLDEN, "synth.hibf.gz"), "rb") as fi, open(p, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    out["synth_hibf"] = p
    return out
