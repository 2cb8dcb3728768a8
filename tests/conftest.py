"""pytest configuration: `-m gpu` tests need a B200, everything else runs on CPU."""
import gzip
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_komb2_case(name):
    cdir = GOLDEN / "komb2" / name
    sam1 = gzip.open(cdir / "r1.sam.gz").read()
    sam2 = gzip.open(cdir / "r2.sam.gz").read()
    exp = json.loads((cdir / "expected.json").read_text())
    exp["edges"] = {tuple(e) for e in exp["edges"]}
    exp["kcore"] = {k: tuple(v) for k, v in exp["kcore"].items()}
    return sam1, sam2, exp


def komb2_case_names():
    return sorted(p.name for p in (GOLDEN / "komb2").iterdir() if p.is_dir())


def corea_case_names():
    return sorted(p.stem for p in (GOLDEN / "corea").glob("*.npz"))


def load_corea_case(name):
    z = np.load(GOLDEN / "corea" / f"{name}.npz")
    return z["core"], z["deg"], z["score"]


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle
