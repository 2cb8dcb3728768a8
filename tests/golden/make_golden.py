"""Regenerate tests/golden/* from the REFERENCE ITSELF (authoring container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) and `make -C oracle ref`, which compiles the
reference's own src/{gfa,graph,komb2}.cpp unmodified against
oracle/igraph_shim into oracle/_ref/komb2_ref, and CoreA::getAnomalyScore
straight from src/CoreA.h into oracle/_ref/corea_ref.  The reference has no
tests or golden vectors of its own, so these fixtures
are its outputs on deterministic generated inputs:

  komb2/<case>/r{1,2}.sam.gz + expected.json   whole-path: komb2_ref -t T
        canonical (Name-keyed) edge set, Name -> (coreness, degree),
        Name -> CoreA score text (%f)
  corea/<case>.npz     (coreness, degree) -> score from CoreA.h, incl. an
        int32-overflow case (quirk Q5)
"""
import gzip
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from komb_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = Path(__file__).resolve().parent

KOMB2_CASES = {
    # name: (tiny_sam_pair kwargs, komb2 -t)
    "quickstart_s1": (dict(seed=1, n_unitigs=30, n_reads=60), 1),
    "mid_s2": (dict(seed=2, n_unitigs=200, n_reads=500), 1),
    "mid_s3_t4": (dict(seed=3, n_unitigs=200, n_reads=500), 4),       # quirk Q1: -t 4 drops lines
    "nosuffix_s6": (dict(seed=6, n_unitigs=60, n_reads=150, qname_suffix=False), 1),
    "noheader_s7": (dict(seed=7, n_unitigs=80, n_reads=200, with_header=False, unmapped_every=0), 1),
    "wide_s4": (dict(seed=4, n_unitigs=2000, n_reads=3000, max_hits=4), 1),
}


def make_komb2():
    for name, (kw, t) in KOMB2_CASES.items():
        s1, s2, _, _ = synth.tiny_sam_pair(**kw)
        with tempfile.TemporaryDirectory() as d:
            ref, stdout = oracle.run_komb2(oracle.REF_KOMB2, s1, s2, d, threads=t)
        cdir = HERE / "komb2" / name
        cdir.mkdir(parents=True, exist_ok=True)
        for fn, data in (("r1.sam.gz", s1), ("r2.sam.gz", s2)):
            with gzip.GzipFile(cdir / fn, "wb", mtime=0) as f:
                f.write(data)
        info = [l.strip() for l in stdout.splitlines()
                if l.strip().startswith(("Number of", "Dense Ratio", "Max CoreA"))]
        exp = {
            "threads": t,
            "edges": sorted([list(e) for e in ref["edges"]]),
            "kcore": {k: list(v) for k, v in sorted(ref["kcore"].items())},
            "score_text": dict(sorted(ref["score_text"].items())),
            "stdout_info": info,
        }
        (cdir / "expected.json").write_text(json.dumps(exp, separators=(",", ":")))
        print(name, "edges", len(exp["edges"]), "vertices", len(exp["kcore"]), info)


def make_corea():
    rng = np.random.Generator(np.random.PCG64(5))
    cases = {}
    # power-law-ish degrees, coreness <= degree, few distinct keys
    n = 4000
    deg = np.floor(1.0 / rng.random(n) ** 0.8).astype(np.int32)
    core = np.minimum(deg, rng.integers(0, 12, n)).astype(np.int32)
    cases["powerlaw_n4000"] = (core, deg)
    # all tied / all distinct / zeros
    cases["all_equal"] = (np.full(50, 3, np.int32), np.full(50, 7, np.int32))
    cases["distinct"] = (np.arange(300, dtype=np.int32) // 3, np.arange(300, dtype=np.int32)[::-1].copy())
    cases["with_isolated"] = (np.array([0, 0, 1, 1, 2, 2, 2, 0], np.int32), np.array([0, 0, 1, 3, 2, 2, 5, 0], np.int32))
    # quirk Q5: coreness * n overflows int32 (n = 40000, coreness up to 120000)
    n = 40000
    core = rng.integers(0, 120000, n).astype(np.int32)
    deg = (core + rng.integers(0, 5000, n)).astype(np.int32)
    cases["overflow_q5_n40000"] = (core, deg)
    (HERE / "corea").mkdir(exist_ok=True)
    for name, (core, deg) in cases.items():
        with tempfile.TemporaryDirectory() as d:
            score = oracle.corea_reference(core, deg, d)
        np.savez_compressed(HERE / "corea" / f"{name}.npz", core=core, deg=deg, score=score)
        print(name, core.shape[0], float(score.max()))


if __name__ == "__main__":
    assert oracle.REF_KOMB2.exists() and oracle.REF_COREA.exists(), "run `make -C oracle ref` first"
    make_komb2()
    make_corea()
