"""Checker leg of tools/scale_probe.py --oracle: the full-size GPU result against the CPU oracle (test infrastructure;
the oracle may only be used from tests/, smoke() and bench.py's CPU leg)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
import komb_b200
from oracle import oracle


def check_against_oracle(g, n, deg, core):
    """g: komb_b200.Graph after analyse(KEY_EXACT64); deg / core: torch device tensors of its results."""
    t0 = time.perf_counter()
    eu, ev = g.edges()
    edges = oracle.pack_edges(eu, ev)
    odeg, ocore = oracle.coreness(n, edges)
    t1 = time.perf_counter()
    ok_d, ok_c = bool(np.array_equal(odeg, deg.cpu().numpy())), bool(np.array_equal(ocore, core.cpu().numpy()))
    score = g.corea(komb_b200.KEY_EXACT64)
    osc = oracle.corea(ocore, odeg, oracle.KEY_EXACT64)
    ok_s = bool(np.allclose(score, osc, rtol=1e-6, atol=1e-12))
    print(f"oracle (CPU BZ, {t1 - t0:.1f} s incl. D2H): degree equal {ok_d}, coreness equal {ok_c}, CORE-A within 1e-6 {ok_s}", flush=True)
    return {"oracle_degree_equal": ok_d, "oracle_coreness_equal": ok_c, "oracle_corea_close": ok_s,
            "oracle_peel_edges_per_s": edges.shape[0] / (t1 - t0)}
