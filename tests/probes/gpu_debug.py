"""Stage-by-stage bring-up check on a GPU box; logs progressively so that a hang
is localised.  usage: python tests/probes/gpu_debug.py [stage...]"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
LOG = open(ROOT / "gpurun_out" / "debug.log", "a", buffering=1)


def log(*a):
    msg = f"[{time.strftime('%H:%M:%S')}] " + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


log("start", sys.argv)
import numpy as np  # noqa: E402
log("numpy ok")
import komb_b200  # noqa: E402
from komb_b200 import synth  # noqa: E402
log("komb_b200 ok")
from oracle import oracle  # noqa: E402
oracle.lib()
log("oracle ok")
ctx = komb_b200.Context(0)
log("ctx ok")

stages = sys.argv[1:] or ["edges_tiny", "edges_mid", "hits_tiny", "hits_mid"]


def run_graph(g, n, exp_edges, tag):
    log(tag, "built", g.counts(), g.stats())
    u, v = g.edges()
    log(tag, "edges equal:", np.array_equal(oracle.pack_edges(u, v), exp_edges))
    exp_deg, exp_core = oracle.coreness(n, exp_edges)
    deg = g.degree()
    log(tag, "deg equal:", np.array_equal(deg, exp_deg))
    row_ptr, col = g.csr()
    log(tag, "row_ptr ok:", np.array_equal(np.diff(row_ptr.astype(np.int64)), exp_deg), "col max", int(col.max()) if col.size else -1)
    eu, ev = oracle.unpack_edges(exp_edges)
    src = np.repeat(np.arange(n, dtype=np.uint64), exp_deg)
    keys = (src << np.uint64(32)) | col.astype(np.uint64)
    sym = np.sort(np.concatenate([oracle.pack_edges(eu, ev), oracle.pack_edges(ev, eu)]))
    log(tag, "csr equal:", np.array_equal(keys, sym))
    log(tag, "peel...")
    core = g.coreness()
    log(tag, "core equal:", np.array_equal(core, exp_core), g.stats())
    for mode in (0, 1):
        s = g.corea(mode)
        e = oracle.corea(exp_core, exp_deg, mode)
        log(tag, "corea mode", mode, "maxabs", float(np.max(np.abs(s - e))) if s.size else 0.0, "summary", g.summary())


for st in stages:
    if st.startswith("edges"):
        scale, n, m = {"edges_tiny": (8, 200, 1500), "edges_mid": (16, 50000, 600000), "edges_big": (22, 3000000, 40000000)}[st]
        u, v = synth.rmat_edges(scale, m, n_vertices=n, seed=1)
        exp = oracle.simplify(u, v)
        log(st, "input ready", m, "E", exp.shape[0])
        with ctx.graph_from_edges(u, v, n) as g:
            run_graph(g, n, exp, st)
    elif st.startswith("hits"):
        n, r = {"hits_tiny": (300, 500), "hits_mid": (20000, 60000), "hits_big": (1000000, 5000000)}[st]
        m1, m2 = synth.metagenome_hits(n, r, seed=5)
        rk = np.concatenate([m1.read_key, m2.read_key])
        ut = np.concatenate([m1.unitig, m2.unitig])
        exp, P, S = oracle.build_edges(rk, ut)
        log(st, "input ready H", rk.shape[0], "P", P, "S", S, "E", exp.shape[0])
        with ctx.build_graph(rk, ut, n) as g:
            run_graph(g, n, exp, st)
log("done")
