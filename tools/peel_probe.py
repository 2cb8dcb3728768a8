"""Peel-only timing probe on a few graph shapes (debug aid; KOMBGPU_DEBUG=1 prints the
kernel's phase breakdown)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("KOMBGPU_DEBUG", "1")
import numpy as np
import komb_b200
from komb_b200 import synth

ctx = komb_b200.Context(0)
which = sys.argv[1:] or ["cfg2", "rmat22", "ramp"]
for w in which:
    if w == "cfg2":
        m1, m2 = synth.metagenome_hits(1_000_000, 5_000_000, seed=11)
        g = ctx.build_graph(np.concatenate([m1.read_key, m2.read_key]), np.concatenate([m1.unitig, m2.unitig]), 1_000_000)
    elif w.startswith("rmat"):
        scale = int(w[4:])
        n = int(0.75 * (1 << scale)); m = 10 * n
        u, v = synth.rmat_edges(scale, m, n_vertices=n, seed=42)
        g = ctx.graph_from_edges(u, v, n)
    elif w == "ramp":
        u, v = synth.ramp_edges(1500, 20)
        g = ctx.graph_from_edges(u, v, 1500 * 20)
    for rep in range(3):
        # re-run the peel: rebuild is not needed, coreness is cached per graph, so rebuild graph cheaply
        st = g.stats()
        print(w, "n", st["n_vertices"], "E", st["n_edges"], "maxdeg", st["max_degree"], flush=True)
        g.coreness(copy=False)
        st = g.stats()
        print(w, "peel ms", round(st["ms_peel"], 3), "kernel ms", round(st["ms_peel_kernel"], 3), "levels", st["peel_levels"],
              "rounds", st["peel_rounds"], "kmax", st["max_coreness"],
              "GB/s", round((24 * st["n_edges"] + 16 * st["n_vertices"]) / st["ms_peel_kernel"] / 1e6, 1), flush=True)
        break
    g.close()
