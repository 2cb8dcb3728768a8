"""A/B timing of the peel kernel variants and knobs on a few graph shapes (measurement aid).

    python tools/peel_ab.py [cfg2] [rmat22] [ramp] [rmat24d]

For each graph: peel with the CTA-wide process phase (reference result + time), then the warp-autonomous
one under several knob settings; every result is compared with the first.  Uses Graph.coreness(again=True) so that one
graph is peeled many times."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import komb_b200
from komb_b200 import synth

VARIANTS = [
    ("cta", {"KOMBGPU_PEEL_MODE": "cta"}),
    ("warp", {"KOMBGPU_PEEL_MODE": "warp"}),
    ("warp thin128", {"KOMBGPU_PEEL_MODE": "warp", "KOMBGPU_PEEL_THIN": "128"}),
    ("warp thin256", {"KOMBGPU_PEEL_MODE": "warp", "KOMBGPU_PEEL_THIN": "256"}),
    ("warp thin1024", {"KOMBGPU_PEEL_MODE": "warp", "KOMBGPU_PEEL_THIN": "1024"}),
    ("warp thin4096", {"KOMBGPU_PEEL_MODE": "warp", "KOMBGPU_PEEL_THIN": "4096"}),
    ("warp thin1024 wsplit256", {"KOMBGPU_PEEL_MODE": "warp", "KOMBGPU_PEEL_THIN": "1024", "KOMBGPU_PEEL_WSPLIT": "256"}),
]
KNOBS = ["KOMBGPU_PEEL_MODE", "KOMBGPU_PEEL_KEEP", "KOMBGPU_PEEL_PARK", "KOMBGPU_PEEL_WSPLIT", "KOMBGPU_PEEL_UNROLL", "KOMBGPU_PEEL_THIN"]


def make_graph(ctx, w):
    if w == "cfg2":
        m1, m2 = synth.metagenome_hits(1_000_000, 5_000_000, seed=11)
        return ctx.build_graph(np.concatenate([m1.read_key, m2.read_key]), np.concatenate([m1.unitig, m2.unitig]), 1_000_000)
    if w == "path":      # one chain: the peel's latency per dependent step, nothing else (two waves meet in the middle)
        n = 200_000
        u = np.arange(n - 1, dtype=np.uint32)
        return ctx.graph_from_edges(u, u + 1, n)
    if w == "ramp":
        u, v = synth.ramp_edges(1500, 20)
        return ctx.graph_from_edges(u, v, 1500 * 20)
    if w.startswith("rmat") and w.endswith("d"):   # device-generated, cfg3 proportions at a smaller scale
        import torch
        sys.path.insert(0, str(Path(__file__).resolve().parent))
        from scale_probe import rmat_device
        scale = int(w[4:-1])
        n = int(0.745 * (1 << scale)); m = int(10.8 * n)
        u, v = rmat_device(scale, m, n, 42)
        return ctx.graph_from_edges(u, v, n)
    if w.startswith("rmat"):
        scale = int(w[4:])
        n = int(0.75 * (1 << scale)); m = 10 * n
        u, v = synth.rmat_edges(scale, m, n_vertices=n, seed=42)
        return ctx.graph_from_edges(u, v, n)
    raise SystemExit(f"unknown graph {w}")


def main():
    ctx = komb_b200.Context(0)
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["cfg2", "rmat22", "ramp"]
    reps = 3
    out = []
    for w in which:
        g = make_graph(ctx, w)
        ref = None
        for name, env in VARIANTS:
            for kname in KNOBS:
                os.environ.pop(kname, None)
            os.environ.update(env)
            best = 1e30
            ok = True
            try:
                for _ in range(reps):
                    core = g.coreness(again=True)
                    st = g.stats()
                    best = min(best, st["ms_peel_kernel"])
                    if ref is None:
                        ref = core
                    elif not np.array_equal(ref, core):
                        ok = False
            except Exception as exc:  # a broken variant must not hide the others
                print(json.dumps({"graph": w, "variant": name, "error": str(exc)}), flush=True)
                continue
            row = {"graph": w, "variant": name, "ms": round(best, 3), "equal": ok, "n": st["n_vertices"], "E": st["n_edges"],
                   "levels": st["peel_levels"], "rounds": st["peel_rounds"], "kmax": st["max_coreness"],
                   "G_edges_s": round(st["n_edges"] / best / 1e6, 2)}
            print(json.dumps(row), flush=True)
            out.append(row)
        g.close()
    ctx.close()


if __name__ == "__main__":
    main()
