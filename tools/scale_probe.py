"""Full-size runs of BASELINE.json configs 3 and 5 on one GPU, with size-independent checks.

    python tools/scale_probe.py cfg3 [--scale 26 --n 50000000 --draws 540000000] [--oracle]
    python tools/scale_probe.py cfg5 [--levels 5000 --per-level 40] [--oracle]

Inputs are generated on the device (torch); the path runs through the C ABI's device-pointer entry
points.  Checks: (1) k-core certificate on the GPU: every vertex has >= core(v) neighbours of coreness
>= core(v) (so the assignment is feasible: core <= true coreness) and <= core(v) neighbours of
coreness > core(v) (necessary for maximality); (2) --oracle: bit-exact comparison with the CPU BZ
oracle on the same edges and CORE-A against the oracle's (the checker itself is tests/probes/scale_oracle_check.py)."""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import komb_b200


def rmat_device(scale, draws, n, seed, chunk=1 << 26):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    us, vs = [], []
    a, b, c = 0.57, 0.19, 0.19
    for s0 in range(0, draws, chunk):
        m = min(chunk, draws - s0)
        u = torch.zeros(m, dtype=torch.int64, device="cuda"); v = torch.zeros_like(u)
        for _ in range(scale):
            r = torch.rand(m, device="cuda", generator=g)
            ub = (r >= a + b).to(torch.int64)
            vb = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
            u = (u << 1) | ub; v = (v << 1) | vb
        def scr(x):
            x = (x ^ (x >> 16)) & 0xFFFFFFFF; x = (x * 0x7FEB352D) & 0xFFFFFFFF
            x = (x ^ (x >> 15)) & 0xFFFFFFFF; x = (x * 0x846CA68B) & 0xFFFFFFFF
            return (x ^ (x >> 16)) & 0xFFFFFFFF
        us.append((scr(u) % n).to(torch.int32)); vs.append((scr(v) % n).to(torch.int32))
    return torch.cat(us), torch.cat(vs)


def ramp_device(levels, per, n_bg, bg_scale, bg_edges, seed):
    L = levels * per
    j = torch.arange(L, dtype=torch.int64, device="cuda"); c = 1 + j // per
    us, vs = [], []
    for d in range(1, levels + 1):
        sel = j[(c >= d) & (j + d < L)]
        us.append(sel); vs.append(sel + d)
    t = torch.arange(L - (levels + 1), L, dtype=torch.int64, device="cuda")
    iu = torch.triu_indices(levels + 1, levels + 1, 1, device="cuda")
    us.append(t[iu[0]]); vs.append(t[iu[1]])
    u = torch.cat(us) + n_bg; v = torch.cat(vs) + n_bg
    if bg_edges:
        bu, bv = rmat_device(bg_scale, bg_edges, n_bg, seed)
        u = torch.cat([u, bu.to(torch.int64)]); v = torch.cat([v, bv.to(torch.int64)])
    n = n_bg + L
    # scramble ids by a fixed bijection so a range partition would be balanced
    a = 2654435761
    while np.gcd(a, n) != 1: a += 2
    f = lambda x: ((x * a + 12345) % n).to(torch.int32)
    return f(u), f(v), n


def certificate(core, col, deg, n, chunk=1 << 27):
    ge = torch.zeros(n, dtype=torch.int32, device="cuda"); gt = torch.zeros(n, dtype=torch.int32, device="cuda")
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); torch.cumsum(deg.to(torch.int64), 0, out=row_ptr[1:])
    total = int(row_ptr[-1])
    for s0 in range(0, total, chunk):
        e = torch.arange(s0, min(total, s0 + chunk), device="cuda")
        src = torch.searchsorted(row_ptr, e, right=True) - 1
        cu = core[col[s0:s0 + e.numel()].to(torch.int64)]; cs = core[src]
        ge.index_add_(0, src, (cu >= cs).to(torch.int32)); gt.index_add_(0, src, (cu > cs).to(torch.int32))
    return bool((ge >= core).all()), bool((gt <= core).all()), bool((core <= deg).all())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfg"); ap.add_argument("--scale", type=int, default=26); ap.add_argument("--n", type=int, default=50_000_000)
    ap.add_argument("--draws", type=int, default=540_000_000); ap.add_argument("--levels", type=int, default=5000)
    ap.add_argument("--per-level", type=int, default=40); ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    ctx = komb_b200.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    if a.cfg == "cfg3":
        u, v = rmat_device(a.scale, a.draws, a.n, 42); n = a.n
    else:
        u, v, n = ramp_device(a.levels, a.per_level, 9_800_000, 24, 40_000_000, 7)
    torch.cuda.synchronize()
    print(f"generated {u.numel()} pairs over {n} vertices in {time.perf_counter() - t0:.1f} s", flush=True)
    out = {"config": a.cfg, "n": n, "input_pairs": u.numel()}
    for rep in range(a.reps):
        g = ctx.graph_from_edges(u, v, n)
        g.analyse(komb_b200.KEY_EXACT64)
        st = g.stats()
        E = st["n_edges"]
        b_peel = 24 * E + 16 * n
        print(f"rep {rep}: E={E} maxdeg={st['max_degree']} kmax={st['max_coreness']} levels={st['peel_levels']} "
              f"build {st['ms_build']:.1f} ms | peel {st['ms_peel']:.1f} ms (kernel {st['ms_peel_kernel']:.1f}; "
              f"{E / st['ms_peel_kernel'] / 1e6:.2f} G edges/s, {b_peel / st['ms_peel_kernel'] / 1e6:.0f} GB/s = "
              f"{b_peel / st['ms_peel_kernel'] / 1e6 / 6539.5:.3f} of HBM) | corea {st['ms_corea']:.1f} ms", flush=True)
        out.update({k: st[k] for k in ("n_edges", "max_degree", "max_coreness", "peel_levels", "peel_rounds", "ms_build", "ms_peel", "ms_peel_kernel", "ms_corea")})
        if rep < a.reps - 1:
            g.close()
    arr = g.device_arrays()
    from komb_b200.distributed import _DevArray
    core = torch.as_tensor(_DevArray(arr["coreness"], n, "<i4", g), device="cuda")
    deg = torch.as_tensor(_DevArray(arr["degree"], n, "<i4", g), device="cuda")
    col = torch.as_tensor(_DevArray(arr["col"], 2 * E, "<i4", g), device="cuda")
    t0 = time.perf_counter()
    c1, c2, c3 = certificate(core, col, deg, n)
    print(f"certificate: feasible={c1} no-vertex-misses-a-higher-core={c2} core<=deg={c3} ({time.perf_counter() - t0:.1f} s)", flush=True)
    out.update({"cert_feasible": c1, "cert_maximal_necessary": c2, "cert_core_le_deg": c3})
    if a.oracle:   # the CPU checker lives with the tests: tests/probes/scale_oracle_check.py
        sys.path.insert(0, str(ROOT / "tests" / "probes"))
        from scale_oracle_check import check_against_oracle
        out.update(check_against_oracle(g, n, deg, core))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
