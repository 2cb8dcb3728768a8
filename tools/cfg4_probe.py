"""BASELINE.json config 4 on N GPUs of one box: 2 B alignment hits -> edge build + dedup to ~1 B edges ->
k-core -> CORE-A, end to end on the partitioned path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/cfg4_probe.py [--unitigs 100000000 --read-pairs 500000000 --p 0.2 --reps 2]

Every rank generates the hits of its own read range on its device (seed 1234 + rank): read pair r has a centre
c ~ scrambled power law(0.5) over the unitigs and k1 + k2 hits (k uniform over {1,1,2,2,2,3,3} per mate) at
(c + Geometric(p) - 1) mod n.  p = 0.2 was calibrated on the CPU oracle at 1/100 scale (n = 1 M, 5 M read pairs:
E/n = 9.65, P/E = 2.30), so n = 1e8 gives E ~ 0.97e9.  Checks at this size are the size-independent ones: sum of
degrees = 2E, coreness <= degree, every rank reports the same global figures."""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist


def gen_hits(n, r_lo, r_cnt, p, seed, chunk=1 << 25):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    a = 2654435761
    while np.gcd(a, n) != 1:
        a += 2
    ks = torch.tensor([1, 1, 2, 2, 2, 3, 3], device="cuda")
    centre = torch.empty(r_cnt, dtype=torch.int64, device="cuda")
    for s0 in range(0, r_cnt, chunk):
        m = min(chunk, r_cnt - s0)
        x = torch.rand(m, device="cuda", generator=g, dtype=torch.float64)
        c = torch.clamp((n * x * x).to(torch.int64), max=n - 1)          # inverse CDF of p(u) ~ (u+1)^-0.5
        centre[s0:s0 + m] = (c * a + 12345) % n                           # fixed bijection: hubs spread over the id range
    mates = []
    for _ in range(2):
        rk_parts, ut_parts = [], []
        for s0 in range(0, r_cnt, chunk):
            m = min(chunk, r_cnt - s0)
            k = ks[torch.randint(0, 7, (m,), device="cuda", generator=g)]
            reads = torch.repeat_interleave(torch.arange(s0, s0 + m, device="cuda"), k)
            off = torch.empty(reads.numel(), device="cuda").geometric_(p, generator=g).to(torch.int64) - 1
            ut_parts.append(((centre[reads] + off) % n).to(torch.int32))
            rk_parts.append((reads + r_lo).to(torch.int32))               # bit pattern of the u32 read key
        mates.append((torch.cat(rk_parts), torch.cat(ut_parts)))
    return torch.cat([mates[0][0], mates[1][0]]), torch.cat([mates[0][1], mates[1][1]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--unitigs", type=int, default=100_000_000)
    ap.add_argument("--read-pairs", type=int, default=500_000_000)
    ap.add_argument("--p", type=float, default=0.2)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg3"],
                    help="cfg3: R-MAT scale 26, 540 M draws over 50 M vertices, every rank draws its share of the edge list")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import komb_b200
    from komb_b200.distributed import Comm, CudaEngine, analyse_partitioned
    ctx = komb_b200.Context(local)
    comm, eng = Comm("nccl"), CudaEngine(ctx)
    per = a.read_pairs // world
    t0 = time.perf_counter()
    if a.config == "cfg3":
        sys.path.insert(0, str(Path(__file__).resolve().parent))
        from scale_probe import rmat_device
        a.unitigs = 50_000_000
        rk, ut = rmat_device(26, 540_000_000 // world, a.unitigs, 42 + rank)    # (u, v) draws of this rank
    else:
        rk, ut = gen_hits(a.unitigs, rank * per, per, a.p, 1234 + rank)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"generated {rk.numel()} {'pairs' if a.config == 'cfg3' else 'hits'} per rank in {time.perf_counter() - t0:.1f} s", flush=True)
    out = None
    for rep in range(a.reps):
        stage = {}
        marks = [None]

        def tick(name):
            torch.cuda.synchronize()
            now = time.perf_counter()
            stage[name] = now - marks[0]
            marks[0] = now
        dist.barrier(); torch.cuda.synchronize()
        marks[0] = time.perf_counter()
        t_start = marks[0]
        if a.config == "cfg3":
            res = analyse_partitioned(eng, comm, a.unitigs, pairs=(rk, ut), timer=tick, key_mode=komb_b200.KEY_EXACT64)
        else:
            res = analyse_partitioned(eng, comm, a.unitigs, read_key=rk, unitig=ut, timer=tick)
        torch.cuda.synchronize(); dist.barrier()
        wall = time.perf_counter() - t_start
        sums = comm.all_gather_ints([int(res.degree.to(torch.int64).sum()), int((res.coreness > res.degree).sum()), rk.numel(),
                                     int(wall * 1e6)])
        H, E, n = int(sums[:, 2].sum()), res.n_edges, a.unitigs
        ms_peel = res.stats.get("ms_peel_kernel", 0.0)
        out = {"config": a.config, "n_gpus": world, "rep": rep, "n_unitigs": n, "n_hits": H, "n_pairs": res.stats["sum_pairs"], "n_edges": E,
               "max_coreness": res.max_coreness, "peel_levels": res.stats["levels"], "peel_mode": res.stats["peel_mode"],
               "wall_ms_max_over_ranks": float(sums[:, 3].max()) / 1e3, "stage_ms_rank0": {k: round(v * 1e3, 2) for k, v in stage.items()},
               "ms_peel_kernel": ms_peel, "ms_corea": res.stats.get("ms_corea"),
               "hits_per_s": H / (float(sums[:, 3].max()) / 1e6),
               "peel_edges_per_s": E / (ms_peel * 1e-3) if ms_peel else None,
               "peel_frac_of_aggregate_hbm": ((24 * E + 16 * n) / (ms_peel * 1e-3) / 1e9 / (6539.5 * world)) if ms_peel else None,
               "check_degree_sum_is_2E": int(sums[:, 0].sum()) == 2 * E, "check_core_le_deg": int(sums[:, 1].sum()) == 0}
        if rank == 0:
            print(json.dumps(out), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
