"""Time the peer-memory multi-GPU path on cfg2 x N (or cfg3 / cfg4 shares) with the peel's own phase profile
(KOMBGPU_DEBUG=1 prints where the leader thread's time goes).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/peer_probe.py [--reps 3] [--debug]
"""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg4_eighth"])
    a = ap.parse_args()
    if a.debug:
        os.environ["KOMBGPU_DEBUG"] = "1"
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import komb_b200
    from komb_b200 import synth
    from komb_b200.peer import Comm, DistGraph
    ctx = komb_b200.Context(local)
    comm = Comm.from_torch(ctx, heap_bytes=1 << 30)
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench
    kind = "hits"
    if a.config == "cfg2":
        n = 1_000_000 * world
        m1, m2 = synth.metagenome_hits(n, 5_000_000, seed=11 + rank, read_offset=rank * 5_000_000, scramble=True)
        rk = torch.from_numpy(np.concatenate([m1.read_key, m2.read_key]).view(np.int32)).cuda()
        ut = torch.from_numpy(np.concatenate([m1.unitig, m2.unitig]).view(np.int32)).cuda()
    elif a.config == "cfg3":
        n, kind = 50_000_000, "pairs"
        rk, ut = bench.rmat_device(26, 540_000_000 // world, n, 42 + 1000 * rank)
    else:
        scale = 8 if a.config == "cfg4_eighth" else 1
        n, per = 100_000_000 // scale, 500_000_000 // scale // world
        rk, ut = bench.cfg4_hits_device(n, rank * per, per, 0.2, 1234 + rank)
    torch.cuda.synchronize()
    for rep in range(a.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = DistGraph.from_hits(comm, rk, ut, n) if kind == "hits" else DistGraph.from_pairs(comm, rk, ut, n)
        g.analyse(komb_b200.KEY_EXACT64 if kind == "pairs" else komb_b200.KEY_REF32)
        torch.cuda.synchronize(); dist.barrier()
        wall = time.perf_counter() - t0
        st = g.stats()
        g.close()
        keys = ("ms_build", "ms_build_route", "ms_build_sort", "ms_build_csr", "ms_peel", "ms_corea", "n_pairs_local", "n_pairs_received",
                "n_fwd_local", "n_directed_local")
        rows = [None] * world
        dist.all_gather_object(rows, {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in keys})
        if rank == 0:
            print(json.dumps({"rep": rep, "wall_ms": wall * 1e3, **{k: st[k] for k in ("peel_subrounds", "peel_solo_subrounds", "peel_levels",
                  "n_edges_global")}, "per_rank": rows}), flush=True)
    dist.barrier()
    comm.close(); ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
