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
    n = 1_000_000 * world
    m1, m2 = synth.metagenome_hits(n, 5_000_000, seed=11 + rank, read_offset=rank * 5_000_000, scramble=True)
    rk = torch.from_numpy(np.concatenate([m1.read_key, m2.read_key]).view(np.int32)).cuda()
    ut = torch.from_numpy(np.concatenate([m1.unitig, m2.unitig]).view(np.int32)).cuda()
    torch.cuda.synchronize()
    for rep in range(a.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = DistGraph.from_hits(comm, rk, ut, n)
        g.analyse()
        torch.cuda.synchronize(); dist.barrier()
        wall = time.perf_counter() - t0
        st = g.stats()
        g.close()
        if rank == 0:
            print(json.dumps({"rep": rep, "wall_ms": wall * 1e3, **{k: st[k] for k in ("ms_build", "ms_build_route", "ms_build_sort", "ms_build_csr",
                  "ms_peel", "ms_corea", "peel_subrounds", "peel_solo_subrounds", "peel_levels", "n_messages_sent", "n_edges_global")}}), flush=True)
    dist.barrier()
    comm.close(); ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
