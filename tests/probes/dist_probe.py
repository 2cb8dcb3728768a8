"""Scale probe of the partitioned path: world 2 as two processes over gloo on whatever GPUs exist.
usage: python tests/probes/dist_probe.py n_unitigs n_read_pairs [world]"""
import os, pickle, socket, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch.multiprocessing as mp


def worker(rank, world, port, n, reads, out_dir):
    import torch, torch.distributed as dist
    import komb_b200
    from komb_b200 import synth
    from komb_b200.distributed import Comm, CudaEngine, analyse_partitioned
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = komb_b200.Context(dev)
    per = reads // world
    m1, m2 = synth.metagenome_hits(n, per, seed=11 + rank, read_offset=rank * per)
    rk = torch.from_numpy(np.concatenate([m1.read_key, m2.read_key]).view(np.int32)).cuda()
    ut = torch.from_numpy(np.concatenate([m1.unitig, m2.unitig]).view(np.int32)).cuda()
    torch.cuda.synchronize()
    print(rank, "inputs ready", rk.numel(), flush=True)
    t0 = time.perf_counter()
    res = analyse_partitioned(CudaEngine(ctx), Comm(), n, read_key=rk, unitig=ut)
    torch.cuda.synchronize()
    print(rank, "done", round(time.perf_counter() - t0, 3), "s", res.n_edges, res.max_coreness, res.stats, flush=True)
    pickle.dump({"core": res.coreness.cpu().numpy(), "deg": res.degree.cpu().numpy()}, open(Path(out_dir) / f"r{rank}.pkl", "wb"))
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    n, reads = int(sys.argv[1]), int(sys.argv[2])
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(world, port, n, reads, d), nprocs=world, join=True)
        parts = [pickle.load(open(Path(d) / f"r{r}.pkl", "rb")) for r in range(world)]
    from komb_b200 import synth
    from oracle import oracle
    per = reads // world
    rks, uts = [], []
    for r in range(world):
        m1, m2 = synth.metagenome_hits(n, per, seed=11 + r, read_offset=r * per)
        rks += [m1.read_key, m2.read_key]; uts += [m1.unitig, m2.unitig]
    edges, _, _ = oracle.build_edges(np.concatenate(rks), np.concatenate(uts))
    deg, core = oracle.coreness(n, edges)
    print("deg equal", np.array_equal(np.concatenate([p["deg"] for p in parts]), deg),
          "core equal", np.array_equal(np.concatenate([p["core"] for p in parts]), core))
