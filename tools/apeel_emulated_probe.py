"""The asynchronous partitioned peel with two EMULATED ranks on one GPU (both ranks inside one grid, "peer" memory is local):
the arrangement a single-GPU ncu capture can see.  apeel_emulated_probe.py [n_unitigs=1000000] [read_pairs=5000000]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["KOMBGPU_DIST_PEEL"] = "async"
import numpy as np, torch
from komb_b200 import synth
from komb_b200.peer import DistGraph, run_local
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
m1, m2 = synth.metagenome_hits(n, pairs, seed=11, scramble=True)
rk = np.concatenate([m1.read_key, m2.read_key]); ut = np.concatenate([m1.unitig, m2.unitig])


def body(comm):
    reads = int(rk.max()) + 1
    lo, hi = reads * comm.rank // comm.world, reads * (comm.rank + 1) // comm.world
    sel = (rk >= lo) & (rk < hi)
    a = torch.from_numpy(rk[sel].view(np.int32).copy()).cuda(); b = torch.from_numpy(ut[sel].view(np.int32).copy()).cuda()
    out = None
    for _ in range(2):
        with DistGraph.from_hits(comm, a, b, n) as g:
            g.coreness()
            out = g.stats()
    return out


for st in run_local(2, body):
    print({k: st[k] for k in ("ms_build", "ms_peel", "peel_levels", "peel_async", "n_messages_sent", "n_directed_local")})
