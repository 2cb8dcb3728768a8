"""Multi-rank host logic on CPU: torch.distributed (gloo), world_size 2 and 3, with a
numpy engine injected in place of the CUDA partition kernels.  Checks the
partitioning, both exchanges, the level / sub-round termination logic and the
result assembly of komb_b200.distributed against the single-process oracle."""
import os
import pickle
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, case, peel_mode="partitioned"):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    from dist_numpy_engine import NumpyEngine
    from komb_b200 import synth
    from komb_b200.distributed import Comm, analyse_partitioned
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = Comm()
    assert comm.mode == "object" and comm.world == world
    n, reads, seed, kind = case
    if kind == "hits":
        # every rank owns a contiguous range of reads (both mates)
        m1, m2 = synth.metagenome_hits(n, reads, seed=seed)
        lo, hi = reads * rank // world, reads * (rank + 1) // world
        sel1 = (m1.read_key >= lo) & (m1.read_key < hi)
        sel2 = (m2.read_key >= lo) & (m2.read_key < hi)
        rk = np.concatenate([m1.read_key[sel1], m2.read_key[sel2]])
        ut = np.concatenate([m1.unitig[sel1], m2.unitig[sel2]])
        res = analyse_partitioned(NumpyEngine(), comm, n, read_key=rk, unitig=ut, peel_mode=peel_mode)
    else:
        u, v = synth.rmat_edges(11, reads, n_vertices=n, seed=seed)
        sl = slice(len(u) * rank // world, len(u) * (rank + 1) // world)
        res = analyse_partitioned(NumpyEngine(), comm, n, pairs=(u[sl], v[sl]), key_mode=1, peel_mode=peel_mode)
    with open(Path(out_dir) / f"rank{rank}.pkl", "wb") as f:
        pickle.dump({"v_lo": res.v_lo, "v_hi": res.v_hi, "deg": np.asarray(res.degree), "core": np.asarray(res.coreness),
                     "score": np.asarray(res.score), "max_score": res.max_score, "max_core": res.max_coreness,
                     "n_edges": res.n_edges, "stats": res.stats}, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case,peel_mode", [(2, (1500, 4000, 3, "hits"), "partitioned"), (3, (1000, 3000, 4, "hits"), "partitioned"),
                                                  (2, (1200, 9000, 5, "pairs"), "partitioned"), (2, (1500, 4000, 3, "hits"), "auto"),
                                                  (3, (1200, 9000, 5, "pairs"), "gather")])
def test_partitioned_path_matches_single_process(tmp_path, oracle_mod, world, case, peel_mode):
    from komb_b200 import synth
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), case, peel_mode), nprocs=world, join=True)
    parts = [pickle.load(open(tmp_path / f"rank{r}.pkl", "rb")) for r in range(world)]
    n, reads, seed, kind = case
    if kind == "hits":
        m1, m2 = synth.metagenome_hits(n, reads, seed=seed)
        edges, _, _ = oracle_mod.build_edges(np.concatenate([m1.read_key, m2.read_key]), np.concatenate([m1.unitig, m2.unitig]))
        mode = oracle_mod.KEY_REF32
    else:
        u, v = synth.rmat_edges(11, reads, n_vertices=n, seed=seed)
        edges = oracle_mod.simplify(u, v)
        mode = oracle_mod.KEY_EXACT64
    deg, core = oracle_mod.coreness(n, edges)
    score = oracle_mod.corea(core, deg, mode)
    assert [p["v_lo"] for p in parts] + [parts[-1]["v_hi"]] == __import__("komb_b200.distributed", fromlist=["x"]).partition_bounds(n, world)
    assert np.array_equal(np.concatenate([p["deg"] for p in parts]), deg)
    assert np.array_equal(np.concatenate([p["core"] for p in parts]), core)          # bit-exact, whatever the partition
    np.testing.assert_allclose(np.concatenate([p["score"] for p in parts]), score, rtol=1e-6, atol=1e-12)
    for p in parts:
        assert p["n_edges"] == edges.shape[0] and p["max_core"] == core.max()
        assert abs(p["max_score"] - score.max()) < 1e-12
    if peel_mode == "partitioned":
        assert parts[0]["stats"]["exchange_subrounds"] > 0      # the peel really crossed partitions
    else:
        assert parts[0]["stats"]["peel_mode"] == "gather"


def test_single_rank_needs_no_process_group(oracle_mod):
    sys.path.insert(0, str(ROOT / "tests"))
    from dist_numpy_engine import NumpyEngine
    from komb_b200 import synth
    from komb_b200.distributed import Comm, analyse_partitioned
    u, v = synth.rmat_edges(9, 3000, n_vertices=400, seed=1)
    res = analyse_partitioned(NumpyEngine(), Comm(), 400, pairs=(u, v), peel_mode="partitioned")
    deg, core = oracle_mod.coreness(400, oracle_mod.simplify(u, v))
    assert np.array_equal(res.coreness, core) and np.array_equal(res.degree, deg)
    assert res.stats["exchange_subrounds"] == 0


def test_async_peel_level_rule_model():
    """A numpy model of the level rule of the asynchronous partitioned peel (komb_b200/csrc/apeel.cu): the next level is the
    global minimum of per-rank LOWER BOUNDS -- the survivors' degrees as the last scan saw them and every degree a decrement
    left above the level -- not of the true surviving degrees.  A unitig that died since can make the bound too low (the
    level it names is then empty: a wasted scan), never too high (a level skipped: a wrong coreness).  The model peels
    partitioned graphs level by level with that rule, decrements applied in an arbitrary (shuffled) order inside a
    level, and must reproduce the oracle's coreness; it also checks bound <= true minimum at every level."""
    from komb_b200 import synth
    from oracle import oracle
    rng = np.random.default_rng(4)
    for seed, n, m, world in [(1, 400, 3000, 2), (2, 1500, 12000, 3), (3, 90, 300, 4)]:
        u, v = synth.rmat_edges(11, m, n_vertices=n, seed=seed)
        edges = oracle.simplify(u, v)
        exp_deg, exp_core = oracle.coreness(n, edges)
        eu, ev = oracle.unpack_edges(edges)
        adj = [[] for _ in range(n)]
        for a, b in zip(eu.tolist(), ev.tolist()):
            adj[a].append(b); adj[b].append(a)
        step = (n + world - 1) // world
        owner = lambda x: x // step                                     # noqa: E731
        deg = exp_deg.astype(np.int64).copy()
        core = np.full(n, -1, np.int64)
        alive = [list(range(q * step, min((q + 1) * step, n))) for q in range(world)]
        bound = [min((deg[x] for x in alive[q]), default=None) for q in range(world)]   # before the first level: a pass
        k_prev, empty_levels, levels = -1, 0, 0
        while True:
            lbs = [b for b in bound if b is not None]
            if not lbs:
                break
            k = min(lbs)
            assert k > k_prev
            true_min = min((int(deg[x]) for q in range(world) for x in alive[q] if core[x] < 0), default=None)
            assert true_min is None or k <= true_min                    # never above the true next level
            k_prev = k
            levels += 1
            # scan: unitigs at degree k are pushed, survivors kept; the scan contributes the survivors' degrees to the bound
            bound = [None] * world
            pool = []
            for q in range(world):
                keep = []
                for x in alive[q]:
                    if core[x] >= 0 or deg[x] < k:
                        continue                                         # peeled at an earlier level
                    if deg[x] == k:
                        pool.append(x)
                    else:
                        keep.append(x)
                        bound[q] = int(deg[x]) if bound[q] is None else min(bound[q], int(deg[x]))
                alive[q] = keep
            if not pool:
                empty_levels += 1
            # the level: entries are taken in arbitrary order; a decrement that leaves old - 1 > k lowers the DECREMENTING
            # rank's bound; the one that takes a degree from k + 1 to k pushes the unitig
            while pool:
                x = pool.pop(int(rng.integers(0, len(pool))))
                assert core[x] < 0
                core[x] = k
                q = owner(x)
                for y in adj[x]:
                    old = int(deg[y])
                    deg[y] = old - 1
                    if old == k + 1:
                        pool.append(y)
                    elif old > k + 1:
                        bound[q] = old - 1 if bound[q] is None else min(bound[q], old - 1)
        assert np.array_equal(core, exp_core.astype(np.int64))
        assert levels - empty_levels == len(set(exp_core.tolist()))      # every non-empty level was visited exactly once
        assert empty_levels <= levels                                    # (stale bounds may add empty ones)


def test_async_peel_end_of_level_detection_model():
    """A randomised interleaving model of the end-of-level test of the asynchronous partitioned peel (apeel.cu): every
    rank counts entries pushed into its pool (tail: bumped by the PRODUCER, at the target, before the producer's own entry
    counts as done) and entries processed (done).  A manager reads done then tail of every rank, then every tail once
    more; equal and unchanged must imply that at some instant nobody had work.  The model interleaves worker steps and
    the manager's individual reads at random and checks that the manager never declares the end while an entry is
    queued or in flight -- and that it does declare it once everything is processed."""
    rng = np.random.default_rng(12)
    for trial in range(300):
        world = int(rng.integers(2, 5))
        tail = [0] * world; done = [0] * world; head = [0] * world
        # initial frontier (the scan's pushes) and a budget of discoveries each processed entry may still make
        for q in range(world):
            tail[q] = int(rng.integers(0, 4))
        budget = int(rng.integers(0, 40))
        if sum(tail) == 0:
            tail[0] = 1
        in_flight = []            # entries taken by a worker: [rank, pushes still to make]
        declared = False
        reads = {}                # the manager's collect in progress
        plan = []                 # remaining reads of the current collect: ("d", q), ("t", q), ("t2", q)
        steps = 0
        while not declared:
            steps += 1
            assert steps < 100000
            work_left = any(head[q] < tail[q] for q in range(world)) or bool(in_flight)
            choice = rng.random()
            if work_left and choice < 0.6:
                movable = [("take", q) for q in range(world) if head[q] < tail[q]] + [("step", i) for i in range(len(in_flight))]
                kind, x = movable[int(rng.integers(0, len(movable)))]
                if kind == "take":
                    head[x] += 1
                    pushes = int(rng.integers(0, 3)) if budget > 0 else 0
                    budget -= pushes
                    in_flight.append([x, pushes])
                else:
                    ent = in_flight[x]
                    if ent[1] > 0:                      # a discovery: the TARGET's tail first ...
                        tail[int(rng.integers(0, world))] += 1
                        ent[1] -= 1
                    else:                               # ... the entry's own done last
                        done[ent[0]] += 1
                        in_flight.pop(x)
            else:
                if not plan:                            # start a collect: all done reads, then all tail reads (any order inside)
                    d_reads = [("d", q) for q in range(world)]; t_reads = [("t", q) for q in range(world)]
                    rng.shuffle(d_reads); rng.shuffle(t_reads)
                    # the two reads of ONE rank may be reordered by the hardware: model that by mixing the groups
                    first = d_reads + t_reads
                    if rng.random() < 0.5:
                        rng.shuffle(first)
                    second = [("t2", q) for q in range(world)]
                    rng.shuffle(second)
                    plan = first + ["check1"] + second + ["check2"]
                    reads = {}
                op = plan.pop(0)
                if op == "check1":
                    if any(reads[("d", q)] != reads[("t", q)] for q in range(world)):
                        plan = []                        # somebody busy: start over
                elif op == "check2":
                    if all(reads[("t2", q)] == reads[("t", q)] for q in range(world)):
                        # declared: there must be no work anywhere, now or (by stability) at the instant it refers to
                        assert not any(head[q] < tail[q] for q in range(world)) and not in_flight, (trial, tail, done, head)
                        assert done == tail
                        declared = True
                    plan = []
                else:
                    kind, q = op
                    reads[op] = done[q] if kind == "d" else tail[q]
        assert sum(done) == sum(tail)
