"""A numpy stand-in for komb_b200.distributed.CudaEngine, used ONLY by the CPU
(gloo, world_size > 1) tests of the multi-rank host logic.  It restates the
per-rank partition semantics of include/kombgpu.h's multi-GPU section on top of
the CPU oracle; it is test infrastructure, not a fallback."""
import numpy as np

from oracle import oracle

INT32_MAX = 2**31 - 1


class Part:
    pass


class NumpyEngine:
    def local_edges(self, read_key, unitig, n_global):
        edges, p, s = oracle.build_edges(read_key, unitig)
        return {"edges": edges, "n_pairs": p, "n_unique_hits": s}

    def edges_from_pairs(self, u, v, n_global):
        return {"edges": oracle.simplify(u, v), "n_pairs": len(u), "n_unique_hits": 0}

    def edgeset_counts(self, es):
        return {"n_edges": int(es["edges"].shape[0]), "n_pairs": int(es["n_pairs"]), "n_unique_hits": int(es["n_unique_hits"])}

    def route_edges(self, es, bounds):
        e = es["edges"]
        u, v = oracle.unpack_edges(e)
        directed = np.concatenate([oracle.pack_edges(u, v), oracle.pack_edges(v, u)])
        src = (directed >> np.uint64(32)).astype(np.int64)
        owner = np.searchsorted(np.asarray(bounds[1:]), src, side="right")
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=len(bounds) - 1)[:len(bounds) - 1]
        return directed[order].view(np.int64), [int(c) for c in counts]

    def build_part(self, entries, v_lo, v_hi, n_global):
        p = Part()
        ent = np.unique(np.asarray(entries).view(np.uint64))
        src = (ent >> np.uint64(32)).astype(np.int64) - v_lo
        assert src.size == 0 or (src.min() >= 0 and src.max() < v_hi - v_lo)
        p.v_lo, p.n_local = v_lo, v_hi - v_lo
        p.col = (ent & np.uint64(0xFFFFFFFF)).astype(np.int64)
        p.deg = np.bincount(src, minlength=p.n_local).astype(np.int32)
        p.row_ptr = np.concatenate([[0], np.cumsum(p.deg)]).astype(np.int64)
        return p

    def part_counts(self, p):
        return {"n_local": p.n_local, "n_directed": int(p.col.shape[0]), "max_degree": int(p.deg.max()) if p.n_local else 0}

    def peel_begin(self, p):
        p.core = p.deg.copy()
        p.alive = np.arange(p.n_local)
        p.front = np.zeros(0, dtype=np.int64)
        p.outbox = []

    def scan(self, p, k):
        d = p.core[p.alive]
        p.front = p.alive[d == k]
        p.alive = p.alive[d > k]
        mn = int(p.core[p.alive].min()) if p.alive.size else INT32_MAX
        return int(p.front.size), int(p.alive.size), mn

    def _decrement(self, p, k, loc, queue):
        if p.core[loc] > k:
            p.core[loc] -= 1
            if p.core[loc] == k:
                queue.append(loc)

    def process(self, p, k):
        queue = list(p.front.tolist())
        p.front = np.zeros(0, dtype=np.int64)
        p.outbox = []
        while queue:
            v = queue.pop()
            for u in p.col[p.row_ptr[v]:p.row_ptr[v + 1]].tolist():
                loc = u - p.v_lo
                if 0 <= loc < p.n_local:
                    self._decrement(p, k, loc, queue)
                else:
                    p.outbox.append(u)
        return len(p.outbox)

    def route_outbox(self, p, n_outbox, bounds):
        ob = np.asarray(p.outbox, dtype=np.int64)
        owner = np.searchsorted(np.asarray(bounds[1:]), ob, side="right")
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=len(bounds) - 1)[:len(bounds) - 1]
        return ob[order].astype(np.int32), [int(c) for c in counts]

    def apply(self, p, k, recv):
        queue = []
        for u in np.asarray(recv).astype(np.int64).tolist():
            self._decrement(p, k, u - p.v_lo, queue)
        p.front = np.asarray(queue, dtype=np.int64)
        return len(queue)

    def degree_core(self, p):
        return p.deg.copy(), p.core.copy()

    def corea(self, core, deg, key_mode):
        s = oracle.corea(np.asarray(core), np.asarray(deg), key_mode)
        return s, float(s.max()) if s.size else 0.0

    def part_csr(self, p):
        return p.deg.copy(), p.col.astype(np.int32)

    def device_memory_bytes(self):
        return 1 << 34

    def gather_peel(self, deg_full, col_full, n, key_mode):
        deg_full = np.asarray(deg_full).astype(np.int64)
        col = np.asarray(col_full).astype(np.int64)
        src = np.repeat(np.arange(n, dtype=np.int64), deg_full)
        keep = src < col
        edges = np.sort(oracle.pack_edges(src[keep].astype(np.uint32), col[keep].astype(np.uint32)))
        deg, core = oracle.coreness(n, edges)
        assert np.array_equal(deg, deg_full)
        score = oracle.corea(core, deg, key_mode)
        return deg, core, score, float(score.max()) if n else 0.0, int(core.max()) if n else 0, {"peel_levels": len(set(core.tolist()))}

    def destroy_part(self, p):
        pass
