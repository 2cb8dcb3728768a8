"""Multi-GPU driver of the KOMB hot path: one process per GPU, the graph
partitioned by unitig-id range, exchanges over
torch.distributed (NCCL on GPUs; the same logic runs over gloo on CPU in the
tests with a numpy engine injected).

    stage 1  rank-local hits -> local simple edges -> both directions of every
             edge to the owner of its source (all-to-all) -> sort/unique -> CSR rows
    stage 2  level-synchronous peel; inside a level every rank peels its frontier
             with the CTA-local cascade kernel, ships the decrements of vertices
             other ranks own (all-to-all), applies what it receives, repeats until
             no rank has a frontier or an outbox entry left
    stage 3  all-gather (coreness, degree) -> CORE-A on every rank -> own slice

`Engine` is the per-rank compute interface.  `CudaEngine` binds the partition
entry points of include/kombgpu.h; there is no CPU engine in this package.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int32, c_uint32, c_uint64, c_void_p
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import KEY_REF32
from .api import Context

INT32_MAX = 2**31 - 1


def partition_bounds(n_vertices: int, n_parts: int) -> list[int]:
    """Equal unitig-id ranges: rank j owns [bounds[j], bounds[j+1])."""
    step = -(-n_vertices // n_parts) if n_vertices else 0
    return [min(n_vertices, j * step) for j in range(n_parts)] + [n_vertices]


# ---------------------------------------------------------------------------
# communication
# ---------------------------------------------------------------------------

def _to_host(arr):
    """(numpy view/copy, function that turns a numpy array back into arr's kind)."""
    if hasattr(arr, "data_ptr"):            # torch tensor (object mode on a GPU box: staged through the host)
        import torch
        dev, dt = arr.device, arr.dtype
        return arr.detach().cpu().numpy(), (lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype=dt).to(dev))
    return np.asarray(arr), (lambda a: a)


class Comm:
    """The collectives the path needs, over torch.distributed.

    mode "nccl": arrays are torch CUDA tensors, collectives run on the device.
    mode "object": arrays are numpy arrays moved with all_gather_object (gloo on
    CPU: the multi-rank tests).  A world of one needs no process group."""

    def __init__(self, mode: str | None = None):
        import torch.distributed as dist
        self.dist = dist
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
            self.mode = mode or ("nccl" if dist.get_backend() == "nccl" else "object")
        else:
            self.rank, self.world, self.mode = 0, 1, mode or "object"

    # small host vectors -------------------------------------------------------
    def all_gather_ints(self, values) -> np.ndarray:
        """[world, len(values)] int64 matrix of every rank's vector."""
        row = np.asarray(values, dtype=np.int64)
        if self.world == 1:
            return row[None, :]
        if self.mode == "nccl":
            import torch
            t = torch.from_numpy(row).cuda()
            out = torch.empty((self.world, row.shape[0]), dtype=torch.int64, device="cuda")
            self.dist.all_gather_into_tensor(out, t)
            return out.cpu().numpy()
        got = [None] * self.world
        self.dist.all_gather_object(got, row)
        return np.stack(got)

    # variable all-to-all ------------------------------------------------------
    def all_to_all(self, send, send_counts, recv_counts=None):
        """send is grouped by destination rank with `send_counts` entries each;
        returns what the other ranks sent here, grouped by source rank."""
        send_counts = [int(c) for c in send_counts]
        if self.world == 1:
            return send
        if recv_counts is None:
            recv_counts = self.all_gather_ints(send_counts)[:, self.rank]
        recv_counts = [int(c) for c in recv_counts]
        if self.mode == "nccl":
            import torch
            out = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
            self.dist.all_to_all_single(out, send, recv_counts, send_counts)
            return out
        host, back = _to_host(send)
        offs = np.concatenate([[0], np.cumsum(send_counts)])
        pieces = [host[offs[j]:offs[j + 1]] for j in range(self.world)]
        got = [None] * self.world
        self.dist.all_gather_object(got, pieces)
        return back(np.concatenate([got[src][self.rank] for src in range(self.world)]) if sum(recv_counts) else host[:0])

    def all_gather_var(self, arr):
        """Concatenation of every rank's 1-D array, in rank order."""
        if self.world == 1:
            return arr
        if self.mode == "nccl":
            import torch
            sizes = [int(x) for x in self.all_gather_ints([arr.numel()])[:, 0]]
            out = torch.empty(sum(sizes), dtype=arr.dtype, device=arr.device)
            mx = max(sizes)
            if len(set(sizes)) == 1:
                self.dist.all_gather_into_tensor(out, arr.contiguous())
            elif sum(sizes) * arr.element_size() >= (8 << 20) and mx * self.world <= 1.25 * sum(sizes):
                # large, nearly even segments (the CSR of a balanced range partition): ONE all-gather of pieces
                # padded to the longest (NCCL's ring / NVLS path), then one device copy that drops the padding.
                # The list form below turns into one broadcast per rank, which ran at a third of that bandwidth.
                # Empirical: with a device-wide synchronisation in front of each of the gather stage's collectives the
                # stage takes 7.3 ms at N=4 instead of 13.1 ms (profiles/r1/r1z_bench_n4*.json); the all-gather itself is
                # 2.7 ms for 1.05 GB in the bench and 1.4 ms in isolation (tools/gather_probe.py).  Not understood yet.
                torch.cuda.synchronize()
                pad = torch.empty(self.world * mx, dtype=arr.dtype, device=arr.device)
                src = pad[self.rank * mx:(self.rank + 1) * mx]      # in place: my slot of the padded buffer
                src[:arr.numel()].copy_(arr)
                self.dist.all_gather_into_tensor(pad, src)
                torch.cat([pad[j * mx:j * mx + sizes[j]] for j in range(self.world)], out=out)
            else:
                # uneven segments: every rank's piece lands at its final offset, no padding and no second copy
                offs = np.concatenate([[0], np.cumsum(sizes)])
                self.dist.all_gather([out[int(offs[j]):int(offs[j + 1])] for j in range(self.world)], arr.contiguous())
            return out
        host, back = _to_host(arr)
        got = [None] * self.world
        self.dist.all_gather_object(got, host)
        return back(np.concatenate(got))


# ---------------------------------------------------------------------------
# engine: per-rank compute
# ---------------------------------------------------------------------------

class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr: int, count: int, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}
        self._owner = owner


class CudaEngine:
    """Partition entry points of libkombgpu.so (include/kombgpu.h, 'multi-GPU')."""

    def __init__(self, ctx: Context):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.lib = ctx._lib
        self.device = torch.device("cuda", ctx.device)
        # the exchanges are torch collectives ordered on torch's current stream: the kernels must run there too
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        self.ctx._check(rc)

    def _bounds(self, bounds):
        return (c_uint32 * len(bounds))(*bounds)

    def local_edges(self, read_key, unitig, n_global: int):
        h = c_void_p()
        self._check(self.lib.kombgpu_local_edges_dev(self.ctx._h, c_void_p(read_key.data_ptr()), c_void_p(unitig.data_ptr()),
                                                     read_key.numel(), n_global, byref(h)))
        return h

    def edges_from_pairs(self, u, v, n_global: int):
        h = c_void_p()
        self._check(self.lib.kombgpu_edgeset_from_pairs_dev(self.ctx._h, c_void_p(u.data_ptr()), c_void_p(v.data_ptr()),
                                                            u.numel(), n_global, byref(h)))
        return h

    def edgeset_counts(self, es) -> dict:
        e, p, s = c_uint64(), c_uint64(), c_uint64()
        self._check(self.lib.kombgpu_edgeset_counts(es, byref(e), byref(p), byref(s)))
        return {"n_edges": e.value, "n_pairs": p.value, "n_unique_hits": s.value}

    def route_edges(self, es, bounds):
        n_edges = self.edgeset_counts(es)["n_edges"]
        send = self.torch.empty(2 * n_edges, dtype=self.torch.int64, device=self.device)
        counts = (c_uint64 * (len(bounds) - 1))()
        self._check(self.lib.kombgpu_edgeset_route_dev(es, self._bounds(bounds), len(bounds) - 1, c_void_p(send.data_ptr()), counts))
        self.lib.kombgpu_edgeset_destroy(es)
        return send, list(counts)

    def build_part(self, entries, v_lo: int, v_hi: int, n_global: int):
        h = c_void_p()
        self._check(self.lib.kombgpu_part_build_dev(self.ctx._h, c_void_p(entries.data_ptr() if entries.numel() else 0),
                                                    entries.numel(), v_lo, v_hi, n_global, byref(h)))
        return h

    def part_counts(self, part) -> dict:
        n, d, m = c_uint32(), c_uint64(), c_int32()
        self._check(self.lib.kombgpu_part_counts(part, byref(n), byref(d), byref(m)))
        return {"n_local": n.value, "n_directed": d.value, "max_degree": m.value}

    def peel_begin(self, part):
        self._check(self.lib.kombgpu_part_peel_begin(part))

    def scan(self, part, k: int):
        nf, na, mn = c_uint32(), c_uint32(), c_int32()
        self._check(self.lib.kombgpu_part_peel_scan(part, k, byref(nf), byref(na), byref(mn)))
        return nf.value, na.value, mn.value

    def process(self, part, k: int) -> int:
        no = c_uint32()
        self._check(self.lib.kombgpu_part_peel_process(part, k, byref(no)))
        return no.value

    def route_outbox(self, part, n_outbox: int, bounds):
        send = self.torch.empty(n_outbox, dtype=self.torch.int32, device=self.device)
        counts = (c_uint64 * (len(bounds) - 1))()
        self._check(self.lib.kombgpu_part_outbox_route_dev(part, self._bounds(bounds), len(bounds) - 1,
                                                           c_void_p(send.data_ptr() if n_outbox else 0), counts))
        return send, list(counts)

    def apply(self, part, k: int, recv) -> int:
        nf = c_uint32()
        self._check(self.lib.kombgpu_part_peel_apply_dev(part, k, c_void_p(recv.data_ptr() if recv.numel() else 0), recv.numel(), byref(nf)))
        return nf.value

    def degree_core(self, part):
        """(degree, coreness) of the local vertices as torch tensors (copies)."""
        n = self.part_counts(part)["n_local"]
        ptrs = [c_void_p() for _ in range(4)]
        self._check(self.lib.kombgpu_part_device_arrays(part, *[byref(p) for p in ptrs]))
        if n == 0:
            z = self.torch.zeros(0, dtype=self.torch.int32, device=self.device)
            return z, z.clone()
        deg = self.torch.as_tensor(_DevArray(ptrs[2].value, n, "<i4", part), device=self.device).clone()
        core = self.torch.as_tensor(_DevArray(ptrs[3].value, n, "<i4", part), device=self.device).clone()
        return deg, core

    def part_csr(self, part):
        """(row lengths int32[n_local], col int32[n_directed]) of the local rows: zero-copy views."""
        c = self.part_counts(part)
        ptrs = [c_void_p() for _ in range(4)]
        self._check(self.lib.kombgpu_part_device_arrays(part, *[byref(p) for p in ptrs]))
        z = self.torch.zeros(0, dtype=self.torch.int32, device=self.device)
        deg = self.torch.as_tensor(_DevArray(ptrs[2].value, c["n_local"], "<i4", part), device=self.device) if c["n_local"] else z
        col = self.torch.as_tensor(_DevArray(ptrs[1].value, c["n_directed"], "<i4", part), device=self.device) if c["n_directed"] else z
        return deg, col

    def gather_peel(self, deg_full, col_full, n: int, key_mode: int):
        """Whole graph on this rank (all-gathered rows): persistent single-GPU peel + CORE-A.
        Returns (degree, coreness, score) as torch tensors plus (max_score, max_coreness, stats)."""
        torch = self.torch
        row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        if n:
            torch.cumsum(deg_full.to(torch.int64), 0, out=row_ptr[1:])
        g = self.ctx.graph_from_csr(row_ptr, col_full, n)
        try:
            g.analyse(key_mode)
            arr = g.device_arrays()
            st = g.stats()
            mc, ms = g.summary()
            if n == 0:
                zi = torch.zeros(0, dtype=torch.int32, device=self.device)
                return zi, zi.clone(), torch.zeros(0, dtype=torch.float64, device=self.device), 0.0, 0, st
            core = torch.as_tensor(_DevArray(arr["coreness"], n, "<i4", g), device=self.device).clone()
            score = torch.as_tensor(_DevArray(arr["score"], n, "<f8", g), device=self.device).clone()
            return deg_full, core, score, ms, mc, st
        finally:
            g.close()

    def device_memory_bytes(self) -> int:
        return int(self.torch.cuda.get_device_properties(self.device).total_memory)

    def corea(self, core, deg, key_mode: int):
        n = core.numel()
        score = self.torch.empty(n, dtype=self.torch.float64, device=self.device)
        mx = c_double()
        self._check(self.lib.kombgpu_corea_dev(self.ctx._h, c_void_p(core.data_ptr() if n else 0), c_void_p(deg.data_ptr() if n else 0),
                                               n, key_mode, c_void_p(score.data_ptr() if n else 0), byref(mx)))
        return score, mx.value

    def destroy_part(self, part):
        self.lib.kombgpu_part_destroy(part)

    def sync(self):
        self.torch.cuda.synchronize(self.device)


# ---------------------------------------------------------------------------
# the distributed path
# ---------------------------------------------------------------------------

@dataclass
class DistResult:
    v_lo: int
    v_hi: int
    degree: object          # local slice
    coreness: object        # local slice
    score: object           # local slice
    max_score: float
    max_coreness: int
    n_edges: int            # global simple edges
    stats: dict = field(default_factory=dict)


def analyse_partitioned(engine, comm: Comm, n_global: int, *, read_key=None, unitig=None, pairs=None,
                        key_mode: int = KEY_REF32, timer=None, peel_mode: str = "auto") -> DistResult:
    """Run build -> k-core -> CORE-A over all ranks.  Every rank passes the hits of
    ITS reads (`read_key`, `unitig`: both mates, global unitig ids) or its share of
    an edge list (`pairs=(u, v)`).

    peel_mode
      "partitioned"  the peel runs on the range partition, remote decrements exchanged per
                     sub-round (works for graphs larger than one device's HBM);
      "gather"       the rank-local CSR rows are all-gathered (NVLink) and every rank runs the
                     persistent single-GPU peel on the whole graph.  The peel is bound by its
                     dependency depth, not by bytes, so splitting it over ranks only adds an
                     exchange latency to every one of its sub-rounds; the build stays distributed;
      "auto"         "gather" when the whole CSR takes under a quarter of one device's memory."""
    r, world = comm.rank, comm.world
    bounds = partition_bounds(n_global, world)
    tick = timer or (lambda name: None)

    # ---- stage 1: build -----------------------------------------------------
    es = engine.edges_from_pairs(pairs[0], pairs[1], n_global) if pairs is not None else engine.local_edges(read_key, unitig, n_global)
    es_counts = engine.edgeset_counts(es)
    send, counts = engine.route_edges(es, bounds)
    recv = comm.all_to_all(send, counts)
    part = engine.build_part(recv, bounds[r], bounds[r + 1], n_global)
    pc = engine.part_counts(part)
    tick("build")

    tot = comm.all_gather_ints([pc["n_directed"], es_counts["n_edges"], es_counts["n_pairs"]])
    n_directed_global = int(tot[:, 0].sum())
    base_stats = {"local_edges": es_counts["n_edges"], "local_pairs": es_counts["n_pairs"],
                  "n_directed_local": pc["n_directed"], "sum_local_edges": int(tot[:, 1].sum()),
                  "sum_pairs": int(tot[:, 2].sum())}
    if peel_mode == "auto":
        csr_bytes = 4 * n_directed_global + 24 * n_global
        peel_mode = "gather" if world > 1 and 4 * csr_bytes < engine.device_memory_bytes() else "partitioned"

    if peel_mode == "gather":
        # ---- stage 2+3 on the gathered graph ------------------------------------
        tick("gather.counts")
        deg_loc, col_loc = engine.part_csr(part)
        deg_full = comm.all_gather_var(deg_loc)
        tick("gather.deg")
        col_full = comm.all_gather_var(col_loc)
        tick("gather.col")
        engine.destroy_part(part)
        tick("gather")
        deg_f, core_f, score_f, max_score, max_core, gst = engine.gather_peel(deg_full, col_full, n_global, key_mode)
        tick("peel+corea")
        lo, hi = bounds[r], bounds[r + 1]
        return DistResult(lo, hi, deg_f[lo:hi], core_f[lo:hi], score_f[lo:hi], float(max_score), int(max_core),
                          n_directed_global // 2,
                          dict(base_stats, peel_mode="gather", levels=int(gst.get("peel_levels", 0)), exchange_subrounds=0,
                               decrements_received=0, ms_peel=float(gst.get("ms_peel", 0.0)), ms_corea=float(gst.get("ms_corea", 0.0)),
                               ms_peel_kernel=float(gst.get("ms_peel_kernel", 0.0))))

    # ---- stage 2: peel --------------------------------------------------------
    engine.peel_begin(part)
    k, levels, subrounds, exchanged = 0, 0, 0, 0
    max_core = 0
    while True:
        nf, na, mn = engine.scan(part, k)
        g = comm.all_gather_ints([nf, na, mn])
        tot_front, tot_alive, gmin = int(g[:, 0].sum()), int(g[:, 1].sum()), int(g[:, 2].min())
        if tot_front == 0:
            if tot_alive == 0 or gmin == INT32_MAX:
                break
            k = gmin
            continue
        levels += 1
        max_core = k
        while True:
            n_out = engine.process(part, k)
            send, counts = engine.route_outbox(part, n_out, bounds)
            matrix = comm.all_gather_ints(counts)            # matrix[src, dst]
            if int(matrix.sum()) == 0:
                break
            recv = comm.all_to_all(send, counts, matrix[:, r])
            exchanged += int(matrix[:, r].sum())
            engine.apply(part, k, recv)
            subrounds += 1
        k += 1
    tick("peel")

    # ---- stage 3: CORE-A --------------------------------------------------------
    deg, core = engine.degree_core(part)
    deg_full = comm.all_gather_var(deg)
    core_full = comm.all_gather_var(core)
    score_full, max_score = engine.corea(core_full, deg_full, key_mode)
    score = score_full[bounds[r]:bounds[r + 1]]
    tick("corea")

    engine.destroy_part(part)
    return DistResult(bounds[r], bounds[r + 1], deg, core, score, float(max_score), max_core, n_directed_global // 2,
                      dict(base_stats, peel_mode="partitioned", levels=levels, exchange_subrounds=subrounds,
                           decrements_received=exchanged))
