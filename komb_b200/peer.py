"""Python face of the peer-memory multi-GPU path (include/kombgpu.h, "multi-GPU, peer-memory path"): a thin
ctypes mirror.  All compute — partitioned build, device-driven peel over NVLink mailboxes, sharded CORE-A — is
in libkombgpu.so; Python only provides the bootstrap all-gather when the ranks are separate processes.

    one process per GPU (torchrun):   comm = Comm.from_torch(ctx)          # bootstrap over torch.distributed
    one process, a thread per GPU:    run_local(world, fn, devices=[...])  # bootstrap built into the library
    tests on a one-GPU box:           run_local(world, fn)                 # every rank on device 0 (emulation)
"""
from __future__ import annotations

import ctypes
import threading
from ctypes import byref, c_double, c_int, c_int32, c_uint32, c_uint64, c_void_p

import numpy as np

from . import _lib
from ._lib import KEY_REF32, DistStats
from .api import Context


class Comm:
    """kombgpu_comm: one rank of the peer-memory communicator."""

    def __init__(self, ctx: Context, handle: c_void_p, keep=None):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = handle
        self._keep = keep
        r, w, sd, hb = c_int(), c_int(), c_int(), c_uint64()
        ctx._check(self._lib.kombgpu_comm_info(self._h, byref(r), byref(w), byref(sd), byref(hb)))
        self.rank, self.world, self.same_device = r.value, w.value, bool(sd.value)

    @classmethod
    def create(cls, ctx: Context, rank: int, world: int, allgather, heap_bytes: int = 0) -> "Comm":
        """`allgather(data: bytes) -> list[bytes]` gathers one blob per rank, in rank order (any transport)."""
        def _cb(_user, send, recv, nbytes):
            try:
                parts = allgather(ctypes.string_at(send, nbytes))
                blob = b"".join(parts)
                if len(blob) != nbytes * world:
                    return -1
                ctypes.memmove(recv, blob, len(blob))
                return 0
            except Exception:   # the C side turns this into KOMBGPU_ESTATE
                return -1
        cb = _lib.ALLGATHER_FN(_cb)
        h = c_void_p()
        ctx._check(ctx._lib.kombgpu_comm_create(ctx._h, rank, world, cb, None, int(heap_bytes), byref(h)))
        return cls(ctx, h, keep=(cb, allgather))

    @classmethod
    def from_torch(cls, ctx: Context, heap_bytes: int = 0) -> "Comm":
        """Ranks = the processes of the initialised torch.distributed group (one GPU each)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()

        def gather(data: bytes):
            out = [None] * world
            dist.all_gather_object(out, data)
            return out
        return cls.create(ctx, rank, world, gather, heap_bytes)

    @classmethod
    def create_local(cls, ctx: Context, rank: int, world: int, group_slot: c_void_p, heap_bytes: int = 0) -> "Comm":
        h = c_void_p()
        ctx._check(ctx._lib.kombgpu_comm_create_local(ctx._h, rank, world, byref(group_slot), int(heap_bytes), byref(h)))
        return cls(ctx, h, keep=group_slot)

    def heap_bytes(self) -> int:
        hb = c_uint64()
        self.ctx._check(self._lib.kombgpu_comm_info(self._h, None, None, None, byref(hb)))
        return hb.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kombgpu_comm_destroy(self._h)
            self._h = None


class DistGraph:
    """kombgpu_dist_graph: this rank's share of the partitioned graph."""

    def __init__(self, comm: Comm, handle: c_void_p, keep=None):
        self.comm = comm
        self._ctx = comm.ctx
        self._lib = comm._lib
        self._h = handle
        self._keep = keep

    @classmethod
    def _build(cls, fn, comm: Comm, a, b, n_global: int) -> "DistGraph":
        assert a.is_cuda and b.is_cuda and a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
        h = c_void_p()
        comm.ctx._check(fn(comm._h, c_void_p(a.data_ptr() if a.numel() else 0), c_void_p(b.data_ptr() if b.numel() else 0),
                           a.numel(), int(n_global), byref(h)))
        return cls(comm, h, keep=(a, b))

    @classmethod
    def from_hits(cls, comm: Comm, read_key, unitig, n_global: int) -> "DistGraph":
        """Hits of THIS rank's reads (torch CUDA int32/uint32 tensors, global unitig ids).  Collective."""
        return cls._build(comm._lib.kombgpu_dist_build_hits_dev, comm, read_key, unitig, n_global)

    @classmethod
    def from_pairs(cls, comm: Comm, u, v, n_global: int) -> "DistGraph":
        """This rank's share of an edge list.  Collective."""
        return cls._build(comm._lib.kombgpu_dist_build_pairs_dev, comm, u, v, n_global)

    def coreness(self):
        self._ctx._check(self._lib.kombgpu_dist_coreness(self._h))

    def corea(self, key_mode: int = KEY_REF32):
        self._ctx._check(self._lib.kombgpu_dist_corea(self._h, int(key_mode)))

    def analyse(self, key_mode: int = KEY_REF32):
        self.coreness()
        self.corea(key_mode)

    def stats(self) -> dict:
        st = DistStats()
        self._ctx._check(self._lib.kombgpu_dist_graph_stats(self._h, byref(st)))
        return st.as_dict()

    def summary(self) -> tuple[int, float]:
        mc, ms = c_int32(), c_double()
        self._ctx._check(self._lib.kombgpu_dist_graph_summary(self._h, byref(mc), byref(ms)))
        return mc.value, ms.value

    def results(self, out: dict | None = None) -> dict:
        """degree / coreness / score of the local unitigs [v_lo, v_lo + n_local) on the host."""
        st = self.stats()
        n = st["n_local"]
        out = out or {}
        r = {"degree": out.get("degree", np.empty(n, np.int32))[:n], "coreness": out.get("coreness", np.empty(n, np.int32))[:n],
             "score": out.get("score", np.empty(n, np.float64))[:n]}
        if any(a.shape[0] != n for a in r.values()):
            raise ValueError("out= buffers are too small for this rank's unitigs")
        self._ctx._check(self._lib.kombgpu_dist_graph_results(self._h, c_void_p(r["degree"].ctypes.data), c_void_p(r["coreness"].ctypes.data),
                                                              c_void_p(r["score"].ctypes.data)))
        r["v_lo"] = st["v_lo"]
        return r

    def edges(self, with_mult: bool = False):
        """This rank's slice of the canonical edge list (u, v[, mult])."""
        m = self.stats()["n_fwd_local"]
        u, v = np.empty(m, np.uint32), np.empty(m, np.uint32)
        mult = np.empty(m, np.uint32) if with_mult else None
        self._ctx._check(self._lib.kombgpu_dist_graph_edges(self._h, c_void_p(u.ctypes.data), c_void_p(v.ctypes.data),
                                                            c_void_p(mult.ctypes.data) if with_mult else None))
        return (u, v, mult) if with_mult else (u, v)

    def edges_csr(self, out: dict | None = None):
        """This rank's slice of the edge list as (fwd_ptr uint64[n_local+1], v uint32[n_fwd_local])."""
        st = self.stats()
        out = out or {}
        fp = out.get("fwd_ptr", np.empty(st["n_local"] + 1, np.uint64))[:st["n_local"] + 1]
        v = out.get("v", np.empty(st["n_fwd_local"], np.uint32))[:st["n_fwd_local"]]
        if fp.shape[0] != st["n_local"] + 1 or v.shape[0] != st["n_fwd_local"]:
            raise ValueError("out= buffers are too small for this rank's slice of the edge list")
        self._ctx._check(self._lib.kombgpu_dist_graph_edges_csr(self._h, c_void_p(fp.ctypes.data), c_void_p(v.ctypes.data)))
        return fp, v

    def device_arrays(self) -> dict:
        ptrs = [c_void_p() for _ in range(6)]
        self._ctx._check(self._lib.kombgpu_dist_graph_device_arrays(self._h, *[byref(p) for p in ptrs]))
        names = ["row_ptr", "col", "edges_packed", "degree", "coreness", "score"]
        return {k: (p.value or 0) for k, p in zip(names, ptrs)}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kombgpu_dist_graph_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def run_local(world: int, fn, devices=None, heap_bytes: int = 0):
    """Run `fn(comm) -> result` on `world` ranks inside this process, one host thread per rank (ctypes releases the
    GIL inside the library).  devices=None puts every rank on device 0: the emulation mode for one-GPU boxes."""
    devices = list(devices) if devices is not None else [0] * world
    assert len(devices) == world
    slot = c_void_p(None)
    results, errors = [None] * world, [None] * world

    def worker(rank):
        ctx = comm = None
        try:
            ctx = Context(devices[rank])
            comm = Comm.create_local(ctx, rank, world, slot, heap_bytes)
            results[rank] = fn(comm)
        except BaseException as e:   # noqa: BLE001 - reported to the caller below
            errors[rank] = e
            if comm is not None:
                comm._lib.kombgpu_comm_abort(comm._h)   # the other ranks must not wait for this one
        finally:
            if comm is not None:
                comm.close()
            if ctx is not None:
                ctx.close()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results
