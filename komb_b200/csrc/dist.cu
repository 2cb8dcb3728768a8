// dist.cu — one rank's share of the multi-GPU path (graph partitioned by unitig-id
// range).  The collectives between these calls are issued
// by the host (komb_b200/distributed.py: torch.distributed over NCCL).
//
//   build  local hits -> local simple edges -> both directions of every edge routed
//          to the owner of the source -> (all-to-all) -> sort + unique -> CSR rows
//   peel   level-synchronous over all ranks; inside a level each rank runs the same
//          CTA-local cascade as the single-GPU kernel (peel_device.cuh, kDist=true):
//          neighbours it owns are decremented in place, neighbours other ranks own
//          go to an outbox that the host exchanges; the owner applies them and the
//          vertices that reach k seed the next sub-round.
// Decrements are clamped at k exactly as on one GPU, so the result does not depend
// on the partition or on the order in which decrements arrive (bit-exact).
#include <cstdlib>

#include "peel_device.cuh"
#include "primitives.cuh"

namespace cg = cooperative_groups;

struct kombgpu_edgeset {
    kombgpu_ctx *ctx = nullptr;
    uint32_t n_global = 0;
    uint64_t n_edges = 0;
    uint64_t *edges = nullptr;  // [n_edges] sorted unique (u << 32 | v), u < v
    kombgpu_stats st{};
};

struct kombgpu_part {
    kombgpu_ctx *ctx = nullptr;
    uint32_t v_lo = 0, n_local = 0, n_global = 0;
    uint64_t n_directed = 0;
    int32_t max_degree = 0;
    uint64_t *row_ptr = nullptr;  // [n_local + 1]
    uint32_t *col = nullptr;      // [n_directed] global ids
    int32_t *deg = nullptr;       // [n_local]
    int32_t *core = nullptr;      // [n_local] working degrees / coreness
    // peel state
    uint32_t *alive[2] = {nullptr, nullptr}, *outbox = nullptr;
    uint64_t *pool = nullptr;     // task pool of the peel (peel_device.cuh)
    uint32_t pool_cap = 0;
    kg::peel::PeelState *state = nullptr;
    uint32_t *counters = nullptr;  // [4]: outbox_cnt, front_cnt (apply), spare, spare
    uint32_t n_alive = 0;
    int alive_cur = -1;            // -1: identity list
    uint32_t n_front = 0;          // entries waiting in `frontier`
    uint32_t n_outbox = 0;
    int grid = 0;
    uint32_t round = 0;            // process launches so far (names the level token)
};

namespace kg {
namespace {

using namespace peel;
constexpr int kThreads = 256;

// level-k scan of one rank (stand-alone launch; state slot 0 holds the results)
__global__ void __launch_bounds__(kPeelThreads) part_scan_kernel(int32_t k, const uint32_t *alive_src, uint32_t n_alive,
                                                                 uint32_t *alive_dst, const int32_t *deg, uint64_t *Q,
                                                                 PeelState *st) {
    __shared__ BlockShared sh;
    int32_t local_min = scan_alive(k, alive_src, n_alive, alive_dst, deg, Q, &st->q_tail, &st->front_cnt[0], &st->alive_out[0], &st->n_isolated, sh);
    local_min = warp_reduce_min(local_min);
    if (lane_id() == 0 && local_min != INT32_MAX) atomicMin(&st->next_min[0], local_min);
}

// local PROCESS of level k: everything in the pool (frontier from the scan or from part_apply_kernel) plus the
// local cascade it triggers.  Cooperative (all CTAs resident: they hand work to one another through the pool).
__global__ void __launch_bounds__(kPeelThreads) part_process_kernel(int32_t k, uint32_t round,
                                                                    const uint64_t *__restrict__ row_ptr,
                                                                    const uint32_t *__restrict__ col, int32_t *deg,
                                                                    uint64_t *Q, uint32_t cap, PeelState *st, PartView part) {
    cg::grid_group grid = cg::this_grid();
    __shared__ BlockShared sh;
    const uint32_t removed = process_level<true>(k, round, Q, cap, row_ptr, col, deg, st, sh, part);
    if (threadIdx.x == 0 && removed) atomicAdd(&st->n_removed, (unsigned long long)removed);
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) st->q_head = __ldcg(&st->q_done);  // un-reserve the slots past the tail
}

// decrements that arrived from other ranks (global ids of vertices this rank owns); vertices that reach k
// are appended to the pool
__global__ void __launch_bounds__(kThreads) part_apply_kernel(int32_t k, const uint32_t *__restrict__ recv, uint64_t count,
                                                              uint32_t v_lo, uint32_t n_local, int32_t *deg, uint64_t *Q,
                                                              uint32_t cap, PeelState *st, uint32_t *front_cnt) {
    const uint32_t lane = lane_id();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += stride) {
        const uint64_t i = base + lane;
        bool push = false;
        uint32_t loc = 0;
        if (i < count) {
            loc = recv[i] - v_lo;
            if (loc >= n_local) {
                atomicExch(&st->error, 5u);
            } else if (__ldcg(&deg[loc]) > k) {
                const int32_t old = atomicSub(&deg[loc], 1);
                if (old == k + 1) push = true;
                else if (old <= k) atomicAdd(&deg[loc], 1);
            }
        }
        const uint32_t pm = __ballot_sync(kFullMask, push);
        if (pm) {
            uint32_t pos = 0;
            if (lane == 0) { pos = atomicAdd(&st->q_tail, (uint32_t)__popc(pm)); atomicAdd(front_cnt, (uint32_t)__popc(pm)); }
            pos = __shfl_sync(kFullMask, pos, 0) + __popc(pm & lanemask_lt());
            if (push) {
                if (pos < cap) Q[pos] = (uint64_t)loc;
                else atomicExch(&st->error, 3u);
            }
        }
    }
}

constexpr int kMaxParts = 64;
struct Bounds {
    uint32_t b[kMaxParts + 1];
    int n;
};
__device__ __forceinline__ int owner_of(const Bounds &bd, uint32_t v) {
    int lo = 0, hi = bd.n;  // b[lo] <= v < b[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (bd.b[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads) outbox_count_kernel(const uint32_t *__restrict__ outbox, uint32_t count, Bounds bd,
                                                                uint32_t *__restrict__ counts) {
    __shared__ uint32_t s_cnt[kMaxParts];
    for (int j = threadIdx.x; j < kMaxParts; j += kThreads) s_cnt[j] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&s_cnt[owner_of(bd, outbox[i])], 1u);
    __syncthreads();
    for (int j = threadIdx.x; j < bd.n; j += kThreads)
        if (s_cnt[j]) atomicAdd(&counts[j], s_cnt[j]);
}

// cursors[j] starts at the offset of owner j's group in `send`
__global__ void __launch_bounds__(kThreads) outbox_scatter_kernel(const uint32_t *__restrict__ outbox, uint32_t count,
                                                                  Bounds bd, uint32_t *__restrict__ cursors,
                                                                  uint32_t *__restrict__ send) {
    __shared__ uint32_t s_cnt[kMaxParts];
    __shared__ uint32_t s_base[kMaxParts];
    const uint64_t tile = (uint64_t)blockDim.x * 8;
    for (uint64_t t0 = (uint64_t)blockIdx.x * tile; t0 < count; t0 += (uint64_t)gridDim.x * tile) {
        for (int j = threadIdx.x; j < kMaxParts; j += kThreads) s_cnt[j] = 0;
        __syncthreads();
        uint32_t v[8], own[8], rank[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint64_t i = t0 + (uint64_t)j * blockDim.x + threadIdx.x;
            own[j] = 0xffffffffu;
            if (i < count) {
                v[j] = outbox[i];
                own[j] = (uint32_t)owner_of(bd, v[j]);
                rank[j] = atomicAdd(&s_cnt[own[j]], 1u);
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < bd.n; j += kThreads) s_base[j] = s_cnt[j] ? atomicAdd(&cursors[j], s_cnt[j]) : 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (own[j] != 0xffffffffu) send[s_base[own[j]] + rank[j]] = v[j];
        __syncthreads();
    }
}

int make_bounds(kombgpu_ctx *ctx, const uint32_t *bounds, int n_parts, Bounds *out) {
    if (!bounds || n_parts < 1 || n_parts > kMaxParts) return ctx_fail(ctx, KOMBGPU_EINVAL, "n_parts must be in [1, %d]", kMaxParts);
    for (int j = 0; j <= n_parts; ++j) {
        if (j && bounds[j] < bounds[j - 1]) return ctx_fail(ctx, KOMBGPU_EINVAL, "partition bounds must be non-decreasing");
        out->b[j] = bounds[j];
    }
    out->n = n_parts;
    return KOMBGPU_OK;
}

void part_release(kombgpu_part *p) {
    kombgpu_ctx *ctx = p->ctx;
    void *ptrs[] = {p->row_ptr, p->col, p->deg, p->core, p->alive[0], p->alive[1], p->outbox, p->pool, p->state, p->counters};
    for (void *q : ptrs) if (q) ws_free(ctx, q);
}

}  // namespace
}  // namespace kg

using namespace kg;

extern "C" {

static int edgeset_finish(kombgpu_ctx *ctx, int rc, kombgpu_edgeset *es, DevBuf<uint64_t> &edges, uint64_t E,
                          kombgpu_edgeset **out) {
    if (rc != KOMBGPU_OK) { delete es; return rc; }
    es->n_edges = E;
    es->edges = edges.take();
    es->st.n_edges = E;
    *out = es;
    return KOMBGPU_OK;
}

int kombgpu_local_edges_dev(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits,
                            uint32_t n_global, kombgpu_edgeset **out) {
    if (!ctx) return KOMBGPU_EINVAL;
    if (!out || (n_hits && (!read_key || !unitig))) return ctx_fail(ctx, KOMBGPU_EINVAL, "null argument");
    *out = nullptr;
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    kombgpu_edgeset *es = new (std::nothrow) kombgpu_edgeset();
    if (!es) return ctx_fail(ctx, KOMBGPU_ENOMEM, "host allocation");
    es->ctx = ctx;
    es->n_global = n_global;
    DevBuf<uint64_t> edges;
    uint64_t E = 0;
    int rc = hits_to_edges(ctx, read_key, unitig, n_hits, n_global, edges, &E, &es->st);
    return edgeset_finish(ctx, rc, es, edges, E, out);
}

int kombgpu_edgeset_from_pairs_dev(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_global,
                                   kombgpu_edgeset **out) {
    if (!ctx) return KOMBGPU_EINVAL;
    if (!out || (n_pairs && (!u || !v))) return ctx_fail(ctx, KOMBGPU_EINVAL, "null argument");
    *out = nullptr;
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    kombgpu_edgeset *es = new (std::nothrow) kombgpu_edgeset();
    if (!es) return ctx_fail(ctx, KOMBGPU_ENOMEM, "host allocation");
    es->ctx = ctx;
    es->n_global = n_global;
    es->st.n_pairs = n_pairs;
    DevBuf<uint64_t> edges;
    uint64_t E = 0;
    int rc = pairs_to_edges(ctx, u, v, n_pairs, n_global, edges, &E);
    return edgeset_finish(ctx, rc, es, edges, E, out);
}

int kombgpu_edgeset_counts(const kombgpu_edgeset *es, uint64_t *n_edges, uint64_t *n_pairs, uint64_t *n_unique_hits) {
    if (!es) return KOMBGPU_EINVAL;
    if (n_edges) *n_edges = es->n_edges;
    if (n_pairs) *n_pairs = es->st.n_pairs;
    if (n_unique_hits) *n_unique_hits = es->st.n_unique_hits;
    return KOMBGPU_OK;
}

int kombgpu_edgeset_route_dev(kombgpu_edgeset *es, const uint32_t *bounds, int n_parts, uint64_t *send, uint64_t *counts) {
    if (!es) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = es->ctx;
    Bounds bd;
    KG_TRY(make_bounds(ctx, bounds, n_parts, &bd));
    if (!counts || (es->n_edges && !send)) return ctx_fail(ctx, KOMBGPU_EINVAL, "null argument");
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t E = es->n_edges;
    // forward entries (u << 32 | v) are the edge list itself; backward entries are its swapped copy sorted by v
    DevBuf<uint64_t> sa, sb;
    uint64_t *swapped = nullptr;
    KG_TRY(swapped_sorted(ctx, es->edges, E, es->n_global, sa, sb, &swapped));
    uint64_t fwd[kMaxParts + 1], bwd[kMaxParts + 1];
    KG_TRY(lower_bounds_hi(ctx, es->edges, E, bd.b, n_parts + 1, fwd));
    KG_TRY(lower_bounds_hi(ctx, swapped, E, bd.b, n_parts + 1, bwd));
    if (fwd[0] != 0 || bwd[0] != 0 || fwd[n_parts] != E || bwd[n_parts] != E)
        return ctx_fail(ctx, KOMBGPU_EINVAL, "partition bounds do not cover every unitig id of the edge set");
    uint64_t off = 0;
    for (int j = 0; j < n_parts; ++j) {
        const uint64_t nf = fwd[j + 1] - fwd[j], nb = bwd[j + 1] - bwd[j];
        if (nf) KG_CUDA(ctx, cudaMemcpyAsync(send + off, es->edges + fwd[j], nf * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        if (nb) KG_CUDA(ctx, cudaMemcpyAsync(send + off + nf, swapped + bwd[j], nb * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        counts[j] = nf + nb;
        off += nf + nb;
    }
    KG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KOMBGPU_OK;
}

void kombgpu_edgeset_destroy(kombgpu_edgeset *es) {
    if (!es) return;
    if (es->edges) ws_free(es->ctx, es->edges);
    delete es;
}

int kombgpu_part_build_dev(kombgpu_ctx *ctx, const uint64_t *entries, uint64_t count, uint32_t v_lo, uint32_t v_hi,
                           uint32_t n_global, kombgpu_part **out) {
    if (!ctx) return KOMBGPU_EINVAL;
    if (!out || (count && !entries) || v_hi < v_lo || v_hi > n_global) return ctx_fail(ctx, KOMBGPU_EINVAL, "bad argument");
    *out = nullptr;
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    kombgpu_part *p = new (std::nothrow) kombgpu_part();
    if (!p) return ctx_fail(ctx, KOMBGPU_ENOMEM, "host allocation");
    p->ctx = ctx;
    p->v_lo = v_lo;
    p->n_local = v_hi - v_lo;
    p->n_global = n_global;
    int rc = csr_from_directed(ctx, entries, count, v_lo, p->n_local, n_global, &p->row_ptr, &p->col, &p->deg, &p->max_degree,
                               &p->n_directed);
    if (rc != KOMBGPU_OK) { part_release(p); delete p; return rc; }
    *out = p;
    return KOMBGPU_OK;
}

void kombgpu_part_destroy(kombgpu_part *p) {
    if (!p) return;
    part_release(p);
    delete p;
}

int kombgpu_part_counts(const kombgpu_part *p, uint32_t *n_local, uint64_t *n_directed, int32_t *max_degree) {
    if (!p) return KOMBGPU_EINVAL;
    if (n_local) *n_local = p->n_local;
    if (n_directed) *n_directed = p->n_directed;
    if (max_degree) *max_degree = p->max_degree;
    return KOMBGPU_OK;
}

int kombgpu_part_device_arrays(const kombgpu_part *p, const uint64_t **row_ptr, const uint32_t **col, const int32_t **degree,
                               const int32_t **coreness) {
    if (!p) return KOMBGPU_EINVAL;
    if (row_ptr) *row_ptr = p->row_ptr;
    if (col) *col = p->col;
    if (degree) *degree = p->deg;
    if (coreness) *coreness = p->core;
    return KOMBGPU_OK;
}

int kombgpu_part_peel_begin(kombgpu_part *p) {
    if (!p) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = p->ctx;
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = p->n_local ? p->n_local : 1;
    auto need = [&](void **slot, size_t bytes) -> bool {
        if (!*slot) *slot = ws_alloc(ctx, bytes);
        return *slot != nullptr;
    };
    int per_sm = 0;
    KG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, part_process_kernel, kPeelThreads, 0));
    if (per_sm < 1) return ctx_fail(ctx, KOMBGPU_ECUDA, "partition peel kernel does not fit on an SM");
    p->grid = per_sm * ctx->sm_count;
    // every local vertex enters the pool at most once, every long row is sliced at most once, plus reserved slots
    const uint64_t cap64 = (uint64_t)p->n_local + p->n_directed / kSliceLen + p->n_directed / kSplit + (uint64_t)p->grid * kClaimMax + 64;
    if (cap64 >= 0xffffffffull || p->n_directed >= (1ull << (63 - kSliceLenBits)))
        return ctx_fail(ctx, KOMBGPU_EINVAL, "partition too large for the pool encoding");
    p->pool_cap = (uint32_t)cap64;
    bool ok = need((void **)&p->core, n * sizeof(int32_t)) && need((void **)&p->pool, (size_t)p->pool_cap * sizeof(uint64_t)) &&
              need((void **)&p->alive[0], n * sizeof(uint32_t)) && need((void **)&p->alive[1], n * sizeof(uint32_t)) &&
              need((void **)&p->outbox, (p->n_directed ? p->n_directed : 1) * sizeof(uint32_t)) &&
              need((void **)&p->state, sizeof(PeelState)) && need((void **)&p->counters, 4 * sizeof(uint32_t));
    if (!ok) return ctx_fail(ctx, KOMBGPU_ENOMEM, "partition peel state");
    KG_CUDA(ctx, cudaMemcpyAsync(p->core, p->deg, p->n_local * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KG_CUDA(ctx, cudaMemsetAsync(p->pool, 0xff, (size_t)p->pool_cap * sizeof(uint64_t), ctx->stream));
    KG_CUDA(ctx, cudaMemsetAsync(p->state, 0, sizeof(PeelState), ctx->stream));
    KG_CUDA(ctx, cudaMemsetAsync(p->counters, 0, 4 * sizeof(uint32_t), ctx->stream));
    p->n_alive = p->n_local;
    p->alive_cur = -1;
    p->n_front = 0;
    p->n_outbox = 0;
    p->round = 0;
    return KOMBGPU_OK;
}

int kombgpu_part_peel_scan(kombgpu_part *p, int32_t k, uint32_t *n_front, uint32_t *n_alive, int32_t *min_next) {
    if (!p || !p->state) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = p->ctx;
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    // reset the per-scan slot, keep the pool counters and the run-long statistics
    PeelState cur{};
    KG_TRY(read_back(ctx, p->state, &cur, 1));
    cur.alive_out[0] = 0;
    cur.front_cnt[0] = 0;
    cur.next_min[0] = INT32_MAX;
    KG_CUDA(ctx, cudaMemcpyAsync(p->state, &cur, sizeof(cur), cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t *src = p->alive_cur < 0 ? nullptr : p->alive[p->alive_cur];
    const int dst_i = p->alive_cur < 0 ? 0 : (p->alive_cur ^ 1);
    if (p->n_alive) {
        uint32_t grid = min(ceil_div_u64(p->n_alive, kPeelThreads), (uint32_t)p->grid);  // scan_alive spreads short lists over the CTAs
        KG_LAUNCH(ctx, part_scan_kernel, grid, kPeelThreads, 0, k, src, p->n_alive, p->alive[dst_i], p->core, p->pool, p->state);
    }
    PeelState res{};
    KG_TRY(read_back(ctx, p->state, &res, 1));
    p->alive_cur = dst_i;
    p->n_alive = res.alive_out[0];
    p->n_front = res.front_cnt[0];
    if (n_front) *n_front = p->n_front;
    if (n_alive) *n_alive = p->n_alive;
    if (min_next) *min_next = res.next_min[0];
    return KOMBGPU_OK;
}

int kombgpu_part_peel_process(kombgpu_part *p, int32_t k, uint32_t *n_outbox) {
    if (!p || !p->state) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = p->ctx;
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    p->n_outbox = 0;
    if (p->n_front) {
        KG_CUDA(ctx, cudaMemsetAsync(p->counters, 0, 4 * sizeof(uint32_t), ctx->stream));  // outbox empty
        PartView pv;
        pv.v_lo = p->v_lo;
        pv.n_local = p->n_local;
        pv.outbox = p->outbox;
        pv.outbox_cnt = &p->counters[0];
        int32_t k_arg = k;
        uint32_t round = ++p->round;
        const uint64_t *row_ptr = p->row_ptr;
        const uint32_t *col = p->col;
        int32_t *deg = p->core;
        uint64_t *Q = p->pool;
        uint32_t cap = p->pool_cap;
        PeelState *st = p->state;
        void *args[] = {&k_arg, &round, &row_ptr, &col, &deg, &Q, &cap, &st, &pv};
        KG_CUDA(ctx, cudaLaunchCooperativeKernel((void *)part_process_kernel, dim3(p->grid), dim3(kPeelThreads), args, 0, ctx->stream));
        ctx->launches++;
        uint32_t c[4];
        KG_TRY(read_back(ctx, p->counters, c, 4));
        p->n_outbox = c[0];
        p->n_front = 0;
        uint32_t err = 0;
        KG_TRY(read_back(ctx, &p->state->error, &err, 1));
        if (err) return ctx_fail(ctx, KOMBGPU_EINTERNAL, "partition peel invariant broken (code %u)", err);
    }
    if (n_outbox) *n_outbox = p->n_outbox;
    return KOMBGPU_OK;
}

int kombgpu_part_outbox_route_dev(kombgpu_part *p, const uint32_t *bounds, int n_parts, uint32_t *send, uint64_t *counts) {
    if (!p || !p->state) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = p->ctx;
    Bounds bd;
    KG_TRY(make_bounds(ctx, bounds, n_parts, &bd));
    if (!counts || (p->n_outbox && !send)) return ctx_fail(ctx, KOMBGPU_EINVAL, "null argument");
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int j = 0; j < n_parts; ++j) counts[j] = 0;
    if (p->n_outbox == 0) return KOMBGPU_OK;
    DevBuf<uint32_t> d_counts, d_cursors;
    KG_ALLOC(ctx, d_counts, kMaxParts);
    KG_ALLOC(ctx, d_cursors, kMaxParts);
    KG_CUDA(ctx, cudaMemsetAsync(d_counts.p, 0, kMaxParts * sizeof(uint32_t), ctx->stream));
    const uint32_t grid = min(ceil_div_u64(p->n_outbox, kThreads * 8), (uint32_t)ctx->sm_count * 8u);
    KG_LAUNCH(ctx, outbox_count_kernel, grid, kThreads, 0, p->outbox, p->n_outbox, bd, d_counts.p);
    uint32_t h_counts[kMaxParts];
    KG_TRY(read_back(ctx, d_counts.p, h_counts, kMaxParts));
    uint32_t h_cursors[kMaxParts];
    uint32_t run = 0;
    for (int j = 0; j < kMaxParts; ++j) {
        h_cursors[j] = run;
        if (j < n_parts) { counts[j] = h_counts[j]; run += h_counts[j]; }
    }
    KG_CUDA(ctx, cudaMemcpyAsync(d_cursors.p, h_cursors, kMaxParts * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    KG_LAUNCH(ctx, outbox_scatter_kernel, grid, kThreads, 0, p->outbox, p->n_outbox, bd, d_cursors.p, send);
    KG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KOMBGPU_OK;
}

int kombgpu_part_peel_apply_dev(kombgpu_part *p, int32_t k, const uint32_t *recv, uint64_t count, uint32_t *n_front) {
    if (!p || !p->state) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = p->ctx;
    if (count && !recv) return ctx_fail(ctx, KOMBGPU_EINVAL, "null argument");
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count) {
        KG_CUDA(ctx, cudaMemsetAsync(&p->counters[1], 0, sizeof(uint32_t), ctx->stream));
        const uint32_t grid = min(ceil_div_u64(count, kThreads), (uint32_t)ctx->sm_count * 8u);
        KG_LAUNCH(ctx, part_apply_kernel, grid, kThreads, 0, k, recv, count, p->v_lo, p->n_local, p->core, p->pool, p->pool_cap,
                  p->state, &p->counters[1]);
        uint32_t c[4];
        KG_TRY(read_back(ctx, p->counters, c, 4));
        p->n_front = c[1];
    } else {
        p->n_front = 0;
    }
    if (n_front) *n_front = p->n_front;
    return KOMBGPU_OK;
}

int kombgpu_corea_dev(kombgpu_ctx *ctx, const int32_t *core, const int32_t *deg, uint32_t n, int key_mode, double *score,
                      double *max_score) {
    if (!ctx) return KOMBGPU_EINVAL;
    if (n && (!core || !deg || !score)) return ctx_fail(ctx, KOMBGPU_EINVAL, "null argument");
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    double mx = 0.0;
    KG_TRY(corea_scores(ctx, core, deg, n, key_mode, score, &mx));
    KG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (max_score) *max_score = mx;
    return KOMBGPU_OK;
}

}  // extern "C"
