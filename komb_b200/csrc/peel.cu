// peel.cu — stage 2: k-core decomposition by frontier peeling.
//
// Replaces igraph_coreness (src/graph.cpp:463; Batagelj-Zaversnik serial bucket
// peel) and, as the same machinery, the serial min-heap peel the reference keeps
// in HashIndexedMinHeap.h.  Coreness is a unique function of the simple graph,
// so the result is bit-exact against any correct implementation.
//
// One persistent cooperative kernel peels the whole graph:
//   level k:  SCAN     one pass over the compacted alive list: vertices with
//                      deg == k are appended to the task pool, vertices above k
//                      are compacted (CTA-wide scan, three global atomics per tile)
//             PROCESS  CTAs claim pool entries (fetch-add tickets) and walk the
//                      rows edge-parallel: for each neighbour u with deg[u] > k:
//                      atomicSub(deg[u]); the thread that takes it to k owns u and
//                      pushes it (ballot/popc aggregated) to the CTA's own
//                      shared-memory list, processed next by the same CTA, so a
//                      cascade chain needs no grid barrier and no global queue; a
//                      decrement that lands below k is undone, so deg[] is clamped
//                      at k and the final deg[] IS the coreness.  A CTA that
//                      discovers more than it can use shares the surplus through
//                      the pool; hub rows go to the pool as slices; idle CTAs park
//                      on reserved pool slots and are handed work with one store.
// Empty levels are skipped through a min-reduction of the survivors' degrees
// done by the scan itself.  The peel is bound by its dependency depth (levels x
// cascade generations), not by bytes: see DESIGN.md "Peel" and peel_device.cuh.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "peel_device.cuh"
#include "peel_warp.cuh"

namespace cg = cooperative_groups;

namespace kg {

namespace {

using namespace peel;

constexpr unsigned long long kTraceWords = 20;   // words per round of the optional trace

union PeelShared {
    BlockShared cta;   // scan phase (both modes) and the CTA-wide process phase
    WarpShared warp;   // warp-autonomous process phase (the phases are separated by grid barriers)
};

template <bool kWarpMode, int kU>
__global__ void __launch_bounds__(kPeelThreads) peel_kernel(uint32_t n, const uint64_t *__restrict__ row_ptr,
                                                            const uint32_t *__restrict__ col, int32_t *deg, int32_t *core_out,
                                                            uint64_t *Q, uint32_t cap, uint32_t *alive_a, uint32_t *alive_b,
                                                            PeelState *st) {
    cg::grid_group grid = cg::this_grid();
    __shared__ PeelShared shu;
    BlockShared &sh = shu.cta;
    const uint32_t tid = threadIdx.x, lane = lane_id();

    int32_t k = 0;
    uint32_t n_alive = n;
    const uint32_t *alive_src = nullptr;  // nullptr = identity (every vertex)
    uint32_t *alive_dst = alive_a;
    uint32_t round = 0;
    uint32_t removed = 0;                 // vertices this CTA peeled
    uint32_t pool_base = 0;               // pool position at the end of the last level

    while (true) {
        const uint32_t par = round % 3;
        // ---------------- SCAN ----------------
        if (blockIdx.x == 0 && tid == 0) {  // re-arm the slot of the NEXT round
            const uint32_t nxt = (round + 1) % 3;
            st->alive_out[nxt] = 0;
            st->front_cnt[nxt] = 0;
            st->next_min[nxt] = INT32_MAX;
            st->rounds = round + 1;
        }
        const bool prof = (blockIdx.x == 0 && tid == 0);
        unsigned long long tp0 = prof ? global_ns() : 0;
        long long stamps[6] = {0, 0, 0, 0, 0, 0};
        int32_t local_min = scan_alive(k, alive_src, n_alive, alive_dst, deg, Q, &st->q_tail, &st->front_cnt[par],
                                       &st->alive_out[par], &st->n_isolated, sh, (prof && st->trace) ? stamps : nullptr);
        local_min = warp_reduce_min(local_min);
        if (lane == 0 && local_min != INT32_MAX) atomicMin(&st->next_min[par], local_min);
        if (prof) stamps[5] = clock64();
        unsigned long long tp1 = prof ? global_ns() : 0;
        grid.sync();
        unsigned long long tp2 = prof ? global_ns() : 0;
        if (prof) { st->prof_ns[0] += tp1 - tp0; st->prof_ns[1] += tp2 - tp1; }

        const uint32_t front_cnt = __ldcg(&st->front_cnt[par]);
        const uint32_t survivors = __ldcg(&st->alive_out[par]);
        const int32_t min_next = __ldcg(&st->next_min[par]);
        // the compacted list becomes the next scan's input
        alive_src = alive_dst;
        alive_dst = (alive_dst == alive_a) ? alive_b : alive_a;
        n_alive = survivors;
        ++round;
        const uint32_t trace_row = round - 1;
        if (prof && st->trace && trace_row < st->trace_cap) {
            unsigned long long *tr = st->trace + kTraceWords * trace_row;
            tr[0] = (unsigned long long)(uint32_t)k; tr[1] = front_cnt; tr[2] = survivors; tr[3] = tp1 - tp0; tr[4] = 0; tr[5] = 0;
            // SM cycles inside CTA 0's scan: id loads, degree gather, CTA scan + counters, stores + tail
            tr[6] = stamps[1] - stamps[0]; tr[7] = stamps[2] - stamps[1]; tr[8] = stamps[3] - stamps[2]; tr[9] = stamps[5] - stamps[3];
        }
        if (front_cnt == 0) {
            // empty level: nothing to process; jump to the smallest remaining degree
            if (survivors == 0 || min_next == INT32_MAX) break;
            k = min_next;
            continue;
        }
        if (blockIdx.x == 0 && tid == 0) {
            st->levels += 1;
            st->max_core = k;
        }

        // ---------------- PROCESS ----------------
        if (kWarpMode) {
            // the scan appended the frontier at the pool position the last level ended at
            removed += process_level_warp<false, kU>(k, round, Q, cap, row_ptr, col, deg, core_out, st, shu.warp, PartView{}, pool_base, front_cnt);
        }
        else removed += process_level<false>(k, round, Q, cap, row_ptr, col, deg, st, sh, PartView{});
        unsigned long long tp3 = prof ? global_ns() : 0;
        grid.sync();
        if (prof) {
            const unsigned long long tp4 = global_ns();
            st->prof_ns[2] += tp3 - tp2; st->prof_ns[3] += tp4 - tp3;
            if (st->trace && trace_row < st->trace_cap) {
                unsigned long long *tr = st->trace + kTraceWords * trace_row;
                tr[4] = tp3 - tp2; tr[5] = tp4 - tp3;
                // cumulative counters over all CTAs (complete: every CTA added its share before the barrier)
                tr[15] = *(volatile unsigned long long *)&st->batches; tr[16] = *(volatile unsigned long long *)&st->carried;
                tr[17] = *(volatile unsigned long long *)&st->ring_pushed; tr[18] = *(volatile unsigned long long *)&st->shared;
            }
        }
        // q_done == q_tail here and nothing moves until the next PROCESS phase
        if (__ldcg(&st->error)) break;
        pool_base = __ldcg(&st->q_done);  // == q_tail, and stable until the next PROCESS phase starts
        if (blockIdx.x == 0 && tid == 0) st->q_head = pool_base;  // un-reserve the slots past the tail
        k += 1;
    }
    if (tid == 0 && removed) atomicAdd(&st->n_removed, (unsigned long long)removed);
}

}  // namespace

int peel_coreness(kombgpu_graph *g) {
    kombgpu_ctx *ctx = g->ctx;
    const uint32_t n = g->n;
    g->st.max_coreness = 0;
    g->st.peel_levels = 0;
    g->st.peel_rounds = 0;
    if (!g->core) {
        g->core = static_cast<int32_t *>(ws_alloc(ctx, (n ? n : 1) * sizeof(int32_t)));
        if (!g->core) return ctx_fail(ctx, KOMBGPU_ENOMEM, "coreness array");
    }
    if (n == 0) { g->has_core = true; return KOMBGPU_OK; }

    const char *mode_env = getenv("KOMBGPU_PEEL_MODE");   // "cta": the CTA-wide process phase (kept for A/B measurements)
    const bool warp_mode = !(mode_env && mode_env[0] == 'c');
    const char *unroll_env = getenv("KOMBGPU_PEEL_UNROLL");
    const bool unroll8 = unroll_env && atoi(unroll_env) == 8;
    void *kernel = !warp_mode ? (void *)peel_kernel<false, 4> : unroll8 ? (void *)peel_kernel<true, 8> : (void *)peel_kernel<true, 4>;
    int per_sm = 0;
    if (!warp_mode) KG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peel_kernel<false, 4>, kPeelThreads, 0));
    else if (unroll8) KG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peel_kernel<true, 8>, kPeelThreads, 0));
    else KG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peel_kernel<true, 4>, kPeelThreads, 0));
    if (per_sm < 1) return ctx_fail(ctx, KOMBGPU_ECUDA, "peel kernel does not fit on an SM");
    const int grid = per_sm * ctx->sm_count;  // persistent: every CTA resident (cooperative launch)

    // warp mode: working degrees in a scratch array (never clamped), coreness written to g->core as vertices are
    // taken (degree-0 vertices are peeled by the scan alone: their 0 comes from the memset);
    // cta mode: g->core holds the working degrees, clamped at the current level, and ends as the coreness
    DevBuf<int32_t> work;
    int32_t *deg_work = g->core;
    if (warp_mode) {
        KG_ALLOC(ctx, work, n);
        deg_work = work.p;
        KG_CUDA(ctx, cudaMemsetAsync(g->core, 0, (size_t)n * sizeof(int32_t), ctx->stream));
    }
    KG_CUDA(ctx, cudaMemcpyAsync(deg_work, g->deg, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    DevBuf<uint32_t> alive_a, alive_b;
    DevBuf<uint64_t> pool;
    DevBuf<PeelState> state(ctx, 1);
    // every vertex enters the pool at most once; a row is sliced at most once (<= 2E/kSplit long rows, each giving
    // <= len/kSliceLen + 1 slices); plus the slots idle CTAs reserve past the tail
    // (warp mode re-queues the pieces of rows longer than kWarpSplit, through the pool when the CTA's ring is busy)
    const uint64_t cap64 = (uint64_t)n + 2 * g->n_edges / kWarpSplitMin + 2 * g->n_edges / kSplit + 4 * g->n_edges / kWarpSplitMin +
                           (uint64_t)grid * kClaimMax + 64;
    if (cap64 >= 0xffffffffull || 2 * g->n_edges >= (1ull << (62 - kSliceLenBits)))
        return ctx_fail(ctx, KOMBGPU_EINVAL, "graph too large for the pool encoding");
    const uint32_t cap = (uint32_t)cap64;
    KG_ALLOC(ctx, pool, cap);
    KG_CUDA(ctx, cudaMemsetAsync(pool.p, 0xff, (size_t)cap * sizeof(uint64_t), ctx->stream));
    KG_ALLOC(ctx, alive_a, n);
    KG_ALLOC(ctx, alive_b, n);
    if (!state) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    PeelState init{};
    for (int i = 0; i < 3; ++i) init.next_min[i] = INT32_MAX;
    init.tune[0] = (uint32_t)kRingKeep;
    init.tune[1] = kWarpSplit;
    init.tune[2] = 0;
    init.tune[3] = kThinEdges;
    init.tune[4] = kSliceLen;
    if (const char *e = getenv("KOMBGPU_PEEL_KEEP")) init.tune[0] = (uint32_t)atoi(e);
    if (const char *e = getenv("KOMBGPU_PEEL_WSPLIT")) init.tune[1] = (uint32_t)atoi(e) < kWarpSplitMin ? kWarpSplitMin : (uint32_t)atoi(e);
    if (const char *e = getenv("KOMBGPU_PEEL_PARK")) init.tune[2] = (uint32_t)atoi(e);
    if (const char *e = getenv("KOMBGPU_PEEL_THIN")) init.tune[3] = (uint32_t)atoi(e);
    if (const char *e = getenv("KOMBGPU_PEEL_HUBSLICE")) init.tune[4] = (uint32_t)atoi(e) < kWarpSplitMin ? kWarpSplitMin : (uint32_t)atoi(e);
    DevBuf<unsigned long long> trace;
    const char *trace_path = getenv("KOMBGPU_TRACE");
    const uint32_t trace_cap = 1u << 16;
    if (trace_path) {
        KG_ALLOC(ctx, trace, kTraceWords * trace_cap);
        KG_CUDA(ctx, cudaMemsetAsync(trace.p, 0, kTraceWords * trace_cap * sizeof(unsigned long long), ctx->stream));
        init.trace = trace.p;
        init.trace_cap = trace_cap;
    }
    KG_CUDA(ctx, cudaMemcpyAsync(state.p, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));

    uint32_t n_arg = n;
    const uint64_t *row_ptr = g->row_ptr;
    const uint32_t *col = g->col;
    int32_t *core = deg_work;
    int32_t *core_out = g->core;
    uint32_t *aa = alive_a.p, *ab = alive_b.p;
    uint64_t *q = pool.p;
    uint32_t cap_arg = cap;
    PeelState *sp = state.p;
    void *args[] = {&n_arg, &row_ptr, &col, &core, &core_out, &q, &cap_arg, &aa, &ab, &sp};
    cudaEvent_t ev0, ev1;
    KG_CUDA(ctx, cudaEventCreate(&ev0));
    KG_CUDA(ctx, cudaEventCreate(&ev1));
    KG_CUDA(ctx, cudaEventRecord(ev0, ctx->stream));
    cudaError_t le = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kPeelThreads), args, 0, ctx->stream);
    cudaEventRecord(ev1, ctx->stream);
    ctx->launches++;
    if (le == cudaSuccess) le = cudaEventSynchronize(ev1);
    if (le == cudaSuccess) cudaEventElapsedTime(&g->st.ms_peel_kernel, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (le != cudaSuccess) return ctx_fail(ctx, KOMBGPU_ECUDA, "peel kernel: %s", cudaGetErrorString(le));

    PeelState fin{};
    KG_TRY(read_back(ctx, state.p, &fin, 1));
    if (getenv("KOMBGPU_DEBUG"))
        fprintf(stderr, "[kombgpu] peel n=%u levels=%u rounds=%u grid=%d kernel=%.3f ms | cta0: scan %.3f ms, sync1 %.3f ms, process %.3f ms, sync2 %.3f ms | batches %llu shared %llu slices %llu\n",
                n, fin.levels, fin.rounds, grid, g->st.ms_peel_kernel, fin.prof_ns[0] * 1e-6, fin.prof_ns[1] * 1e-6,
                fin.prof_ns[2] * 1e-6, fin.prof_ns[3] * 1e-6, fin.batches, fin.shared, fin.sliced);
    if (trace_path) {
        const uint32_t rows = fin.rounds < trace_cap ? fin.rounds : trace_cap;
        std::vector<unsigned long long> h(kTraceWords * rows);
        KG_CUDA(ctx, cudaMemcpy(h.data(), trace.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(trace_path, "w")) {
            fprintf(f, "round,k,frontier,survivors,scan_ns,process_ns,wait_ns,scan_cyc_ids,scan_cyc_deg,scan_cyc_counters,scan_cyc_tail,p_init,p_first,p_lastdone,p_over,p_exit,cum_batches,cum_carried,cum_ring,cum_pool\n");
            for (uint32_t r = 0; r < rows; ++r) {
                const unsigned long long *t = &h[kTraceWords * r];
                fprintf(f, "%u,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu\n", r, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15], t[16], t[17], t[18]);
            }
            fclose(f);
        }
    }
    if (fin.error) return ctx_fail(ctx, KOMBGPU_EINTERNAL, "peel invariant broken (code %u, %llu of %u vertices peeled)", fin.error, fin.n_removed, n);
    if (fin.n_removed + fin.n_isolated != n)
        return ctx_fail(ctx, KOMBGPU_EINTERNAL, "peel ended with %llu of %u vertices peeled", fin.n_removed + fin.n_isolated, n);
    g->st.max_coreness = fin.max_core;
    g->st.peel_levels = fin.levels + (fin.n_isolated ? 1u : 0u);  // level 0 is handled by the scan alone
    g->st.peel_rounds = fin.rounds;
    g->has_core = true;
    return KOMBGPU_OK;
}

}  // namespace kg
