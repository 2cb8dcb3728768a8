// peel.cu — stage 2: k-core decomposition by frontier peeling.
//
// Replaces igraph_coreness (src/graph.cpp:463; Batagelj-Zaversnik serial bucket
// peel) and, as the same machinery, the serial min-heap peel the reference keeps
// in HashIndexedMinHeap.h.  Coreness is a unique function of the simple graph,
// so the result is bit-exact against any correct implementation.
//
// One persistent cooperative kernel peels the whole graph:
//   level k:  SCAN    compact the alive list; vertices with deg == k enter the
//                     peel queue (ballot/popc warp-aggregated appends)
//             PROCESS every queued vertex v: for each neighbour u with
//                     deg[u] > k: atomicSub(deg[u]); the thread that takes it to
//                     k appends u to the SAME queue (so the level-k cascade
//                     needs no further grid-wide barrier); a decrement that
//                     lands below k is undone, so deg[] is clamped at k and the
//                     final deg[] IS the coreness.
// The queue is the peel order: each vertex is appended exactly once over the
// whole run, so it is never reset; CTAs claim chunks from it with a CAS and a
// level ends when q_done == q_tail.  Empty levels are skipped through a
// min-reduction of the survivors' degrees done by the scan itself.
#include <cooperative_groups.h>

#include "graph.cuh"

namespace cg = cooperative_groups;

namespace kg {

namespace {

constexpr int kPeelThreads = 512;
constexpr int kPeelWarps = kPeelThreads / 32;
constexpr uint32_t kSentinel = 0xffffffffu;
constexpr uint32_t kMaxChunk = 4 * kPeelWarps;               // queue entries one CTA claims at once
constexpr unsigned long long kWatchdogNs = 10ull * 1000000000ull;  // a wait this long means a broken invariant

// Per-round scan results live in three rotating slots: round r uses slot r % 3
// and CTA 0 re-arms slot (r + 1) % 3 at the start of round r.  That slot was last
// read right after the grid barrier of round r - 2, and CTA 0 can only be in
// round r once every CTA has arrived at the barrier of round r - 1, so nobody
// can still be reading it.  All control-flow decisions are taken from these
// slots (or from q_done / error at points where they cannot change), never from
// q_tail, which other CTAs may already be advancing.
struct PeelState {
    uint32_t q_tail;        // queue entries appended (monotone, ends at n)
    uint32_t q_head;        // queue entries claimed
    uint32_t q_done;        // queue entries fully processed
    uint32_t alive_out[3];  // survivors written by the scan of round r (slot r % 3)
    uint32_t front_cnt[3];  // vertices that scan put on the queue
    int32_t next_min[3];    // min degree of the survivors
    uint32_t error;         // watchdog / invariant flag
    uint32_t levels;        // non-empty levels
    uint32_t rounds;        // scan phases executed
    int32_t max_core;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Claim protocol.  q_head only moves by atomicAdd, so claims never fail or retry
// (a CAS loop admits one winner per L2 round trip and serialises the whole
// level).  A claim made while q_head < q_tail can still overshoot q_tail; the
// claimer then OWNS slots that are not written yet and waits for them.  A slot
// is abandoned only when the level is quiescent (q_done == q_tail): every
// appended entry is processed, nobody can append any more, so the slot cannot
// fill during this level.  CTA 0 pulls q_head back to q_tail before the next
// level, which makes abandoned slots claimable again.
__global__ void __launch_bounds__(kPeelThreads) peel_kernel(uint32_t n, const uint64_t *__restrict__ row_ptr,
                                                            const uint32_t *__restrict__ col, int32_t *deg,
                                                            uint32_t *queue, uint32_t *alive_a, uint32_t *alive_b,
                                                            PeelState *st) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_begin, s_cnt;
    __shared__ uint64_t s_row[kMaxChunk];      // first edge of each claimed vertex
    __shared__ uint32_t s_off[kMaxChunk + 1];  // exclusive prefix of their row lengths
    __shared__ uint32_t s_scan[kPeelWarps + 1];
    const uint32_t lane = lane_id();
    const uint32_t warp_in_block = threadIdx.x >> 5;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kPeelWarps + warp_in_block;
    const uint64_t warps_total = (uint64_t)gridDim.x * kPeelWarps;

    int32_t k = 0;
    uint32_t n_alive = n;
    const uint32_t *alive_src = nullptr;  // nullptr = identity (every vertex)
    uint32_t *alive_dst = alive_a;
    uint32_t round = 0;

    while (true) {
        const uint32_t par = round % 3;
        // ---------------- SCAN ----------------
        if (blockIdx.x == 0 && threadIdx.x == 0) {  // re-arm the slot of the NEXT round
            const uint32_t nxt = (round + 1) % 3;
            st->alive_out[nxt] = 0;
            st->front_cnt[nxt] = 0;
            st->next_min[nxt] = INT32_MAX;
            st->rounds = round + 1;
        }
        int32_t local_min = INT32_MAX;
        for (uint64_t base = warp_global * 32; base < n_alive; base += warps_total * 32) {
            const uint64_t i = base + lane;
            bool front = false, surv = false;
            uint32_t v = 0;
            if (i < n_alive) {
                v = alive_src ? alive_src[i] : (uint32_t)i;
                const int32_t d = __ldcg(&deg[v]);
                front = (d == k);
                surv = (d > k);  // d < k: peeled at an earlier level, drop from the list
                if (surv) local_min = min(local_min, d);
            }
            const uint32_t fm = __ballot_sync(kFullMask, front);
            const uint32_t sm = __ballot_sync(kFullMask, surv);
            uint32_t fbase = 0, sbase = 0;
            if (lane == 0) {
                if (fm) {
                    fbase = atomicAdd(&st->q_tail, (uint32_t)__popc(fm));
                    atomicAdd(&st->front_cnt[par], (uint32_t)__popc(fm));
                }
                if (sm) sbase = atomicAdd(&st->alive_out[par], (uint32_t)__popc(sm));
            }
            fbase = __shfl_sync(kFullMask, fbase, 0);
            sbase = __shfl_sync(kFullMask, sbase, 0);
            if (front) queue[fbase + __popc(fm & lanemask_lt())] = v;
            if (surv) alive_dst[sbase + __popc(sm & lanemask_lt())] = v;
        }
        local_min = warp_reduce_min(local_min);
        if (lane == 0 && local_min != INT32_MAX) atomicMin(&st->next_min[par], local_min);
        grid.sync();

        const uint32_t front_cnt = __ldcg(&st->front_cnt[par]);
        const uint32_t survivors = __ldcg(&st->alive_out[par]);
        const int32_t min_next = __ldcg(&st->next_min[par]);
        // the compacted list becomes the next scan's input
        alive_src = alive_dst;
        alive_dst = (alive_dst == alive_a) ? alive_b : alive_a;
        n_alive = survivors;
        ++round;

        if (front_cnt == 0) {
            // empty level: nothing to process; jump to the smallest remaining degree
            if (survivors == 0 || min_next == INT32_MAX) break;
            k = min_next;
            continue;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->levels += 1;
            st->max_core = k;
        }

        // ---------------- PROCESS ----------------
        // Thread i of the CTA owns queue slot s_begin + i of the current claim and keeps
        // it `pending` until that slot has been written AND processed.
        bool pending = false;
        uint32_t my_slot = 0;
        unsigned long long idle_since = 0;
        while (true) {
            if (__syncthreads_count(pending) == 0) {
                if (threadIdx.x == 0) {
                    uint32_t begin = 0, cnt = 0;
                    while (true) {
                        const uint32_t h = ld_volatile_u32(&st->q_head);
                        const uint32_t t = ld_volatile_u32(&st->q_tail);
                        if (h < t) {
                            uint32_t take = (t - h + gridDim.x - 1) / gridDim.x;
                            take = min(max(take, 1u), kMaxChunk);
                            begin = atomicAdd(&st->q_head, take);
                            cnt = begin < n ? min(take, n - begin) : 0;
                            if (cnt) break;
                            continue;
                        }
                        // nothing to claim: the level is over once everything appended is processed
                        const uint32_t d = ld_volatile_u32(&st->q_done);
                        const uint32_t t2 = ld_volatile_u32(&st->q_tail);
                        if (d == t2 || ld_volatile_u32(&st->error)) break;
                        if (idle_since == 0) idle_since = global_ns();
                        else if (global_ns() - idle_since > kWatchdogNs) { atomicExch(&st->error, 1u); break; }
                        __nanosleep(32);
                    }
                    s_begin = begin;
                    s_cnt = cnt;
                }
                __syncthreads();
                if (s_cnt == 0) break;  // level over
                if (threadIdx.x < s_cnt) { pending = true; my_slot = s_begin + threadIdx.x; }
            }
            // poll the owned slots once; rows that are there get processed now
            bool ready = false;
            uint32_t my_len = 0;
            uint64_t my_row = 0;
            if (pending) {
                const uint32_t v = ld_volatile_u32(&queue[my_slot]);
                if (v != kSentinel) {
                    ready = true;
                    my_row = row_ptr[v];
                    my_len = (uint32_t)(row_ptr[v + 1] - my_row);
                }
            }
            uint32_t total = 0;
            const uint32_t ex = block_excl_scan_add<uint32_t, kPeelThreads>(my_len, s_scan, &total);
            if (threadIdx.x < kMaxChunk) { s_off[threadIdx.x] = ex; s_row[threadIdx.x] = my_row; }
            if (threadIdx.x == 0) s_off[kMaxChunk] = total;
            const int n_ready = __syncthreads_count(ready);  // also publishes s_off / s_row
            if (n_ready == 0) {
                // owned slots still empty: give them up only when the level is quiescent
                if (threadIdx.x == 0) {
                    const uint32_t d = ld_volatile_u32(&st->q_done);
                    const uint32_t t2 = ld_volatile_u32(&st->q_tail);
                    uint32_t over = (d == t2 || ld_volatile_u32(&st->error)) ? 1u : 0u;
                    if (!over) {
                        if (idle_since == 0) idle_since = global_ns();
                        else if (global_ns() - idle_since > kWatchdogNs) { atomicExch(&st->error, 2u); over = 1u; }
                        __nanosleep(32);
                    }
                    s_cnt = over;
                }
                __syncthreads();
                if (s_cnt) break;  // level over (pending slots stay unwritten this level)
                continue;
            }
            idle_since = 0;

            // block-wide edge-parallel traversal of the ready rows
            for (uint32_t base = warp_in_block * 32; base < total; base += kPeelThreads) {
                const uint32_t j = base + lane;
                bool push = false;
                uint32_t u = 0;
                if (j < total) {
                    uint32_t lo = 0, hi = kMaxChunk;  // s_off[lo] <= j < s_off[hi]
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (s_off[mid] <= j) lo = mid; else hi = mid;
                    }
                    u = col[s_row[lo] + (j - s_off[lo])];
                    if (__ldcg(&deg[u]) > k) {
                        const int32_t old = atomicSub(&deg[u], 1);
                        if (old == k + 1) push = true;             // u just reached level k: ours to enqueue
                        else if (old <= k) atomicAdd(&deg[u], 1);  // already at level k: undo (clamp)
                    }
                }
                const uint32_t pm = __ballot_sync(kFullMask, push);
                if (pm) {
                    uint32_t slot = 0;
                    if (lane == 0) slot = atomicAdd(&st->q_tail, (uint32_t)__popc(pm));
                    slot = __shfl_sync(kFullMask, slot, 0);
                    if (push) {
                        volatile uint32_t *q = queue;  // consumers poll the slot leaving the sentinel value
                        q[slot + __popc(pm & lanemask_lt())] = u;
                    }
                }
            }
            if (ready) pending = false;
            __syncthreads();  // every decrement of this batch is issued before it counts as done
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(&st->q_done, (uint32_t)n_ready);
            }
        }
        grid.sync();
        // q_done == q_tail here and neither moves until the next PROCESS phase
        const uint32_t done = __ldcg(&st->q_done);
        if (done >= n || __ldcg(&st->error)) break;
        if (blockIdx.x == 0 && threadIdx.x == 0) st->q_head = done;  // un-claim overshoot / abandoned slots
        k += 1;
    }
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
    KG_CUDA(ctx, cudaMemcpyAsync(g->core, g->deg, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));

    DevBuf<uint32_t> queue, alive_a, alive_b;
    DevBuf<PeelState> state(ctx, 1);
    KG_ALLOC(ctx, queue, n);
    KG_ALLOC(ctx, alive_a, n);
    KG_ALLOC(ctx, alive_b, n);
    if (!state) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(queue.p, 0xff, (size_t)n * sizeof(uint32_t), ctx->stream));
    PeelState init{};
    for (int i = 0; i < 3; ++i) init.next_min[i] = INT32_MAX;
    KG_CUDA(ctx, cudaMemcpyAsync(state.p, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));

    int per_sm = 0;
    KG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peel_kernel, kPeelThreads, 0));
    if (per_sm < 1) return ctx_fail(ctx, KOMBGPU_ECUDA, "peel kernel does not fit on an SM");
    const int grid = per_sm * ctx->sm_count;  // persistent: every CTA resident (cooperative launch)
    uint32_t n_arg = n;
    const uint64_t *row_ptr = g->row_ptr;
    const uint32_t *col = g->col;
    int32_t *core = g->core;
    uint32_t *q = queue.p, *aa = alive_a.p, *ab = alive_b.p;
    PeelState *sp = state.p;
    void *args[] = {&n_arg, &row_ptr, &col, &core, &q, &aa, &ab, &sp};
    cudaEvent_t ev0, ev1;
    KG_CUDA(ctx, cudaEventCreate(&ev0));
    KG_CUDA(ctx, cudaEventCreate(&ev1));
    KG_CUDA(ctx, cudaEventRecord(ev0, ctx->stream));
    cudaError_t le = cudaLaunchCooperativeKernel((void *)peel_kernel, dim3(grid), dim3(kPeelThreads), args, 0, ctx->stream);
    cudaEventRecord(ev1, ctx->stream);
    ctx->launches++;
    if (le == cudaSuccess) le = cudaEventSynchronize(ev1);
    if (le == cudaSuccess) cudaEventElapsedTime(&g->st.ms_peel_kernel, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (le != cudaSuccess) return ctx_fail(ctx, KOMBGPU_ECUDA, "peel kernel: %s", cudaGetErrorString(le));

    PeelState fin{};
    KG_TRY(read_back(ctx, state.p, &fin, 1));
    if (fin.error) return ctx_fail(ctx, KOMBGPU_EINTERNAL, "peel watchdog tripped (code %u, tail %u of %u)", fin.error, fin.q_tail, n);
    if (fin.q_tail != n) return ctx_fail(ctx, KOMBGPU_EINTERNAL, "peel ended with %u of %u vertices queued", fin.q_tail, n);
    g->st.max_coreness = fin.max_core;
    g->st.peel_levels = fin.levels;
    g->st.peel_rounds = fin.rounds;
    g->has_core = true;
    return KOMBGPU_OK;
}

}  // namespace kg
