// peel_device.cuh — device code shared by the single-GPU peel (peel.cu) and the
// per-rank partition kernels of the multi-GPU path (dist.cu).
#pragma once

#include <cooperative_groups.h>

#include "graph.cuh"

namespace kg {
namespace peel {

constexpr int kPeelThreads = 512;
constexpr int kPeelWarps = kPeelThreads / 32;
constexpr int kLocalQ = 2048;        // capacity of each CTA-local vertex list (two lists: current, next)
constexpr int kBatch = 256;          // adjacency ranges one traversal covers
constexpr int kBatchLog2 = 8;
constexpr int kUnroll = 4;           // independent edge chains per thread (memory-level parallelism)
constexpr uint32_t kKeep = 64;       // discoveries a CTA keeps per generation; the surplus is shared through the pool
constexpr uint32_t kClaimMax = 1024; // pool entries one claim may take
constexpr uint32_t kSplit = 4096;    // rows longer than this are cut into slices shared by all CTAs
constexpr uint32_t kSliceLen = 2048; // edges per slice: one traversal iteration of a CTA
constexpr int kSliceLenBits = 20;
constexpr uint32_t kDirectEdges = kPeelThreads * kUnroll;  // batches up to this size skip the degree pre-load
constexpr int kScanItems = 8;        // alive-list entries per thread per scan tile
constexpr int kScanTileV = kPeelThreads * kScanItems;
constexpr unsigned long long kWatchdogNs = 10ull * 1000000000ull;  // a wait this long means a broken invariant

// Pool entries (64 bit):
//   vertex task : v                                     (bit 63 clear)
//   slice task  : 1 << 63 | first_edge << 20 | length   (a piece of a long row)
//   kEmpty      : slot not written yet
//   level token : 0xFFFFFFFE'<round>  "the level of <round> is over", written into reserved slots
constexpr uint64_t kEmpty = ~0ull;
constexpr uint64_t kSliceBit = 1ull << 63;
constexpr uint32_t kTokenHi = 0xfffffffeu;

__device__ __forceinline__ bool entry_is_task(uint64_t e) { return e != kEmpty && (uint32_t)(e >> 32) != kTokenHi; }

// Per-round scan results live in three rotating slots: round r uses slot r % 3
// and CTA 0 re-arms slot (r + 1) % 3 at the start of round r.  That slot was last
// read right after the grid barrier of round r - 2, and CTA 0 can only be in
// round r once every CTA has arrived at the last barrier of round r - 1, so
// nobody can still be reading it.  All control-flow decisions that lead to a grid
// barrier are taken from these slots (or from q_done / error at points where they
// cannot change), so every CTA takes the same path to the same barriers.
struct PeelState {
    // hot 16 bytes, read with one vector load
    uint32_t q_head;   // pool slots claimed or reserved
    uint32_t q_tail;   // pool slots appended (monotone over the whole run)
    uint32_t q_done;   // pool tasks fully processed (credited when the claimer's local cascade has drained)
    uint32_t error;    // watchdog / invariant flag
    uint32_t alive_out[3];  // survivors written by the scan of round r (slot r % 3)
    uint32_t front_cnt[3];  // vertices that scan appended to the pool
    int32_t next_min[3];    // min degree of the survivors
    uint32_t levels;        // non-empty levels
    uint32_t rounds;        // scan phases executed
    int32_t max_core;
    unsigned long long n_removed;  // vertices peeled through the pool
    unsigned long long n_isolated; // degree-0 vertices, peeled by the level-0 scan itself (n_removed + n_isolated must end at n)
    unsigned long long shared;     // discoveries handed to other CTAs through the pool
    unsigned long long sliced;     // slices published
    // CTA 0's view of where the time goes (ns): scan, barrier after scan, process, barrier after process
    unsigned long long prof_ns[4];
    unsigned long long batches;    // traversals over all CTAs
    unsigned long long carried;    // warp mode: discoveries handed on in registers (followed cascades)
    unsigned long long ring_pushed; // warp mode: discoveries and row pieces queued in a CTA's ring
    unsigned long long *trace;     // optional (KOMBGPU_TRACE): 6 words per round, CTA 0's view
    uint32_t trace_cap;
    uint32_t tune[8];              // warp mode knobs (peel_warp.cuh WarpTune): keep, wsplit, park_ns, thin, hub_slice
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 ld_volatile_u4(const void *p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct BlockShared {
    uint32_t list[2][kLocalQ];   // current / next CTA-local vertex lists
    uint32_t next_cnt;           // entries pushed to the next list (clamped to kLocalQ when read)
    uint64_t row[kBatch];        // first edge of every range of the batch
    uint32_t off[kBatch + 1];    // exclusive prefix of the range lengths
    uint32_t scan[kPeelWarps + 1];
    uint32_t tile_base[3];       // scan: where this tile's frontier / survivor entries go
    uint32_t ctl[4];             // thread 0 -> CTA broadcasts
    uint32_t fin[2];             // [tail, head) to fill with level tokens (set by the CTA that ends the level)
    uint32_t ready_mask[kBatch / 32];  // which of the polled pool slots hold a task
};

// kDist: the CTA works on one rank's vertex range [part.v_lo, part.v_lo + part.n_local).  col[] holds GLOBAL
// ids, deg[] / lists / pool hold LOCAL ids.  A neighbour owned by another rank is not touched here: its
// global id is appended to the outbox and the owner applies the decrement after the exchange.
struct PartView {
    uint32_t v_lo = 0;
    uint32_t n_local = 0;
    uint32_t *outbox = nullptr;      // global ids of remote decrement targets
    uint32_t *outbox_cnt = nullptr;
};

// SCAN of one level for one CTA: one pass over its tiles of the alive list.
// Vertices at deg == k are appended to the pool, vertices above k are compacted
// into alive_dst, the rest (peeled at an earlier level) are dropped.  Positions
// come from a CTA-wide scan, so a tile of kScanTileV entries costs three global
// atomics in all.  Returns the thread's minimum survivor degree.
__device__ __forceinline__ int32_t scan_alive(const int32_t k, const uint32_t *alive_src, const uint32_t n_alive,
                                              uint32_t *alive_dst, const int32_t *deg, uint64_t *Q, uint32_t *q_tail,
                                              uint32_t *front_cnt, uint32_t *alive_out, unsigned long long *n_isolated,
                                              BlockShared &sh, long long *stamps = nullptr) {
    const uint32_t tid = threadIdx.x;
    if (stamps) stamps[0] = clock64();
    int32_t local_min = INT32_MAX;
    uint32_t isolated = 0;
    // short lists are spread over all CTAs: a tile of `items` entries per thread, items = 1..kScanItems.  (A CTA
    // gathering 4096 random degrees is bound by its own L1->L2 request rate: ~7000 cycles measured, against
    // ~1200 for 512.)
    const uint32_t per_cta = (uint32_t)(((uint64_t)n_alive + gridDim.x - 1) / gridDim.x);
    const int items = (int)min(max((per_cta + kPeelThreads - 1) / kPeelThreads, 1u), (uint32_t)kScanItems);
    const uint64_t tile_v = (uint64_t)items * kPeelThreads;
    for (uint64_t tile = (uint64_t)blockIdx.x * tile_v; tile < n_alive; tile += (uint64_t)gridDim.x * tile_v) {
        uint32_t v[kScanItems];
        uint32_t flag[kScanItems];  // 1 = frontier, 0x10000 = survivor
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint64_t i = (j < items) ? tile + (uint64_t)j * kPeelThreads + tid : ~0ull;
            flag[j] = 0;
            v[j] = 0;
            if (i < n_alive) v[j] = alive_src ? __ldcg(&alive_src[i]) : (uint32_t)i;  // rewritten every round: skip L1
        }
        if (stamps && tile == 0) { uint32_t x = 0; for (int j = 0; j < kScanItems; ++j) x ^= v[j]; if (x == 0xfffffff3u) stamps[4] = 1; stamps[1] = clock64(); }
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint64_t i = (j < items) ? tile + (uint64_t)j * kPeelThreads + tid : ~0ull;
            if (i < n_alive) {
                const int32_t d = __ldcg(&deg[v[j]]);
                if (d == k) { if (k > 0) flag[j] = 1u; else ++isolated; }  // a degree-0 vertex has no row to walk: peeled right here
                else if (d > k) { flag[j] = 0x10000u; local_min = min(local_min, d); }
            }
            mine += flag[j];
        }
        if (stamps && tile == 0) { if (mine == 0xfffffff3u) stamps[4] = 1; stamps[2] = clock64(); }
        uint32_t total = 0;
        uint32_t ex = block_excl_scan_add<uint32_t, kPeelThreads>(mine, sh.scan, &total);  // both counts: 16 bits each
        if (tid == 0) {
            const uint32_t nf = total & 0xffffu, ns = total >> 16;
            sh.tile_base[0] = nf ? atomicAdd(q_tail, nf) : 0;
            sh.tile_base[1] = ns ? atomicAdd(alive_out, ns) : 0;
            if (nf) atomicAdd(front_cnt, nf);
        }
        __syncthreads();
        if (stamps && tile == 0) stamps[3] = clock64();
        uint32_t fpos = sh.tile_base[0] + (ex & 0xffffu), spos = sh.tile_base[1] + (ex >> 16);
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            if (flag[j] == 1u) Q[fpos++] = (uint64_t)v[j];
            else if (flag[j]) alive_dst[spos++] = v[j];
        }
        __syncthreads();  // tile_base is reused by the next tile
    }
    if (k == 0) {
        isolated = warp_reduce_add(isolated);
        if (lane_id() == 0 && isolated) atomicAdd(n_isolated, (unsigned long long)isolated);
    }
    return local_min;
}

// Edge-parallel traversal of the batch described by sh.row / sh.off (kBatch ranges,
// `total` edges) by the whole CTA: for every neighbour u with deg[u] > k the
// degree is decremented; the thread whose decrement takes it to k owns u and
// pushes it to the CTA's next list.  A decrement that lands below k is undone,
// so deg[] is clamped at k and ends as the coreness.  Every thread keeps kUnroll
// independent col -> deg -> atomic chains in flight.
template <bool kDist>
__device__ __forceinline__ void traverse_batch(const uint32_t total, const int32_t k, const uint32_t *__restrict__ col,
                                               int32_t *deg, uint32_t *next, uint64_t *Q, const uint32_t cap, PeelState *st,
                                               BlockShared &sh, const PartView &part) {
    const uint32_t tid = threadIdx.x, lane = lane_id();
    const bool direct = total <= kDirectEdges;
    for (uint32_t base = 0; base < total; base += kPeelThreads * kUnroll) {
        uint32_t u[kUnroll];
        int32_t d[kUnroll];
        bool push[kUnroll];
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const uint32_t e = base + t * kPeelThreads + tid;
            u[t] = kFullMask;
            if (e < total) {
                uint32_t lo_i = 0, hi_i = kBatch;  // off[lo_i] <= e < off[hi_i]
#pragma unroll
                for (int sgm = 0; sgm < kBatchLog2; ++sgm) {
                    const uint32_t mid = (lo_i + hi_i) >> 1;
                    if (sh.off[mid] <= e) lo_i = mid; else hi_i = mid;
                }
                u[t] = col[sh.row[lo_i] + (e - sh.off[lo_i])];
            }
        }
        if (kDist) {
            // split off the neighbours other ranks own: ship their ids, keep local ones as local ids
#pragma unroll
            for (int t = 0; t < kUnroll; ++t) {
                const bool valid = u[t] != kFullMask;
                const uint32_t loc = u[t] - part.v_lo;
                const bool remote = valid && loc >= part.n_local;
                const uint32_t rm = __ballot_sync(kFullMask, remote);
                if (rm) {
                    uint32_t pos = 0;
                    if (lane == 0) pos = atomicAdd(part.outbox_cnt, (uint32_t)__popc(rm));
                    pos = __shfl_sync(kFullMask, pos, 0) + __popc(rm & lanemask_lt());
                    if (remote) part.outbox[pos] = u[t];
                }
                u[t] = (valid && !remote) ? loc : kFullMask;
            }
        }
        if (direct) {
            // latency-bound batch (one iteration): no degree pre-load, the decrement goes out right away and
            // is undone if it lands at or below k -- one dependent round trip less on the cascade's critical path
#pragma unroll
            for (int t = 0; t < kUnroll; ++t) d[t] = (u[t] != kFullMask) ? atomicSub(&deg[u[t]], 1) : INT32_MAX;
#pragma unroll
            for (int t = 0; t < kUnroll; ++t) {
                push[t] = d[t] == k + 1;
                if (d[t] <= k) atomicAdd(&deg[u[t]], 1);
            }
        } else {
#pragma unroll
            for (int t = 0; t < kUnroll; ++t) d[t] = (u[t] != kFullMask) ? __ldcg(&deg[u[t]]) : INT32_MIN;
#pragma unroll
            for (int t = 0; t < kUnroll; ++t) {
                push[t] = false;
                if (d[t] > k) {
                    const int32_t old = atomicSub(&deg[u[t]], 1);
                    if (old == k + 1) push[t] = true;             // u just reached level k: ours to peel
                    else if (old <= k) atomicAdd(&deg[u[t]], 1);  // already at level k: undo (clamp)
                }
            }
        }
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const uint32_t pm = __ballot_sync(kFullMask, push[t]);
            if (pm == 0) continue;
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&sh.next_cnt, (uint32_t)__popc(pm));
            pos = __shfl_sync(kFullMask, pos, 0) + __popc(pm & lanemask_lt());
            if (push[t]) {
                if (pos < kLocalQ) {
                    next[pos] = u[t];
                } else {
                    // the CTA-local list is full: this discovery goes straight to the pool
                    const uint32_t idx = atomicAdd(&st->q_tail, 1u);
                    if (idx < cap) st_volatile_u64(&Q[idx], (uint64_t)u[t]);
                    else atomicExch(&st->error, 3u);
                }
            }
        }
    }
}

// PROCESS phase of one level for one CTA.
//
// Two mechanisms, one for each thing that bounds a level:
//  * latency of cascade chains: vertices a CTA discovers (its decrement took deg[u] to k) go to the CTA's own
//    shared-memory list and are processed by the same CTA next.  A chain step costs row_ptr -> col -> atomic round
//    trips and a few CTA barriers; no global queue, no other CTA.
//  * balance: everything else goes through the POOL, a ticket queue in global memory.  The scan appends the level's
//    frontier to it; a CTA that discovers more than kKeep vertices in one generation appends the surplus; rows
//    longer than kSplit edges are appended as slices.  q_head / q_tail only move by fetch-add.  A CTA with nothing
//    left claims the next entries; if the pool is empty it RESERVES the next slot and polls that slot's own address,
//    so whoever appends the next task hands it to exactly that CTA with one store (no contended counter).
// Credit for claimed tasks is added to q_done only when the CTA's local cascade has drained, which makes
// "q_done == q_tail" a sound and final quiescence test.  The CTA whose credit makes them equal ends the level: it
// writes a level token into every reserved slot [q_tail, q_head).  CTAs that reserved later notice through a slow
// periodic check of the counters.  CTA 0 pulls q_head back to q_tail before the next level.
template <bool kDist>
__device__ __forceinline__ uint32_t process_level(const int32_t k, const uint32_t round, uint64_t *Q, const uint32_t cap,
                                                  const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                                  int32_t *deg, PeelState *st, BlockShared &sh, const PartView &part) {
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t token = ((uint64_t)kTokenHi << 32) | round;
    uint32_t cur = 0;          // index of the current local list
    uint32_t rb = 0, re = 0;   // owned pool range [rb, re)
    uint32_t removed = 0;
    // thread 0 only:
    uint32_t credit = 0, polls = 0, batches = 0, shared = 0, sliced = 0;
    unsigned long long idle_since = 0;
    if (tid == 0) sh.next_cnt = 0;
    __syncthreads();

    while (true) {
        uint64_t ent = kEmpty;  // this thread's task of the coming batch (threads < kBatch)
        uint32_t n_next = min(sh.next_cnt, (uint32_t)kLocalQ);
        __syncthreads();  // everyone has read next_cnt
        bool from_pool = false;
        uint32_t m = 0;
        if (n_next > 0) {
            // ---- own discoveries first: they are the critical path of the cascade ----
            cur ^= 1;
            if (tid == 0) sh.next_cnt = 0;
            if (n_next > kKeep) {
                // share the surplus: other CTAs (parked on reserved slots) pick it up at once
                const uint32_t surplus = n_next - kKeep;
                if (tid == 0) { sh.ctl[0] = atomicAdd(&st->q_tail, surplus); shared += surplus; }
                __syncthreads();
                const uint32_t base = sh.ctl[0];
                for (uint32_t i = tid; i < surplus; i += kPeelThreads) {
                    if (base + i < cap) st_volatile_u64(&Q[base + i], (uint64_t)sh.list[cur][kKeep + i]);
                    else atomicExch(&st->error, 3u);
                }
                n_next = kKeep;
            }
            m = n_next;
            if (tid < m) ent = (uint64_t)sh.list[cur][tid];
        } else {
            // ---- local cascade drained: take work from the pool ----
            if (rb == re) {
                // owned range exhausted: settle the credit (everything claimed so far is processed and its
                // cascades have drained), then claim more -- or reserve the slot the next task will land in
                if (tid == 0) {
                    uint32_t over = 0, begin = 0, end = 0, fin_tail = 0, fin_head = 0;
                    uint4 a;
                    if (credit) {
                        __threadfence();
                        const uint32_t old = atomicAdd(&st->q_done, credit);
                        a = ld_volatile_u4(st);  // q_head, q_tail, q_done, error
                        if (old + credit == a.y) { fin_tail = a.y; fin_head = min(a.x, cap); over = 1; }  // quiescent, final
                        credit = 0;
                    } else {
                        a = ld_volatile_u4(st);
                    }
                    if (!over) {
                        if (a.z == a.y || a.w) {
                            over = 1;  // q_done == q_tail: the level has ended
                        } else {
                            uint32_t take = 1;
                            if (a.x < a.y) take = min(max((a.y - a.x + gridDim.x - 1) / gridDim.x, 1u), kClaimMax);
                            begin = atomicAdd(&st->q_head, take);
                            end = min(begin + take, cap);
                            if (begin >= cap) { atomicExch(&st->error, 4u); over = 1; }
                        }
                    }
                    sh.ctl[0] = over; sh.ctl[1] = begin; sh.ctl[2] = end;
                    sh.fin[0] = fin_tail; sh.fin[1] = fin_head;
                    polls = 0;
                }
                __syncthreads();
                {   // this CTA ended the level: wake every CTA parked on a reserved slot
                    const uint32_t fin_tail = sh.fin[0], fin_head = sh.fin[1];
                    for (uint32_t i = fin_tail + tid; i < fin_head; i += kPeelThreads) st_volatile_u64(&Q[i], token);
                }
                if (sh.ctl[0]) break;
                rb = sh.ctl[1];
                re = sh.ctl[2];
            }
            // poll the head of the owned range; take the leading run of written slots
            const uint32_t m_try = min((uint32_t)kBatch, re - rb);
            bool ready = false, saw_token = false;
            if (tid < m_try) {
                ent = ld_volatile_u64(&Q[rb + tid]);
                if (ent == token) { saw_token = true; st_volatile_u64(&Q[rb + tid], kEmpty); }  // leave the slot clean
                ready = entry_is_task(ent);
            }
            if (warp < kBatch / 32) {
                const uint32_t rm = __ballot_sync(kFullMask, ready);
                if (lane == 0) sh.ready_mask[warp] = rm;
            }
            if (__syncthreads_or(saw_token)) break;  // level over: nothing can be pending anywhere
            m = 0;
#pragma unroll
            for (int w = 0; w < kBatch / 32; ++w) {   // length of the leading run of written slots
                const uint32_t rm = sh.ready_mask[w];
                if (m == 32u * w) m += (rm == kFullMask) ? 32u : (uint32_t)__ffs(~rm) - 1u;
            }
            m = min(m, m_try);
            if (m == 0) {
                // owned slots still empty.  Settle the credit first (nothing else may hold the level open),
                // then poll; only rarely look at the shared counters
                if (tid == 0) {
                    uint32_t over = 0, fin_tail = 0, fin_head = 0;
                    if (credit) {
                        __threadfence();
                        const uint32_t old = atomicAdd(&st->q_done, credit);
                        const uint4 a = ld_volatile_u4(st);
                        if (old + credit == a.y) { fin_tail = a.y; fin_head = min(a.x, cap); over = 1; }
                        credit = 0;
                    } else if (++polls == 1 || (polls & 7u) == 0) {
                        const uint4 a = ld_volatile_u4(st);
                        if (a.z == a.y || a.w) over = 1;  // q_done == q_tail: final, nobody can append any more
                        else if (idle_since == 0) idle_since = global_ns();
                        else if (global_ns() - idle_since > kWatchdogNs) { atomicExch(&st->error, 2u); over = 1; }
                    }
                    if (!over) __nanosleep(100);
                    sh.ctl[0] = over;
                    sh.fin[0] = fin_tail; sh.fin[1] = fin_head;
                }
                __syncthreads();
                {
                    const uint32_t fin_tail = sh.fin[0], fin_head = sh.fin[1];
                    for (uint32_t i = fin_tail + tid; i < fin_head; i += kPeelThreads) st_volatile_u64(&Q[i], token);
                }
                if (sh.ctl[0]) break;
                continue;
            }
            if (tid >= m) ent = kEmpty;
            rb += m;
            from_pool = true;
            if (tid == 0) { credit += m; idle_since = 0; }
        }

        // ---- batch assembly: one adjacency range per task --------------------------------
        uint32_t my_len = 0;
        uint64_t my_row = 0;
        bool is_vertex = false;
        if (ent != kEmpty) {
            if (ent & kSliceBit) {
                my_row = (ent & ~kSliceBit) >> kSliceLenBits;
                my_len = (uint32_t)(ent & ((1u << kSliceLenBits) - 1));
            } else {
                const uint32_t v = (uint32_t)ent;
                is_vertex = true;
                my_row = row_ptr[v];
                my_len = (uint32_t)(row_ptr[v + 1] - my_row);
                if (my_len > kSplit) {
                    // hub row: hand it to the whole grid as slices
                    const uint32_t n_sl = (my_len + kSliceLen - 1) / kSliceLen;
                    const uint32_t s0 = atomicAdd(&st->q_tail, n_sl);
                    for (uint32_t i = 0; i < n_sl; ++i) {
                        const uint64_t e = kSliceBit | ((my_row + (uint64_t)i * kSliceLen) << kSliceLenBits) |
                                           min(kSliceLen, my_len - i * kSliceLen);
                        if (s0 + i < cap) st_volatile_u64(&Q[s0 + i], e);
                        else atomicExch(&st->error, 3u);
                    }
                    atomicAdd(&st->sliced, (unsigned long long)n_sl);
                    my_len = 0;
                }
            }
        }
        (void)from_pool;
        uint32_t total = 0;
        const uint32_t ex = block_excl_scan_add<uint32_t, kPeelThreads>(my_len, sh.scan, &total);
        if (tid < kBatch) { sh.off[tid] = ex; sh.row[tid] = my_row; }
        if (tid == 0) { sh.off[kBatch] = total; ++batches; }
        removed += __syncthreads_count(is_vertex);  // also publishes off / row
        traverse_batch<kDist>(total, k, col, deg, sh.list[cur ^ 1], Q, cap, st, sh, part);
        __syncthreads();  // row/off are reused by the next batch; next_cnt is complete
    }
    if (tid == 0) {
        if (shared) atomicAdd(&st->shared, (unsigned long long)shared);
        if (batches) atomicAdd(&st->batches, (unsigned long long)batches);
        (void)sliced;
    }
    return removed;
}

}  // namespace peel
}  // namespace kg
