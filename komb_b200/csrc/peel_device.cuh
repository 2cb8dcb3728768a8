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
constexpr int kBatch = 64;           // adjacency ranges one traversal covers
constexpr int kUnroll = 4;           // independent edge chains per thread (memory-level parallelism)
constexpr uint32_t kSplit = 2048;    // rows longer than this are cut into slices shared by all CTAs
constexpr uint32_t kSliceLen = 2048; // edges per slice: one traversal iteration of a CTA
constexpr int kSliceLenBits = 20;    // slice entry = first_edge << 20 | length
constexpr uint32_t kDirectEdges = kPeelThreads * kUnroll;  // batches up to this size skip the degree pre-load
constexpr int kScanItems = 8;        // alive-list entries per thread per scan tile
constexpr int kScanTileV = kPeelThreads * kScanItems;

// Per-round results live in three rotating slots: round r uses slot r % 3 and
// CTA 0 re-arms slot (r + 1) % 3 at the start of round r.  That slot was last
// read right after a grid barrier of round r - 2, and CTA 0 can only be in
// round r once every CTA has arrived at the last barrier of round r - 1, so
// nobody can still be reading it.  All control-flow decisions are taken from
// these slots at points where they cannot change, so every CTA takes the same
// path to the same barriers.
struct PeelState {
    uint32_t alive_out[3];  // survivors written by the scan of round r (slot r % 3)
    uint32_t front_cnt[3];  // length of the level's frontier list: scan output + CTA-list overflow
    uint32_t slice_cnt[3];  // length of the level's slice list (pieces of long rows)
    int32_t next_min[3];    // min degree of the survivors
    uint32_t error;
    uint32_t levels;        // non-empty levels
    uint32_t rounds;        // scan phases executed
    uint32_t subrounds;     // process phases executed (grid-wide)
    int32_t max_core;
    unsigned long long n_removed;  // vertices peeled (must end at n)
    unsigned long long overflowed; // discoveries that did not fit a CTA-local list
    unsigned long long sliced;     // slices published
    // CTA 0's view of where the time goes (ns): scan, barrier after scan, process, barrier after process
    unsigned long long prof_ns[4];
    unsigned long long batches;    // traversals over all CTAs
    unsigned long long *trace;     // optional (KOMBGPU_TRACE): 6 words per round, CTA 0's view
    uint32_t trace_cap;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct BlockShared {
    uint32_t list[2][kLocalQ];   // current / next CTA-local vertex lists
    uint32_t next_cnt;           // entries pushed to the next list (may exceed kLocalQ: the excess went global)
    uint64_t row[kBatch];        // first edge of every range of the batch
    uint32_t off[kBatch + 1];    // exclusive prefix of the range lengths
    uint32_t scan[kPeelWarps + 1];
    uint32_t tile_base[2];       // scan: where this tile's frontier / survivor entries go
};

// Edge-parallel traversal of the batch described by sh.row / sh.off (kBatch ranges,
// `total` edges) by the whole CTA: for every neighbour u with deg[u] > k the
// degree is decremented; the thread whose decrement takes it to k owns u and
// pushes it to the CTA's next list.  A decrement that lands below k is undone,
// so deg[] is clamped at k and ends as the coreness.  Every thread keeps kUnroll
// independent col -> deg -> atomic chains in flight.
//
// kDist: the CTA works on one rank's vertex range [part.v_lo, part.v_lo + part.n_local).  col[] holds GLOBAL
// ids, deg[] / lists hold LOCAL ids.  A neighbour owned by another rank is not touched here: its global id
// is appended to the outbox and the owner applies the decrement after the exchange.
struct PartView {
    uint32_t v_lo = 0;
    uint32_t n_local = 0;
    uint32_t *outbox = nullptr;      // global ids of remote decrement targets
    uint32_t *outbox_cnt = nullptr;
};

template <bool kDist>
__device__ __forceinline__ void traverse_batch(const uint32_t total, const int32_t k, const uint32_t *__restrict__ col,
                                               int32_t *deg, uint32_t *next, uint32_t *F, uint32_t *front_cnt,
                                               BlockShared &sh, uint32_t &overflowed, const PartView &part) {
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
                for (int sgm = 0; sgm < 6; ++sgm) {
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
                if (pos < kLocalQ) next[pos] = u[t];
                else { F[atomicAdd(front_cnt, 1u)] = u[t]; ++overflowed; }  // rare: the next sub-round takes it
            }
        }
    }
}

// SCAN of one level for one CTA: one pass over its tiles of the alive list.
// Vertices at deg == k go to the frontier list F, vertices above k are compacted
// into alive_dst, the rest (peeled at an earlier level) are dropped.  Positions
// come from a CTA-wide scan, so a tile of kScanTileV entries costs two global
// atomics in all.  Returns the thread's minimum survivor degree.
__device__ __forceinline__ int32_t scan_alive(const int32_t k, const uint32_t *alive_src, const uint32_t n_alive,
                                              uint32_t *alive_dst, const int32_t *deg, uint32_t *F, uint32_t *front_cnt,
                                              uint32_t *alive_out, BlockShared &sh) {
    const uint32_t tid = threadIdx.x;
    int32_t local_min = INT32_MAX;
    for (uint64_t tile = (uint64_t)blockIdx.x * kScanTileV; tile < n_alive; tile += (uint64_t)gridDim.x * kScanTileV) {
        uint32_t v[kScanItems];
        uint32_t flag[kScanItems];  // 1 = frontier, 0x10000 = survivor
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint64_t i = tile + (uint64_t)j * kPeelThreads + tid;
            flag[j] = 0;
            v[j] = 0;
            if (i < n_alive) v[j] = alive_src ? __ldcg(&alive_src[i]) : (uint32_t)i;  // rewritten every round: skip L1
        }
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint64_t i = tile + (uint64_t)j * kPeelThreads + tid;
            if (i < n_alive) {
                const int32_t d = __ldcg(&deg[v[j]]);
                if (d == k) flag[j] = 1u;
                else if (d > k) { flag[j] = 0x10000u; local_min = min(local_min, d); }
            }
            mine += flag[j];
        }
        uint32_t total = 0;
        uint32_t ex = block_excl_scan_add<uint32_t, kPeelThreads>(mine, sh.scan, &total);  // both counts: 16 bits each
        if (tid == 0) {
            const uint32_t nf = total & 0xffffu, ns = total >> 16;
            sh.tile_base[0] = nf ? atomicAdd(front_cnt, nf) : 0;
            sh.tile_base[1] = ns ? atomicAdd(alive_out, ns) : 0;
        }
        __syncthreads();
        uint32_t fpos = sh.tile_base[0] + (ex & 0xffffu), spos = sh.tile_base[1] + (ex >> 16);
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            if (flag[j] == 1u) F[fpos++] = v[j];
            else if (flag[j]) alive_dst[spos++] = v[j];
        }
        __syncthreads();  // tile_base is reused by the next tile
    }
    return local_min;
}

// PROCESS phase of one sub-round for one CTA (PKC-style, CTA-local cascade).
//
// Work of a sub-round: the slice list S[s_lo, s_hi) (pieces of long rows) and the
// frontier list F[f_lo, f_hi) (vertices from the scan, or overflow of an earlier
// sub-round).  Both are dealt to the CTAs round-robin: no claiming, no atomics.
// A CTA traverses a batch of ranges edge-parallel with all its threads.
// Vertices it discovers go to its own shared-memory "next" list and are
// processed by the same CTA right after the current list: a cascade chain costs
// row_ptr -> col -> deg -> atomic round trips and a few CTA barriers per step,
// and never touches a global queue or waits for another CTA.  A row longer than
// kSplit is not traversed by the CTA that meets it: it is cut into slices
// appended to S for the next sub-round, so a hub is shared by the whole grid.
// What does not fit the shared-memory list is appended to F for the next
// sub-round as well.
template <bool kDist>
__device__ __forceinline__ uint32_t process_subround(const int32_t k, uint32_t *F, const uint32_t f_lo, const uint32_t f_hi,
                                                     uint32_t *front_cnt, uint64_t *S, const uint32_t s_lo,
                                                     const uint32_t s_hi, uint32_t *slice_cnt,
                                                     const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                                     int32_t *deg, PeelState *st, BlockShared &sh, const PartView &part) {
    const uint32_t tid = threadIdx.x, lane = lane_id();
    uint32_t cur = 0;  // index of the current list
    uint32_t removed = 0, batches = 0, overflowed = 0, sliced = 0;
    if (tid == 0) sh.next_cnt = 0;
    __syncthreads();

    // ---- slices: kBatch of them per traversal, dealt round-robin ----
    const uint32_t n_slices = s_hi - s_lo;
    // as many slices per traversal as keeps every CTA busy: a tail level has a few dozen slices, and
    // dealing them kBatch at a time would serialise them on one or two CTAs
    const uint32_t slice_sz = min((uint32_t)kBatch, max(1u, (n_slices + gridDim.x - 1) / gridDim.x));
    for (uint32_t g0 = blockIdx.x * slice_sz; g0 < n_slices; g0 += gridDim.x * slice_sz) {
        uint32_t my_len = 0;
        uint64_t my_row = 0;
        if (tid < slice_sz && g0 + tid < n_slices) {
            const uint64_t e = __ldcg(&S[s_lo + g0 + tid]);
            my_row = e >> kSliceLenBits;
            my_len = (uint32_t)(e & ((1u << kSliceLenBits) - 1));
        }
        uint32_t total = 0;
        const uint32_t ex = block_excl_scan_add<uint32_t, kPeelThreads>(my_len, sh.scan, &total);
        if (tid < kBatch) { sh.off[tid] = ex; sh.row[tid] = my_row; }
        if (tid == 0) sh.off[kBatch] = total;
        __syncthreads();
        ++batches;
        traverse_batch<kDist>(total, k, col, deg, sh.list[cur ^ 1], F, front_cnt, sh, overflowed, part);
        __syncthreads();
    }

    // ---- vertices: own discoveries first, then the CTA's share of the frontier list ----
    const uint32_t n_front = f_hi - f_lo;
    const uint32_t chunk_sz = min((uint32_t)kBatch, max(1u, (n_front + gridDim.x - 1) / gridDim.x));
    const uint32_t n_chunks = (n_front + chunk_sz - 1) / chunk_sz;
    // start dealing where the slices stopped, so that CTA 0 does not get the first share of both
    uint32_t chunk = (blockIdx.x + gridDim.x - ((n_slices + slice_sz - 1) / slice_sz) % gridDim.x) % gridDim.x;
    while (true) {
        uint32_t n_cur = min(sh.next_cnt, (uint32_t)kLocalQ);
        __syncthreads();  // everyone has read next_cnt
        if (n_cur > 0) {
            cur ^= 1;     // discoveries first: they are the critical path of the cascade
            if (tid == 0) sh.next_cnt = 0;
        } else if (chunk < n_chunks) {
            const uint32_t b = f_lo + chunk * chunk_sz;
            n_cur = min(chunk_sz, f_hi - b);
            if (tid < n_cur) sh.list[cur][tid] = __ldcg(&F[b + tid]);  // F is rewritten every level: skip L1
            chunk += gridDim.x;
        } else {
            break;
        }
        __syncthreads();
        const uint32_t *list = sh.list[cur];
        for (uint32_t b0 = 0; b0 < n_cur; b0 += kBatch) {
            const uint32_t m = min((uint32_t)kBatch, n_cur - b0);
            uint32_t my_len = 0;
            uint64_t my_row = 0;
            if (tid < m) {
                const uint32_t v = list[b0 + tid];
                my_row = row_ptr[v];
                my_len = (uint32_t)(row_ptr[v + 1] - my_row);
                if (my_len > kSplit) {
                    // hub row: hand it to the whole grid as slices of the next sub-round
                    const uint32_t n_sl = (my_len + kSliceLen - 1) / kSliceLen;
                    const uint32_t s0 = atomicAdd(slice_cnt, n_sl);
                    for (uint32_t i = 0; i < n_sl; ++i)
                        S[s0 + i] = ((my_row + (uint64_t)i * kSliceLen) << kSliceLenBits) | min(kSliceLen, my_len - i * kSliceLen);
                    sliced += n_sl;
                    my_len = 0;
                }
            }
            uint32_t total = 0;
            const uint32_t ex = block_excl_scan_add<uint32_t, kPeelThreads>(my_len, sh.scan, &total);
            if (tid < kBatch) { sh.off[tid] = ex; sh.row[tid] = my_row; }
            if (tid == 0) sh.off[kBatch] = total;
            __syncthreads();
            ++batches;
            traverse_batch<kDist>(total, k, col, deg, sh.list[cur ^ 1], F, front_cnt, sh, overflowed, part);
            __syncthreads();  // row/off are reused by the next batch; next_cnt is complete
        }
        removed += n_cur;
    }
    overflowed = warp_reduce_add(overflowed);
    sliced = warp_reduce_add(sliced);
    if (lane == 0 && overflowed) atomicAdd(&st->overflowed, (unsigned long long)overflowed);
    if (lane == 0 && sliced) atomicAdd(&st->sliced, (unsigned long long)sliced);
    if (tid == 0 && batches) atomicAdd(&st->batches, (unsigned long long)batches);
    return removed;
}


}  // namespace peel
}  // namespace kg
