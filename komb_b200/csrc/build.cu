// build.cu — stage 1 of the KOMB hot path on the device:
//   hits (read_key, unitig)  ->  per-read unitig sets  ->  clique pairs
//   ->  simple undirected edge set  ->  CSR of the symmetric graph.
//
// Restates, as sort/scan/compact kernels over integer arrays, what the reference
// does with string-keyed hash maps:
//   per-read set insert + mate union   src/graph.cpp:235,259-285   sort + adjacent-unique of (read, unitig)
//   all i<j pairs of every set         src/graph.cpp:332-347       segment scan + load-balanced pair emission
//   dedup / igraph_simplify            src/graph.cpp:342-347,438   sort + adjacent-unique of (min, max)
//   igraph_create adjacency index      src/graph.cpp:418           boundary detection + scan -> CSR
// No atomics are needed anywhere in this stage: every count falls out of run
// boundaries in sorted arrays, so the result (including the order inside each
// CSR row) is deterministic.
#include <cstdlib>

#include "graph.cuh"
#include "primitives.cuh"

namespace kg {

namespace {

constexpr int kThreads = 256;

inline uint32_t grid_for(uint64_t n, int per_block) { return ceil_div_u64(n ? n : 1, per_block); }

// ---- packing ---------------------------------------------------------------

// key = read_key << 32 | unitig; also the max read key (for the sort plan), an out-of-range flag, and how the
// read keys are ordered: info[0] = max read key, info[1] = error flag, info[2] = number of descents
// (read_key[i] < read_key[i-1]), info[3] = position of the first one.  SAM files list reads in input order, so the
// two mate files arrive as two runs of non-decreasing keys (one descent, at the file boundary).
__global__ void __launch_bounds__(kThreads) pack_hits_kernel(const uint32_t *__restrict__ read_key,
                                                             const uint32_t *__restrict__ unitig, uint64_t n_hits,
                                                             uint32_t n_vertices, uint64_t *__restrict__ keys,
                                                             uint32_t *__restrict__ info) {
    uint32_t local_max = 0, descents = 0, first = 0xffffffffu;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_hits; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r = read_key[i], u = unitig[i];
        bad |= (u >= n_vertices);
        local_max = max(local_max, r);
        if (i > 0 && r < read_key[i - 1]) { ++descents; first = min(first, (uint32_t)i); }
        keys[i] = ((uint64_t)r << 32) | u;
    }
    local_max = warp_reduce_max(local_max);
    descents = warp_reduce_add(descents);
    first = warp_reduce_min(first);
    if (lane_id() == 0) {
        if (local_max) atomicMax(&info[0], local_max);
        if (descents) { atomicAdd(&info[2], descents); atomicMin(&info[3], first); }
    }
    if (__any_sync(kFullMask, bad) && lane_id() == 0) atomicOr(&info[1], 1u);
}

// ---- reads in file order: merge the two runs, no sort -----------------------------------------------------

constexpr int kMergeItems = 8;
constexpr int kMergeTile = kThreads * kMergeItems;

__device__ __forceinline__ uint32_t read_of(uint64_t key) { return (uint32_t)(key >> 32); }

// number of A elements among the first `diag` outputs of the stable merge (A before B on equal read keys)
template <typename KeyA, typename KeyB>
__device__ __forceinline__ uint32_t merge_path(uint32_t diag, uint32_t n_a, uint32_t n_b, KeyA key_a, KeyB key_b) {
    uint32_t lo = diag > n_b ? diag - n_b : 0, hi = min(diag, n_a);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (key_a(mid) <= key_b(diag - 1 - mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// part[t] = A elements before output t * kMergeTile
__global__ void __launch_bounds__(kThreads) merge_partition_kernel(const uint64_t *__restrict__ a, uint32_t n_a,
                                                                   const uint64_t *__restrict__ b, uint32_t n_b,
                                                                   uint32_t n_tiles, uint32_t *__restrict__ part) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const uint32_t diag = (uint32_t)min((uint64_t)t * kMergeTile, (uint64_t)n_a + n_b);
    part[t] = merge_path(diag, n_a, n_b, [&](uint32_t i) { return read_of(a[i]); }, [&](uint32_t j) { return read_of(b[j]); });
}

// stable merge by read key of two runs of packed hits; the unitigs of a read stay in arrival order
__global__ void __launch_bounds__(kThreads) merge_runs_kernel(const uint64_t *__restrict__ a, uint32_t n_a,
                                                              const uint64_t *__restrict__ b, uint32_t n_b,
                                                              const uint32_t *__restrict__ part, uint64_t *__restrict__ out) {
    __shared__ uint64_t s_key[kMergeTile];
    const uint32_t tile = blockIdx.x;
    const uint64_t o0 = (uint64_t)tile * kMergeTile;
    const uint32_t count = (uint32_t)min((uint64_t)kMergeTile, (uint64_t)n_a + n_b - o0);
    const uint32_t a0 = part[tile], a1 = part[tile + 1];
    const uint32_t b0 = (uint32_t)(o0 - a0);
    const uint32_t ca = a1 - a0, cb = count - ca;   // the tile's A keys sit at s_key[0, ca), its B keys behind them
    for (uint32_t i = threadIdx.x; i < count; i += kThreads) s_key[i] = i < ca ? a[a0 + i] : b[b0 + (i - ca)];
    __syncthreads();
    const uint32_t diag = min(threadIdx.x * kMergeItems, count);
    uint32_t ia = merge_path(diag, ca, cb, [&](uint32_t i) { return read_of(s_key[i]); },
                             [&](uint32_t j) { return read_of(s_key[ca + j]); });
    uint32_t ib = diag - ia;
    uint64_t res[kMergeItems];
#pragma unroll
    for (int j = 0; j < kMergeItems; ++j) {
        const bool has_a = ia < ca, has_b = ib < cb;
        const uint64_t ka = has_a ? s_key[ia] : 0, kb = has_b ? s_key[ca + ib] : 0;
        const bool take_a = has_a && (!has_b || read_of(ka) <= read_of(kb));
        res[j] = take_a ? ka : kb;
        if (take_a) ++ia; else ++ib;
    }
#pragma unroll
    for (int j = 0; j < kMergeItems; ++j)
        if (diag + j < count) out[o0 + diag + j] = res[j];
}

// hits grouped by read (unitigs in any order inside a read): 1 for the first occurrence of a (read, unitig).
// A read longer than kMaxReadScan hits is left to the general sort path (*too_long is raised).
constexpr uint32_t kMaxReadScan = 64;
struct FirstInReadFlag {
    const uint64_t *hits;
    uint32_t *too_long;
    __device__ uint32_t operator()(uint64_t i) const {
        const uint64_t k = hits[i];
        for (uint32_t back = 1; back <= i; ++back) {
            const uint64_t p = hits[i - back];
            if ((p >> 32) != (k >> 32)) return 1u;
            if (p == k) return 0u;
            if (back == kMaxReadScan) { atomicExch(too_long, 1u); return 1u; }
        }
        return 1u;
    }
};

// canonical undirected key (min << 32 | max); loops keep u == v and are dropped
// by the unique pass.
__global__ void __launch_bounds__(kThreads) pack_pairs_kernel(const uint32_t *__restrict__ u, const uint32_t *__restrict__ v,
                                                              uint64_t n_pairs, uint32_t n_vertices,
                                                              uint64_t *__restrict__ keys, uint32_t *__restrict__ info) {
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t a = u[i], b = v[i];
        bad |= (a >= n_vertices) | (b >= n_vertices);
        uint32_t lo = min(a, b), hi = max(a, b);
        keys[i] = ((uint64_t)lo << 32) | hi;
    }
    if (__any_sync(kFullMask, bad) && lane_id() == 0) atomicOr(&info[1], 1u);
}

// ---- per-read segments -------------------------------------------------------

// A "good" segment is a read with >= 2 distinct unitigs.  hi word counts good
// heads, lo word counts good tails; the s-th good head and the s-th good tail
// delimit the s-th good segment.
struct SegFlagIn {
    const uint64_t *hits;  // sorted unique (read << 32 | unitig)
    uint64_t n;
    __device__ uint64_t operator()(uint64_t i) const {
        uint32_t r = (uint32_t)(hits[i] >> 32);
        bool same_prev = i > 0 && (uint32_t)(hits[i - 1] >> 32) == r;
        bool same_next = i + 1 < n && (uint32_t)(hits[i + 1] >> 32) == r;
        uint64_t head = (!same_prev && same_next) ? 1ull : 0ull;
        uint64_t tail = (same_prev && !same_next) ? 1ull : 0ull;
        return (head << 32) | tail;
    }
};
struct SegFlagOut {
    uint32_t *seg_head;
    uint32_t *seg_tail;
    __device__ void operator()(uint64_t i, uint64_t prefix, uint64_t v) const {
        if (v >> 32) seg_head[prefix >> 32] = (uint32_t)i;
        if (v & 0xffffffffu) seg_tail[prefix & 0xffffffffu] = (uint32_t)i;
    }
};

// pairs of segment s = C(size, 2)
struct SegPairsIn {
    const uint32_t *seg_head;
    const uint32_t *seg_tail;
    __device__ uint64_t operator()(uint64_t s) const {
        uint64_t k = (uint64_t)(seg_tail[s] - seg_head[s]) + 1;
        return k * (k - 1) / 2;
    }
};
struct SegPairsOut {
    uint64_t *seg_base;
    __device__ void operator()(uint64_t s, uint64_t prefix, uint64_t) const { seg_base[s] = prefix; }
};

// ---- load-balanced pair emission ----------------------------------------------

constexpr int kEmitItems = 8;
constexpr int kEmitTile = kThreads * kEmitItems;  // outputs per CTA

// tile_seg[t] = last segment s with seg_base[s] <= t * kEmitTile
__global__ void __launch_bounds__(kThreads) emit_partition_kernel(const uint64_t *__restrict__ seg_base, uint32_t n_seg,
                                                                  uint32_t n_tiles, uint32_t *__restrict__ tile_seg) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    uint64_t p = (uint64_t)t * kEmitTile;
    uint32_t lo = 0, hi = n_seg;  // invariant: seg_base[lo] <= p, (hi == n_seg or seg_base[hi] > p)
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (seg_base[mid] <= p) lo = mid; else hi = mid;
    }
    tile_seg[t] = lo;
}

__global__ void __launch_bounds__(kThreads) emit_pairs_kernel(const uint64_t *__restrict__ hits,
                                                              const uint32_t *__restrict__ seg_head,
                                                              const uint64_t *__restrict__ seg_base, uint32_t n_seg,
                                                              const uint32_t *__restrict__ tile_seg, uint32_t n_tiles,
                                                              uint64_t n_pairs, uint64_t *__restrict__ pairs) {
    __shared__ uint64_t s_base[kEmitTile + 2];
    __shared__ uint32_t s_head[kEmitTile + 2];
    const uint32_t tile = blockIdx.x;
    const uint64_t p0 = (uint64_t)tile * kEmitTile;
    const uint64_t p1 = min(n_pairs, p0 + kEmitTile);
    const uint32_t s_first = tile_seg[tile];
    const uint32_t s_last = tile + 1 < n_tiles ? tile_seg[tile + 1] : n_seg - 1;
    const uint32_t cnt = s_last - s_first + 1;  // <= kEmitTile + 1: every segment is non-empty
    for (uint32_t i = threadIdx.x; i < cnt; i += kThreads) {
        s_base[i] = seg_base[s_first + i];
        s_head[i] = seg_head[s_first + i];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kEmitItems; ++j) {
        uint64_t p = p0 + (uint64_t)j * kThreads + threadIdx.x;
        if (p >= p1) break;
        uint32_t lo = 0, hi = cnt;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_base[mid] <= p) lo = mid; else hi = mid;
        }
        // t-th pair (a < b) of the segment in colexicographic order: t = b(b-1)/2 + a
        uint64_t t = p - s_base[lo];
        uint64_t b = (uint64_t)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
        while (b * (b - 1) / 2 > t) --b;
        while ((b + 1) * b / 2 <= t) ++b;
        uint64_t a = t - b * (b - 1) / 2;
        const uint64_t h = s_head[lo];
        const uint32_t ua = (uint32_t)hits[h + a], ub = (uint32_t)hits[h + b];  // distinct; ascending only after a sort
        pairs[p] = ((uint64_t)min(ua, ub) << 32) | max(ua, ub);
    }
}

// ---- unique of sorted pair keys, dropping loops ---------------------------------

struct EdgeFlagIn {
    const uint64_t *keys;
    __device__ uint32_t operator()(uint64_t i) const {
        uint64_t k = keys[i];
        bool head = (i == 0 || keys[i - 1] != k);
        bool loop = (uint32_t)(k >> 32) == (uint32_t)k;
        return (head && !loop) ? 1u : 0u;
    }
};

// adjacent-unique of the sorted pair keys that also records how many pairs collapsed into each edge (the number of
// reads supporting it: the "weight" of the north star's weighted graph; the reference itself keeps no weights,
// src/graph.cpp:438 passes comb = NULL).  Runs are short (1.0005 pairs per edge on cfg2), so the head looks ahead a
// few keys and only then falls back to a binary search for the end of its run.
struct CompactEdgesMult {
    const uint64_t *keys;
    uint64_t n;
    uint64_t *dst;
    uint32_t *mult;
    __device__ void operator()(uint64_t i, uint32_t pos, uint32_t flag) const {
        if (!flag) return;
        const uint64_t k = keys[i];
        dst[pos] = k;
        uint64_t r = 1;
        while (r < 8 && i + r < n && keys[i + r] == k) ++r;
        if (r == 8 && i + r < n && keys[i + r] == k) {
            uint64_t lo = i + r, hi = n;   // first position past i + r whose key differs
            while (lo < hi) {
                const uint64_t mid = lo + (hi - lo) / 2;
                if (keys[mid] == k) lo = mid + 1; else hi = mid;
            }
            r = lo - i;
        }
        mult[pos] = (uint32_t)r;
    }
};

// ---- CSR ---------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads) swap_pack_kernel(const uint64_t *__restrict__ edges, uint64_t n_edges,
                                                             uint64_t *__restrict__ swapped) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t e = edges[i];
        swapped[i] = (e << 32) | (e >> 32);
    }
}

// keys sorted by their high word, which lies in [base, base + n_rows): start[x] = first index whose high word is
// >= base + x, for x in [0, n_rows]  (start[n_rows] = count).  No atomics on the result: thread i fills the rows between
// its predecessor's and its own -- but only short gaps; a long run of empty rows (a rank of the partitioned path sees
// only the neighbourhood of its own id range; an edge list may leave whole id ranges without edges) is recorded and
// filled by a whole CTA in a second kernel, so one thread never writes millions of entries.
// *err (optional) is raised when a key lies outside [base, base + n_rows) or its low word is >= lo_bound.
constexpr int64_t kGapShort = 32;
struct RowGap {
    uint32_t lo, hi, val;   // start[lo .. hi] = val
};
__global__ void __launch_bounds__(kThreads) row_bounds_kernel(const uint64_t *__restrict__ keys, uint64_t count, uint32_t base,
                                                              uint32_t n_rows, uint32_t lo_bound, uint32_t *__restrict__ start,
                                                              RowGap *__restrict__ gaps, uint32_t *__restrict__ n_gaps, uint32_t *__restrict__ err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= count; i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t cur = (int64_t)n_rows, prev = -1;
        if (i < count) {
            const uint64_t k = keys[i];
            cur = (int64_t)(uint32_t)(k >> 32) - (int64_t)base;
            if (cur < 0 || cur >= (int64_t)n_rows || (uint32_t)k >= lo_bound) {
                if (err) atomicExch(err, 2u);
                cur = (int64_t)n_rows;
            }
        }
        if (i > 0) {
            prev = (int64_t)(uint32_t)(keys[i - 1] >> 32) - (int64_t)base;
            if (prev < 0 || prev >= (int64_t)n_rows) prev = (int64_t)n_rows;
        }
        if (cur - prev > kGapShort) {
            const uint32_t g = atomicAdd(n_gaps, 1u);
            gaps[g] = RowGap{(uint32_t)(prev + 1), (uint32_t)cur, (uint32_t)i};
        } else {
            for (int64_t x = prev + 1; x <= cur; ++x) start[x] = (uint32_t)i;
        }
    }
}
__global__ void __launch_bounds__(kThreads) row_gaps_kernel(const RowGap *__restrict__ gaps, const uint32_t *__restrict__ n_gaps,
                                                            uint32_t *__restrict__ start) {
    const uint32_t n = *n_gaps;
    for (uint32_t g = blockIdx.x; g < n; g += gridDim.x) {
        const RowGap gp = gaps[g];
        for (uint64_t x = (uint64_t)gp.lo + threadIdx.x; x <= gp.hi; x += blockDim.x) start[x] = gp.val;
    }
}

struct DegreeIn {
    const uint32_t *fwd_start;
    const uint32_t *back_start;
    __device__ uint64_t operator()(uint64_t v) const {
        return (uint64_t)(fwd_start[v + 1] - fwd_start[v]) + (uint64_t)(back_start[v + 1] - back_start[v]);
    }
};
struct DegreeOut {
    uint64_t *row_ptr;
    int32_t *deg;
    __device__ void operator()(uint64_t v, uint64_t prefix, uint64_t d) const {
        row_ptr[v] = prefix;
        deg[v] = (int32_t)d;
    }
};

// max of an int32 array: warp shuffle -> one atomic per warp
__global__ void __launch_bounds__(kThreads) reduce_max_i32_kernel(const int32_t *__restrict__ x, uint64_t n,
                                                                  int32_t *__restrict__ out) {
    int32_t m = INT32_MIN;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, x[i]);
    m = warp_reduce_max(m);
    if (lane_id() == 0 && m != INT32_MIN) atomicMax(out, m);
}

// Row v = [back neighbours (< v, ascending) | forward neighbours (> v, ascending)].
__global__ void __launch_bounds__(kThreads) fill_col_kernel(const uint64_t *__restrict__ edges,
                                                            const uint64_t *__restrict__ swapped, uint64_t n_edges,
                                                            const uint64_t *__restrict__ row_ptr,
                                                            const uint32_t *__restrict__ fwd_start,
                                                            const uint32_t *__restrict__ back_start,
                                                            uint32_t *__restrict__ col) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t e = edges[i];
        uint32_t u = (uint32_t)(e >> 32);
        uint64_t nback = back_start[u + 1] - back_start[u];
        col[row_ptr[u] + nback + (i - fwd_start[u])] = (uint32_t)e;
        uint64_t s = swapped[i];
        uint32_t v = (uint32_t)(s >> 32);
        col[row_ptr[v] + (i - back_start[v])] = (uint32_t)s;
    }
}

__global__ void set_u64_kernel(uint64_t *p, uint64_t v) { *p = v; }

// keys sorted by high word: idx[j] = first position whose high word is >= bounds[j]
__global__ void lower_bounds_kernel(const uint64_t *__restrict__ keys, uint64_t count, const uint32_t *__restrict__ bounds,
                                    int nb, uint64_t *__restrict__ idx) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    const uint64_t target = (uint64_t)bounds[j] << 32;
    uint64_t lo = 0, hi = count;  // first position with keys[pos] >= target
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    idx[j] = lo;
}

// row r of a partition owns the keys whose high word is base + r
// *err is raised when a key's row lies outside [base, base + n_rows) or its low word is >= lo_bound (an entry routed to
// the wrong rank, or a corrupt one, must not write past the arrays)
__global__ void __launch_bounds__(kThreads) row_bounds_base_kernel(const uint64_t *__restrict__ keys, uint64_t count,
                                                                   uint32_t base, uint32_t n_rows, uint32_t lo_bound,
                                                                   uint64_t *__restrict__ start, uint32_t *__restrict__ err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= count; i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t cur = (int64_t)n_rows, prev = -1;
        if (i < count) {
            const uint64_t k = keys[i];
            cur = (int64_t)(uint32_t)(k >> 32) - (int64_t)base;
            if (cur < 0 || cur >= (int64_t)n_rows || (uint32_t)k >= lo_bound) { atomicExch(err, 2u); cur = (int64_t)n_rows; }
        }
        if (i > 0) {
            prev = (int64_t)(uint32_t)(keys[i - 1] >> 32) - (int64_t)base;
            if (prev < 0 || prev >= (int64_t)n_rows) prev = (int64_t)n_rows;
        }
        for (int64_t x = prev + 1; x <= cur; ++x) start[x] = i;
    }
}

__global__ void __launch_bounds__(kThreads) part_cols_kernel(const uint64_t *__restrict__ keys, uint64_t count,
                                                             uint32_t *__restrict__ col) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        col[i] = (uint32_t)keys[i];
}

__global__ void __launch_bounds__(kThreads) part_degree_kernel(const uint64_t *__restrict__ row_ptr, uint32_t n_rows,
                                                               int32_t *__restrict__ deg, int32_t *__restrict__ max_deg) {
    int32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (uint64_t)gridDim.x * blockDim.x) {
        int32_t d = (int32_t)(row_ptr[i + 1] - row_ptr[i]);
        deg[i] = d;
        m = max(m, d);
    }
    m = warp_reduce_max(m);
    if (lane_id() == 0 && m) atomicMax(max_deg, m);
}

struct AnyHeadFlag {  // unique of sorted directed entries (no loop test: routed entries never hold loops)
    const uint64_t *keys;
    __device__ uint32_t operator()(uint64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
};

// hits -> all clique pairs, sorted (duplicates still in).  pa / pb are the ping-pong buffers.
int hits_to_sorted_pairs(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits,
                         uint32_t n_vertices, DevBuf<uint64_t> &pa, DevBuf<uint64_t> &pb, uint64_t **psorted_out,
                         uint64_t *n_pairs_out, kombgpu_stats *st, bool sort_pairs = true) {
    if (n_hits >= (1ull << 32)) return ctx_fail(ctx, KOMBGPU_EINVAL, "n_hits >= 2^32 is not supported on one device");
    const int bn = bits_for(n_vertices > 0 ? n_vertices - 1 : 0);
    st->n_hits = n_hits;

    // 1. (read, unitig) keys, sorted + unique  == per-read unitig SETS, mates merged
    DevBuf<uint64_t> ha, hb;
    DevBuf<uint32_t> info(ctx, 5);
    KG_ALLOC(ctx, ha, n_hits);
    KG_ALLOC(ctx, hb, n_hits);
    if (!info) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    const uint32_t info0[5] = {0, 0, 0, 0xffffffffu, 0};   // max key, error, descents, first descent, read too long
    KG_CUDA(ctx, cudaMemcpyAsync(info.p, info0, sizeof(info0), cudaMemcpyHostToDevice, ctx->stream));
    if (n_hits)
        KG_LAUNCH(ctx, pack_hits_kernel, min(grid_for(n_hits, kThreads), 148u * 16u), kThreads, 0, read_key, unitig, n_hits,
                  n_vertices, ha.p, info.p);
    uint32_t h_info[5] = {0, 0, 0, 0, 0};
    KG_TRY(read_back(ctx, info.p, h_info, 5));
    if (h_info[1]) return ctx_fail(ctx, KOMBGPU_EINVAL, "unitig id >= n_vertices (%u)", n_vertices);
    DevBuf<uint32_t> d_cnt(ctx, 1);
    if (!d_cnt) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    uint64_t *sorted = ha.p, *other = hb.p;
    uint32_t n_uniq = 0;
    RadixPass passes[8];
    bool grouped = false;   // unique hits grouped by read without a sort?
    DevBuf<uint64_t> hc;    // third buffer of the merge path (see below)
    if (n_hits && h_info[2] <= 1 && !getenv("KOMBGPU_NO_MERGE")) {
        // The reads arrive in file order (at most two runs of non-decreasing keys: the two mate files): a stable
        // merge by read key replaces the radix sort (6 passes for cfg2), and since the per-read unitig SET is all
        // the path needs, duplicates inside a read are found by looking back through the read.
        uint64_t *by_read = ha.p;
        if (h_info[2] == 1) {
            const uint32_t n_a = h_info[3], n_b = (uint32_t)(n_hits - n_a);
            const uint32_t n_tiles = ceil_div_u64(n_hits, kMergeTile);
            DevBuf<uint32_t> part;
            KG_ALLOC(ctx, part, (size_t)n_tiles + 1);
            KG_LAUNCH(ctx, merge_partition_kernel, grid_for((uint64_t)n_tiles + 1, kThreads), kThreads, 0, ha.p, n_a, ha.p + n_a, n_b,
                      n_tiles, part.p);
            KG_LAUNCH(ctx, merge_runs_kernel, n_tiles, kThreads, 0, ha.p, n_a, ha.p + n_a, n_b, part.p, hb.p);
            by_read = hb.p;
        }
        // compaction needs a third buffer only when the merge used both: the packed input is dead after the merge,
        // but a read that is too long sends us back to it, so keep it and compact into scratch
        uint64_t *dst = by_read == ha.p ? hb.p : nullptr;
        if (!dst) { KG_ALLOC(ctx, hc, n_hits); dst = hc.p; }
        KG_TRY((device_scan<uint32_t>(ctx, n_hits, FirstInReadFlag{by_read, info.p + 4}, CompactKeysU64{by_read, dst}, d_cnt.p)));
        uint32_t too_long = 0;
        KG_TRY(read_back(ctx, info.p + 4, &too_long, 1));
        if (!too_long) {
            KG_TRY(read_back(ctx, d_cnt.p, &n_uniq, 1));
            grouped = true;
            other = dst;
        }
    }
    if (!grouped) {
        int np = plan_radix_passes(0, bn, 32, 32 + bits_for(h_info[0]), passes);
        KG_TRY(radix_sort_u64(ctx, ha.p, hb.p, n_hits, passes, np, &sorted));
        other = sorted == ha.p ? hb.p : ha.p;
        KG_TRY((device_scan<uint32_t>(ctx, n_hits, HeadFlagU64{sorted}, CompactKeysU64{sorted, other}, d_cnt.p)));
        KG_TRY(read_back(ctx, d_cnt.p, &n_uniq, 1));
    }
    const uint64_t *hits = other;  // unique hits grouped by read (sorted when the general path ran), n_uniq of them
    st->n_unique_hits = n_uniq;
    // the segment scan below carries two 32-bit counters in one word whose top two bits are the scan's status flags
    if (n_uniq >= (1u << 31)) return ctx_fail(ctx, KOMBGPU_EINVAL, "%u distinct (read, unitig) hits exceed the 2^31 per-device limit", n_uniq);

    // 2. reads with >= 2 unitigs -> segments; pairs per segment -> offsets
    DevBuf<uint32_t> seg_head, seg_tail;
    DevBuf<uint64_t> d_tot(ctx, 1);
    KG_ALLOC(ctx, seg_head, (size_t)n_uniq / 2 + 1);
    KG_ALLOC(ctx, seg_tail, (size_t)n_uniq / 2 + 1);
    if (!d_tot) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_TRY((device_scan<uint64_t>(ctx, n_uniq, SegFlagIn{hits, n_uniq}, SegFlagOut{seg_head.p, seg_tail.p}, d_tot.p)));
    uint64_t seg_tot = 0;
    KG_TRY(read_back(ctx, d_tot.p, &seg_tot, 1));
    const uint32_t n_seg = (uint32_t)(seg_tot >> 32);
    DevBuf<uint64_t> seg_base;
    KG_ALLOC(ctx, seg_base, (size_t)n_seg + 1);
    KG_TRY((device_scan<uint64_t>(ctx, n_seg, SegPairsIn{seg_head.p, seg_tail.p}, SegPairsOut{seg_base.p}, d_tot.p)));
    uint64_t n_pairs = 0;
    KG_TRY(read_back(ctx, d_tot.p, &n_pairs, 1));
    st->n_pairs = n_pairs;
    *n_pairs_out = n_pairs;
    if (n_pairs >= (1ull << 32))
        return ctx_fail(ctx, KOMBGPU_EINVAL, "%llu clique pairs exceed the 2^32 per-device limit", (unsigned long long)n_pairs);

    // 3. emit all pairs, sort
    KG_ALLOC(ctx, pa, n_pairs);
    if (sort_pairs) KG_ALLOC(ctx, pb, n_pairs);
    uint64_t *psorted = pa.p;
    if (n_pairs) {
        const uint32_t n_tiles = ceil_div_u64(n_pairs, kEmitTile);
        DevBuf<uint32_t> tile_seg;
        KG_ALLOC(ctx, tile_seg, n_tiles);
        KG_LAUNCH(ctx, emit_partition_kernel, grid_for(n_tiles, kThreads), kThreads, 0, seg_base.p, n_seg, n_tiles, tile_seg.p);
        KG_LAUNCH(ctx, emit_pairs_kernel, n_tiles, kThreads, 0, hits, seg_head.p, seg_base.p, n_seg, tile_seg.p, n_tiles,
                  n_pairs, pa.p);
        if (sort_pairs) {
            int npp = plan_radix_passes(0, bn, 32, 32 + bn, passes);
            KG_TRY(radix_sort_u64(ctx, pa.p, pb.p, n_pairs, passes, npp, &psorted));
        }
    }
    *psorted_out = psorted;
    return KOMBGPU_OK;
}

// (u, v) pairs -> canonical keys, sorted (duplicates and loops still in)
int pairs_to_sorted_keys(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices,
                         DevBuf<uint64_t> &ka, DevBuf<uint64_t> &kb, uint64_t **sorted_out) {
    if (n_pairs >= (1ull << 32)) return ctx_fail(ctx, KOMBGPU_EINVAL, "n_pairs >= 2^32 is not supported on one device");
    const int bn = bits_for(n_vertices > 0 ? n_vertices - 1 : 0);
    DevBuf<uint32_t> info(ctx, 2);
    KG_ALLOC(ctx, ka, n_pairs);
    KG_ALLOC(ctx, kb, n_pairs);
    if (!info) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(info.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    uint64_t *sorted = ka.p;
    if (n_pairs) {
        KG_LAUNCH(ctx, pack_pairs_kernel, min(grid_for(n_pairs, kThreads), 148u * 16u), kThreads, 0, u, v, n_pairs, n_vertices,
                  ka.p, info.p);
        RadixPass passes[8];
        int np = plan_radix_passes(0, bn, 32, 32 + bn, passes);
        KG_TRY(radix_sort_u64(ctx, ka.p, kb.p, n_pairs, passes, np, &sorted));
    }
    uint32_t h_info[2] = {0, 0};
    KG_TRY(read_back(ctx, info.p, h_info, 2));
    if (h_info[1]) return ctx_fail(ctx, KOMBGPU_EINVAL, "edge endpoint >= n_vertices (%u)", n_vertices);
    *sorted_out = sorted;
    return KOMBGPU_OK;
}

}  // namespace

// sorted pair keys (duplicates, loops possible) -> unique simple edges
int unique_edges(kombgpu_ctx *ctx, const uint64_t *keys, uint64_t count, DevBuf<uint64_t> &edges, uint64_t *n_edges,
                 DevBuf<uint32_t> *mult) {
    DevBuf<uint32_t> d_count(ctx, 1);
    if (!d_count) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_ALLOC(ctx, edges, count);
    if (mult) {
        KG_ALLOC(ctx, *mult, count);
        KG_TRY((device_scan<uint32_t>(ctx, count, EdgeFlagIn{keys}, CompactEdgesMult{keys, count, edges.p, mult->p}, d_count.p)));
    } else
    KG_TRY((device_scan<uint32_t>(ctx, count, EdgeFlagIn{keys}, CompactKeysU64{keys, edges.p}, d_count.p)));
    uint32_t n32 = 0;
    KG_TRY(read_back(ctx, d_count.p, &n32, 1));
    *n_edges = n32;
    return KOMBGPU_OK;
}

// (v << 32 | u) copy of an edge list, stably sorted on v: the backward half of the adjacency in order
int swapped_sorted(kombgpu_ctx *ctx, const uint64_t *edges, uint64_t n_edges, uint32_t n_vertices, DevBuf<uint64_t> &a,
                   DevBuf<uint64_t> &b, uint64_t **out) {
    const int bn = bits_for(n_vertices > 0 ? n_vertices - 1 : 0);
    KG_ALLOC(ctx, a, n_edges);
    KG_ALLOC(ctx, b, n_edges);
    *out = a.p;
    if (n_edges) {
        KG_LAUNCH(ctx, swap_pack_kernel, min(grid_for(n_edges, kThreads), 148u * 16u), kThreads, 0, edges, n_edges, a.p);
        RadixPass passes[8];
        int np = plan_radix_passes(32, 32 + bn, 0, 0, passes);
        KG_TRY(radix_sort_u64(ctx, a.p, b.p, n_edges, passes, np, out));
    }
    return KOMBGPU_OK;
}

int lower_bounds_hi(kombgpu_ctx *ctx, const uint64_t *keys, uint64_t count, const uint32_t *bounds_host, int nb,
                    uint64_t *idx_host) {
    DevBuf<uint32_t> d_b;
    DevBuf<uint64_t> d_i;
    KG_ALLOC(ctx, d_b, nb);
    KG_ALLOC(ctx, d_i, nb);
    KG_CUDA(ctx, cudaMemcpyAsync(d_b.p, bounds_host, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    KG_LAUNCH(ctx, lower_bounds_kernel, 1, 64, 0, keys, count, d_b.p, nb, d_i.p);
    return read_back(ctx, d_i.p, idx_host, nb);
}

int hits_to_edges(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits, uint32_t n_vertices,
                  DevBuf<uint64_t> &edges, uint64_t *n_edges, kombgpu_stats *st, DevBuf<uint32_t> *mult) {
    DevBuf<uint64_t> pa, pb;
    uint64_t *psorted = nullptr, n_pairs = 0;
    KG_TRY(hits_to_sorted_pairs(ctx, read_key, unitig, n_hits, n_vertices, pa, pb, &psorted, &n_pairs, st));
    return unique_edges(ctx, psorted, n_pairs, edges, n_edges, mult);
}

int pairs_to_edges(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices,
                   DevBuf<uint64_t> &edges, uint64_t *n_edges, DevBuf<uint32_t> *mult) {
    DevBuf<uint64_t> ka, kb;
    uint64_t *sorted = nullptr;
    KG_TRY(pairs_to_sorted_keys(ctx, u, v, n_pairs, n_vertices, ka, kb, &sorted));
    return unique_edges(ctx, sorted, n_pairs, edges, n_edges, mult);
}

// hits -> every clique pair as a canonical key (min << 32 | max), in emission order (not sorted, duplicates in)
int hits_to_pairs(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits, uint32_t n_vertices,
                  DevBuf<uint64_t> &pairs, uint64_t *n_pairs, kombgpu_stats *st) {
    DevBuf<uint64_t> unused;
    uint64_t *p = nullptr;
    return hits_to_sorted_pairs(ctx, read_key, unitig, n_hits, n_vertices, pairs, unused, &p, n_pairs, st, false);
}

// (u, v) pairs -> canonical keys (min << 32 | max), input order; ids are range-checked
int pairs_to_keys(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices, DevBuf<uint64_t> &keys) {
    if (n_pairs >= (1ull << 32)) return ctx_fail(ctx, KOMBGPU_EINVAL, "n_pairs >= 2^32 is not supported on one device");
    DevBuf<uint32_t> info(ctx, 2);
    KG_ALLOC(ctx, keys, n_pairs);
    if (!info) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(info.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    if (n_pairs)
        KG_LAUNCH(ctx, pack_pairs_kernel, min(grid_for(n_pairs, kThreads), 148u * 16u), kThreads, 0, u, v, n_pairs, n_vertices, keys.p, info.p);
    uint32_t h_info[2] = {0, 0};
    KG_TRY(read_back(ctx, info.p, h_info, 2));
    if (h_info[1]) return ctx_fail(ctx, KOMBGPU_EINVAL, "edge endpoint >= n_vertices (%u)", n_vertices);
    return KOMBGPU_OK;
}

int row_starts(kombgpu_ctx *ctx, const uint64_t *keys, uint64_t count, uint32_t base, uint32_t n_rows, uint32_t lo_bound, uint32_t *start,
               uint32_t *err_dev) {
    // a gap longer than kGapShort rows is recorded: at most n_rows / kGapShort + 1 of them
    DevBuf<RowGap> gaps;
    DevBuf<uint32_t> n_gaps(ctx, 1);
    KG_ALLOC(ctx, gaps, (size_t)n_rows / (size_t)kGapShort + 2);
    if (!n_gaps) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(n_gaps.p, 0, sizeof(uint32_t), ctx->stream));
    KG_LAUNCH(ctx, row_bounds_kernel, min(grid_for(count + 1, kThreads), 148u * 16u), kThreads, 0, keys, count, base, n_rows, lo_bound, start,
              gaps.p, n_gaps.p, err_dev);
    KG_LAUNCH(ctx, row_gaps_kernel, 148u * 4u, kThreads, 0, gaps.p, n_gaps.p, start);
    return KOMBGPU_OK;
}

int forward_index(kombgpu_ctx *ctx, const uint64_t *edges, uint64_t E, uint32_t n, uint32_t **fwd_start_out) {
    DevBuf<uint32_t> fwd_start;
    KG_ALLOC(ctx, fwd_start, (size_t)n + 1);
    KG_TRY(row_starts(ctx, edges, E, 0u, n, 0xffffffffu, fwd_start.p, nullptr));
    *fwd_start_out = fwd_start.take();
    return KOMBGPU_OK;
}

// unique simple edges (sorted, u < v) -> CSR of the symmetric graph
int csr_from_edges(kombgpu_ctx *ctx, DevBuf<uint64_t> &edges, uint64_t E, uint32_t n, kombgpu_graph *g) {
    g->n = n;
    g->n_edges = E;
    g->st.n_edges = E;
    g->st.n_vertices = n;
    DevBuf<uint64_t> sw_a, sw_b;
    uint64_t *swapped = nullptr;
    KG_TRY(swapped_sorted(ctx, edges.p, E, n, sw_a, sw_b, &swapped));
    DevBuf<uint32_t> back_start;
    KG_ALLOC(ctx, back_start, (size_t)n + 1);
    if (!g->fwd_start) KG_TRY(forward_index(ctx, edges.p, E, n, &g->fwd_start));   // kept: the CSR form of the edge list
    const uint32_t *fwd_start_p = g->fwd_start;
    KG_TRY(row_starts(ctx, swapped, E, 0u, n, 0xffffffffu, back_start.p, nullptr));

    DevBuf<uint64_t> row_ptr;
    DevBuf<int32_t> deg, max_deg(ctx, 1);
    DevBuf<uint32_t> col;
    KG_ALLOC(ctx, row_ptr, (size_t)n + 1);
    KG_ALLOC(ctx, deg, n);
    KG_ALLOC(ctx, col, 2 * E);
    if (!max_deg) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(max_deg.p, 0, sizeof(int32_t), ctx->stream));
    KG_TRY((device_scan<uint64_t>(ctx, n, DegreeIn{fwd_start_p, back_start.p}, DegreeOut{row_ptr.p, deg.p},
                                  (uint64_t *)nullptr)));
    if (n) KG_LAUNCH(ctx, reduce_max_i32_kernel, min(grid_for(n, kThreads), 148u * 8u), kThreads, 0, deg.p, (uint64_t)n, max_deg.p);
    KG_LAUNCH(ctx, set_u64_kernel, 1, 1, 0, row_ptr.p + n, 2 * E);
    if (E)
        KG_LAUNCH(ctx, fill_col_kernel, min(grid_for(E, kThreads), 148u * 16u), kThreads, 0, edges.p, swapped, E, row_ptr.p,
                  fwd_start_p, back_start.p, col.p);
    KG_TRY(read_back(ctx, max_deg.p, &g->st.max_degree, 1));
    g->edges = edges.take();
    g->row_ptr = row_ptr.take();
    g->col = col.take();
    g->deg = deg.take();
    return KOMBGPU_OK;
}

// directed entries (src << 32 | dst) with src in [v_lo, v_lo + n_local), any order, duplicates allowed
// -> CSR rows of this rank: row_ptr[n_local + 1], col[] = global ids (every row ascending), deg[]
int csr_from_directed(kombgpu_ctx *ctx, const uint64_t *entries, uint64_t count, uint32_t v_lo, uint32_t n_local,
                      uint32_t n_global, uint64_t **row_ptr_out, uint32_t **col_out, int32_t **deg_out, int32_t *max_deg_out,
                      uint64_t *n_directed_out) {
    if (count >= (1ull << 32)) return ctx_fail(ctx, KOMBGPU_EINVAL, "more than 2^32 directed entries on one device");
    const int bn = bits_for(n_global > 0 ? n_global - 1 : 0);
    DevBuf<uint64_t> ka, kb, uniq;
    KG_ALLOC(ctx, ka, count);
    KG_ALLOC(ctx, kb, count);
    uint64_t *sorted = ka.p;
    if (count) {
        KG_CUDA(ctx, cudaMemcpyAsync(ka.p, entries, count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        RadixPass passes[8];
        int np = plan_radix_passes(0, bn, 32, 32 + bn, passes);
        KG_TRY(radix_sort_u64(ctx, ka.p, kb.p, count, passes, np, &sorted));
    }
    DevBuf<uint32_t> d_count(ctx, 1);
    if (!d_count) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_ALLOC(ctx, uniq, count);
    KG_TRY((device_scan<uint32_t>(ctx, count, AnyHeadFlag{sorted}, CompactKeysU64{sorted, uniq.p}, d_count.p)));
    uint32_t nd32 = 0;
    KG_TRY(read_back(ctx, d_count.p, &nd32, 1));
    const uint64_t nd = nd32;
    DevBuf<uint64_t> row_ptr;
    DevBuf<uint32_t> col;
    DevBuf<int32_t> deg, max_deg(ctx, 1);
    KG_ALLOC(ctx, row_ptr, (size_t)n_local + 1);
    KG_ALLOC(ctx, col, nd);
    KG_ALLOC(ctx, deg, n_local);
    if (!max_deg) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(max_deg.p, 0, sizeof(int32_t), ctx->stream));
    DevBuf<uint32_t> d_err(ctx, 1);
    if (!d_err) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(uint32_t), ctx->stream));
    KG_LAUNCH(ctx, row_bounds_base_kernel, min(grid_for(nd + 1, kThreads), 148u * 16u), kThreads, 0, uniq.p, nd, v_lo, n_local,
              n_global ? n_global : 1u, row_ptr.p, d_err.p);
    uint32_t h_err = 0;
    KG_TRY(read_back(ctx, d_err.p, &h_err, 1));
    if (h_err) return ctx_fail(ctx, KOMBGPU_EINVAL, "a directed entry lies outside this rank's rows [%u, %u) or names a unitig >= %u", v_lo, v_lo + n_local, n_global);
    if (nd) KG_LAUNCH(ctx, part_cols_kernel, min(grid_for(nd, kThreads), 148u * 16u), kThreads, 0, uniq.p, nd, col.p);
    if (n_local)
        KG_LAUNCH(ctx, part_degree_kernel, min(grid_for(n_local, kThreads), 148u * 8u), kThreads, 0, row_ptr.p, n_local, deg.p,
                  max_deg.p);
    KG_TRY(read_back(ctx, max_deg.p, max_deg_out, 1));
    *n_directed_out = nd;
    *row_ptr_out = row_ptr.take();
    *col_out = col.take();
    *deg_out = deg.take();
    return KOMBGPU_OK;
}

// copy an existing CSR (device) into a graph; degrees from the row lengths
int adopt_csr(kombgpu_ctx *ctx, const uint64_t *row_ptr, const uint32_t *col, uint32_t n, kombgpu_graph *g) {
    uint64_t n_dir = 0;
    KG_TRY(read_back(ctx, row_ptr + n, &n_dir, 1));
    if (n_dir && !col) return ctx_fail(ctx, KOMBGPU_EINVAL, "null col array");
    DevBuf<uint64_t> rp;
    DevBuf<uint32_t> cl;
    DevBuf<int32_t> deg, max_deg(ctx, 1);
    KG_ALLOC(ctx, rp, (size_t)n + 1);
    KG_ALLOC(ctx, cl, n_dir);
    KG_ALLOC(ctx, deg, n);
    if (!max_deg) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemcpyAsync(rp.p, row_ptr, ((size_t)n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (n_dir) KG_CUDA(ctx, cudaMemcpyAsync(cl.p, col, n_dir * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    KG_CUDA(ctx, cudaMemsetAsync(max_deg.p, 0, sizeof(int32_t), ctx->stream));
    if (n) KG_LAUNCH(ctx, part_degree_kernel, min(grid_for(n, kThreads), 148u * 8u), kThreads, 0, rp.p, n, deg.p, max_deg.p);
    KG_TRY(read_back(ctx, max_deg.p, &g->st.max_degree, 1));
    g->n = n;
    g->n_edges = n_dir / 2;
    g->st.n_vertices = n;
    g->st.n_edges = n_dir / 2;
    g->row_ptr = rp.take();
    g->col = cl.take();
    g->deg = deg.take();
    return KOMBGPU_OK;
}

int build_from_pairs(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices,
                     kombgpu_graph *g) {
    DevBuf<uint64_t> edges;
    uint64_t E = 0;
    g->st.n_pairs = n_pairs;
    DevBuf<uint32_t> mult;
    KG_TRY(pairs_to_edges(ctx, u, v, n_pairs, n_vertices, edges, &E, &mult));
    g->mult = mult.take();
    return csr_from_edges(ctx, edges, E, n_vertices, g);
}

int build_from_hits(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits,
                    uint32_t n_vertices, kombgpu_graph *g) {
    DevBuf<uint64_t> edges;
    uint64_t E = 0;
    DevBuf<uint32_t> mult;
    KG_TRY(hits_to_edges(ctx, read_key, unitig, n_hits, n_vertices, edges, &E, &g->st, &mult));
    g->mult = mult.take();
    return csr_from_edges(ctx, edges, E, n_vertices, g);
}

void graph_release(kombgpu_graph *g) {
    if (!g || !g->ctx) return;
    kombgpu_ctx *ctx = g->ctx;
    truss_release(g);
    format_release(g);
    if (g->edges) ws_free(ctx, g->edges);
    if (g->fwd_start) ws_free(ctx, g->fwd_start);
    if (g->mult) ws_free(ctx, g->mult);
    g->fwd_start = nullptr; g->mult = nullptr;
    if (g->row_ptr) ws_free(ctx, g->row_ptr);
    if (g->col) ws_free(ctx, g->col);
    if (g->deg) ws_free(ctx, g->deg);
    if (g->core) ws_free(ctx, g->core);
    if (g->score) ws_free(ctx, g->score);
    g->edges = nullptr; g->row_ptr = nullptr; g->col = nullptr; g->deg = nullptr; g->core = nullptr; g->score = nullptr;
}

}  // namespace kg
