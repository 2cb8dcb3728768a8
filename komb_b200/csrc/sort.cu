// sort.cu — stable LSD radix sort of 64-bit keys (hits, unitig pairs, CORE-A
// rank keys).  Integer, HBM-bound: per pass the keys are read twice (digit
// histogram, scatter) and written once; the scatter stages a tile in shared
// memory in digit order so global stores are contiguous runs.  A single-read
// "onesweep" variant (decoupled look-back) was measured and was not faster here:
// with ~300 tiles in flight the look-back chains cost what the histogram pass costs.
#include "primitives.cuh"

namespace kg {

namespace {

constexpr int kRsThreads = 512;
constexpr int kRsItems = 8;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per CTA
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRadix = 256;

// table[d * n_tiles + tile] = number of keys of this tile whose digit is d
__global__ void __launch_bounds__(kRsThreads) radix_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift,
                                                                uint32_t mask, uint32_t *__restrict__ table,
                                                                uint32_t n_tiles) {
    __shared__ uint32_t s_hist[kRadix];
    for (int d = threadIdx.x; d < kRadix; d += kRsThreads) s_hist[d] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        uint64_t i = base + (uint64_t)j * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d <= (int)mask; d += kRsThreads) table[(uint64_t)d * n_tiles + blockIdx.x] = s_hist[d];
}

struct TableIn {
    const uint32_t *t;
    __device__ uint32_t operator()(uint64_t i) const { return t[i]; }
};
struct TableOut {
    uint32_t *t;
    __device__ void operator()(uint64_t i, uint32_t prefix, uint32_t) const { t[i] = prefix; }
};

// dynamic shared memory layout of the scatter kernel
struct ScatterSmem {
    uint64_t keys[kRsTile];
    uint32_t warp_cnt[kRsWarps][kRadix];
    uint32_t digit_local[kRadix];   // first slot of digit d inside the staged tile
    uint32_t digit_global[kRadix];  // global position of that slot minus digit_local
    uint32_t scan_tmp[kRadix / 32 + 1];
};

__global__ void __launch_bounds__(kRsThreads) radix_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out,
                                                                   uint64_t n, int shift, uint32_t mask,
                                                                   const uint32_t *__restrict__ table, uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    ScatterSmem &s = *reinterpret_cast<ScatterSmem *>(s_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t tile_base = (uint64_t)blockIdx.x * kRsTile;
    const uint32_t tile_count = (uint32_t)min((uint64_t)kRsTile, n - tile_base);

    for (int i = threadIdx.x; i < kRsWarps * kRadix; i += kRsThreads) (&s.warp_cnt[0][0])[i] = 0;

    // warp-striped load: item j of lane l is tile element warp*256 + j*32 + l,
    // so (warp, j, lane) order is memory order and the sort stays stable.
    uint64_t key[kRsItems];
    uint32_t rank[kRsItems];
    const uint32_t warp_base = warp * (32 * kRsItems);
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        uint32_t e = warp_base + j * 32 + lane;
        key[j] = e < tile_count ? in[tile_base + e] : 0;
    }
    __syncthreads();

    // rank inside the warp.  Lanes with the same digit are found with one ballot per digit bit
    // (__match_any_sync does it in one instruction, but MATCH runs on the ADU pipe at a few instructions per
    // hundred cycles and made this kernel ADU-bound: profiles/r1g).  The group's leader takes the digit's running
    // count with a shared-memory atomic; a warp issues those in j order, so earlier items get smaller bases and
    // the sort stays stable without a warp-sync chain.
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t e = warp_base + j * 32 + lane;
        const bool valid = e < tile_count;
        const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & mask) : 0u;
        uint32_t same = __ballot_sync(kFullMask, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if ((mask >> b) == 0) break;  // uniform: digits narrower than 8 bits need fewer ballots
            const uint32_t vote = __ballot_sync(kFullMask, (d >> b) & 1u);
            same &= ((d >> b) & 1u) ? vote : ~vote;
        }
        const uint32_t lead = valid ? (uint32_t)__ffs(same) - 1u : lane;
        uint32_t base = 0;
        if (valid && lead == lane) base = atomicAdd(&s.warp_cnt[warp][d], (uint32_t)__popc(same));
        rank[j] = __popc(same & lanemask_lt()) + __shfl_sync(kFullMask, base, lead);
    }
    __syncthreads();

    // per digit: exclusive prefix over warps, then over digits
    uint32_t digit_total = 0;
    if (threadIdx.x < kRadix) {
        for (int w = 0; w < kRsWarps; ++w) {
            uint32_t c = s.warp_cnt[w][threadIdx.x];
            s.warp_cnt[w][threadIdx.x] = digit_total;
            digit_total += c;
        }
    }
    {
        // exclusive scan of digit_total over the first 256 threads (8 warps)
        uint32_t incl = warp_incl_scan_add(digit_total);
        if (threadIdx.x < kRadix && lane == 31) s.scan_tmp[warp] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < kRadix / 32 ? s.scan_tmp[threadIdx.x] : 0;
            uint32_t wi = warp_incl_scan_add(w);
            if (threadIdx.x < kRadix / 32) s.scan_tmp[threadIdx.x] = wi - w;
        }
        __syncthreads();
        if (threadIdx.x < kRadix) {
            uint32_t local = s.scan_tmp[warp] + incl - digit_total;
            s.digit_local[threadIdx.x] = local;
            uint32_t g = threadIdx.x <= mask ? table[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] : 0;
            s.digit_global[threadIdx.x] = g - local;
        }
    }
    __syncthreads();

    // stage the tile in digit order
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t e = warp_base + j * 32 + lane;
        if (e < tile_count) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & mask;
            s.keys[s.digit_local[d] + s.warp_cnt[warp][d] + rank[j]] = key[j];
        }
    }
    __syncthreads();

    // contiguous runs out to global memory
    for (uint32_t i = threadIdx.x; i < tile_count; i += kRsThreads) {
        const uint64_t k = s.keys[i];
        const uint32_t d = (uint32_t)(k >> shift) & mask;
        out[(uint64_t)(s.digit_global[d] + i)] = k;
    }
}

}  // namespace

int plan_radix_passes(int lo0, int hi0, int lo1, int hi1, RadixPass *out) {
    int n = 0;
    const int lo[2] = {lo0, lo1}, hi[2] = {hi0, hi1};
    for (int r = 0; r < 2; ++r) {
        int bits = hi[r] - lo[r];
        if (bits <= 0) continue;
        int np = (bits + 7) / 8;
        int w = (bits + np - 1) / np;
        int s = lo[r];
        while (s < hi[r]) {
            int b = hi[r] - s < w ? hi[r] - s : w;
            out[n++] = RadixPass{s, b};
            s += b;
        }
    }
    return n;
}

int radix_sort_u64(kombgpu_ctx *ctx, uint64_t *a, uint64_t *b, uint64_t n, const RadixPass *passes, int n_passes,
                   uint64_t **sorted) {
    *sorted = a;
    if (n < 2 || n_passes == 0) return KOMBGPU_OK;
    if (n >= (1ull << 32)) return ctx_fail(ctx, KOMBGPU_EINVAL, "radix_sort_u64: %llu keys exceed the 2^32 per-array limit", (unsigned long long)n);
    static bool attr_set = false;
    if (!attr_set) {
        KG_CUDA(ctx, cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
        attr_set = true;
    }
    const uint32_t n_tiles = ceil_div_u64(n, kRsTile);
    DevBuf<uint32_t> table;
    KG_ALLOC(ctx, table, (size_t)kRadix * n_tiles);
    uint64_t *src = a, *dst = b;
    for (int p = 0; p < n_passes; ++p) {
        const int shift = passes[p].shift;
        const uint32_t mask = (1u << passes[p].bits) - 1u;
        const uint64_t table_len = (uint64_t)(mask + 1) * n_tiles;
        KG_LAUNCH(ctx, radix_hist_kernel, n_tiles, kRsThreads, 0, src, n, shift, mask, table.p, n_tiles);
        KG_TRY((device_scan<uint32_t>(ctx, table_len, TableIn{table.p}, TableOut{table.p}, (uint32_t *)nullptr)));
        KG_LAUNCH(ctx, radix_scatter_kernel, n_tiles, kRsThreads, sizeof(ScatterSmem), src, dst, n, shift, mask, table.p, n_tiles);
        uint64_t *t = src; src = dst; dst = t;
    }
    *sorted = src;
    return KOMBGPU_OK;
}

}  // namespace kg
