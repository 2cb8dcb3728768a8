// densest.cu — densest k-core: the bulk analogue of the serial greedy densest-block peel the reference keeps
// (unreachable from main) in CombineCoreA::runMerge over HashIndexedMinHeap (src/CombineCoreA.h:45-219,
// src/HashIndexedMinHeap.h).  That code pops the minimum-priority node again and again, tracks
// density = (edges + suspiciousness left) / (nodes left) and returns the densest prefix of the removal order
// (Charikar's greedy 2-approximation run on a row copy and a column copy of every vertex; its tie order and an
// uninitialised `removed[]` make its exact output accidental, so it is not a parity target: SURVEY.md row A9).
// The peel this library already ran visits the same nested family of blocks level by level: the k-cores.  The
// densest of them is a 2-approximation of the densest subgraph as well (the ceil(rho*)-core is not empty and has
// minimum degree >= rho*), and it falls out of two histograms over the finished coreness array:
//   V_k = #{v : core(v) >= k},  E_k = #{(u,v) in E : min(core(u), core(v)) >= k},  density_k = E_k / V_k
// (the reference's suspSum / numNodes for suspiciousness == nullptr: 2 E_k directed entries over 2 V_k row+column
// nodes).  Among blocks of equal density the largest wins, as with the reference's strict `>` while it shrinks the
// block; k* is reported as the smallest coreness inside the block.
#include <vector>

#include "graph.cuh"

namespace kg {
namespace {

constexpr int kThreads = 256;
constexpr int kSmemBins = 4096;

// kEdges = false: hist[core[v]] += 1 over the `count` vertices; kEdges = true: hist[min(core[u], core[v])] += 1 over the
// `count` edges of the canonical edge list
template <bool kEdges>
__global__ void __launch_bounds__(kThreads) level_hist_kernel(const int32_t *__restrict__ core, const uint64_t *__restrict__ edges,
                                                              uint64_t count, uint32_t n_bins,
                                                              unsigned long long *__restrict__ hist) {
    __shared__ uint32_t s_hist[kSmemBins];
    const bool use_smem = n_bins <= kSmemBins;
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < n_bins; i += kThreads) s_hist[i] = 0;
        __syncthreads();
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        int32_t c;
        if (kEdges) {
            const uint64_t e = edges[i];
            c = min(core[(uint32_t)(e >> 32)], core[(uint32_t)e]);
        } else {
            c = core[i];
        }
        if (use_smem) atomicAdd(&s_hist[c], 1u);
        else atomicAdd(&hist[c], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_bins; i += kThreads)
            if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
    }
}

}  // namespace
}  // namespace kg

extern "C" int kombgpu_graph_densest_core(kombgpu_graph *g, int32_t *k_star, uint32_t *n_vertices, uint64_t *n_edges,
                                          double *density) {
    using namespace kg;
    if (!g) return KOMBGPU_EINVAL;
    kombgpu_ctx *ctx = g->ctx;
    if (!g->has_core) return ctx_fail(ctx, KOMBGPU_ESTATE, "kombgpu_graph_densest_core needs kombgpu_coreness first");
    if (g->n_edges && !g->edges) return ctx_fail(ctx, KOMBGPU_ESTATE, "graph was adopted from a CSR: no canonical edge list");
    KG_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t best_k = 0;
    uint64_t best_v = g->n, best_e = g->n_edges;
    double best_d = g->n ? (double)g->n_edges / (double)g->n : 0.0;
    if (g->n && g->n_edges) {
        const uint32_t n_bins = (uint32_t)g->st.max_coreness + 1;
        DevBuf<unsigned long long> hist;
        KG_ALLOC(ctx, hist, 2 * (size_t)n_bins);
        KG_CUDA(ctx, cudaMemsetAsync(hist.p, 0, 2 * (size_t)n_bins * sizeof(unsigned long long), ctx->stream));
        const uint32_t cap = (uint32_t)ctx->sm_count * 8u;
        KG_LAUNCH(ctx, level_hist_kernel<false>, min(ceil_div_u64(g->n, kThreads), cap), kThreads, 0, g->core, (const uint64_t *)nullptr,
                  (uint64_t)g->n, n_bins, hist.p);
        KG_LAUNCH(ctx, level_hist_kernel<true>, min(ceil_div_u64(g->n_edges, kThreads), cap), kThreads, 0, g->core, g->edges, g->n_edges,
                  n_bins, hist.p + n_bins);
        std::vector<unsigned long long> h(2 * (size_t)n_bins);
        KG_TRY(read_back(ctx, hist.p, h.data(), h.size()));
        uint64_t v = 0, e = 0;
        best_d = -1.0;
        for (int64_t k = (int64_t)n_bins - 1; k >= 0; --k) {   // suffix sums over the levels, densest first
            v += h[k];
            e += h[n_bins + k];
            if (v == 0) continue;
            const double d = (double)e / (double)v;
            // equal density: the larger block wins; an empty level leaves the block (and the reported k) as it is
            if (d > best_d || (d == best_d && v > best_v)) { best_d = d; best_k = (int32_t)k; best_v = v; best_e = e; }
        }
    }
    if (k_star) *k_star = best_k;
    if (n_vertices) *n_vertices = (uint32_t)best_v;
    if (n_edges) *n_edges = best_e;
    if (density) *density = best_d;
    return KOMBGPU_OK;
}
