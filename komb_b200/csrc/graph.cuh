// graph.cuh — the device-resident unitig graph and the stage entry points that
// the C ABI (capi.cu) strings together.
#pragma once

#include "common.cuh"

struct kombgpu_graph {
    kombgpu_ctx *ctx = nullptr;
    uint32_t n = 0;            // vertices (hit unitigs)
    uint64_t n_edges = 0;      // simple undirected edges
    uint64_t *edges = nullptr;   // [E]   packed (u << 32 | v), u < v, ascending
    uint64_t *row_ptr = nullptr; // [n+1] CSR offsets of the symmetric graph
    uint32_t *col = nullptr;     // [2E]  neighbours, every row ascending
    uint32_t *fwd_start = nullptr; // [n+1] forward index of the edge list: edges [fwd_start[u], fwd_start[u+1]) have source u
    uint32_t *mult = nullptr;    // [E]   pairs (reads, or input duplicates) that support each edge
    int32_t *deg = nullptr;      // [n]
    int32_t *core = nullptr;     // [n]   after peel
    double *score = nullptr;     // [n]   after CORE-A
    double max_score = 0.0;
    bool has_core = false, has_score = false;
    kombgpu_stats st{};
    // maximal core + trussness of its edges (truss.cu), computed on demand
    bool has_truss = false;
    uint32_t tr_n_core = 0, tr_n_vertices = 0;
    uint64_t tr_m = 0;
    int32_t tr_max = 0;
    uint32_t *tr_core_vid = nullptr;   // [tr_n_core] original ids of the maximal core, ascending
    uint64_t *tr_edges = nullptr;      // [tr_m] induced edges, compact ids (a << 32 | b), canonical order
    int32_t *tr_truss = nullptr;       // [tr_m]
    uint32_t *tr_vertices = nullptr;   // [tr_n_vertices] original ids of the unitigs on edges of maximal trussness
    // output files formatted on the device (format.cu): edgelist.txt, kcore.tsv, CoreA_anomaly.txt
    char *text[3] = {nullptr, nullptr, nullptr};
    uint64_t text_bytes[3] = {0, 0, 0};
};

namespace kg {

// stage 1 (build.cu) — inputs are device pointers
int build_from_hits(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits,
                    uint32_t n_vertices, kombgpu_graph *g);
int build_from_pairs(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices,
                     kombgpu_graph *g);

// building blocks of stage 1 shared with the partitioned (multi-GPU) path (build.cu)
// `mult` (optional): receives, per unique edge, how many emitted pairs collapsed into it
int hits_to_edges(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits, uint32_t n_vertices,
                  DevBuf<uint64_t> &edges, uint64_t *n_edges, kombgpu_stats *st, DevBuf<uint32_t> *mult = nullptr);
int pairs_to_edges(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices,
                   DevBuf<uint64_t> &edges, uint64_t *n_edges, DevBuf<uint32_t> *mult = nullptr);
int hits_to_pairs(kombgpu_ctx *ctx, const uint32_t *read_key, const uint32_t *unitig, uint64_t n_hits, uint32_t n_vertices,
                  DevBuf<uint64_t> &pairs, uint64_t *n_pairs, kombgpu_stats *st);
int pairs_to_keys(kombgpu_ctx *ctx, const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t n_vertices, DevBuf<uint64_t> &keys);
// sorted canonical keys (duplicates and loops in) -> unique simple edges (+ optional multiplicities)
int unique_edges(kombgpu_ctx *ctx, const uint64_t *keys, uint64_t count, DevBuf<uint64_t> &edges, uint64_t *n_edges,
                 DevBuf<uint32_t> *mult);
// keys sorted by high word in [base, base + n_rows): start[x] = first index whose high word is >= base + x, x in [0, n_rows];
// long runs of empty rows are filled by whole CTAs.  *err_dev (optional) is set to 2 when a key is out of range or its
// low word is >= lo_bound.
int row_starts(kombgpu_ctx *ctx, const uint64_t *keys, uint64_t count, uint32_t base, uint32_t n_rows, uint32_t lo_bound, uint32_t *start,
               uint32_t *err_dev);
// fwd_start[x] = first edge whose source is >= x, x in [0, n]  (the edge list is sorted by source)
int forward_index(kombgpu_ctx *ctx, const uint64_t *edges, uint64_t n_edges, uint32_t n, uint32_t **fwd_start_out);
int swapped_sorted(kombgpu_ctx *ctx, const uint64_t *edges, uint64_t n_edges, uint32_t n_vertices, DevBuf<uint64_t> &a,
                   DevBuf<uint64_t> &b, uint64_t **out);
int lower_bounds_hi(kombgpu_ctx *ctx, const uint64_t *keys, uint64_t count, const uint32_t *bounds_host, int nb,
                    uint64_t *idx_host);
int csr_from_directed(kombgpu_ctx *ctx, const uint64_t *entries, uint64_t count, uint32_t v_lo, uint32_t n_local,
                      uint32_t n_global, uint64_t **row_ptr_out, uint32_t **col_out, int32_t **deg_out, int32_t *max_deg_out,
                      uint64_t *n_directed_out);

int csr_from_edges(kombgpu_ctx *ctx, DevBuf<uint64_t> &edges, uint64_t n_edges, uint32_t n, kombgpu_graph *g);

int adopt_csr(kombgpu_ctx *ctx, const uint64_t *row_ptr, const uint32_t *col, uint32_t n, kombgpu_graph *g);

// stage 2 (peel.cu)
int peel_coreness(kombgpu_graph *g);

// stage 3 (corea.cu) — device pointers in, device score out; *max_score on host
int corea_scores(kombgpu_ctx *ctx, const int32_t *core, const int32_t *deg, uint32_t n, int key_mode, double *score,
                 double *max_score_host);

void graph_release(kombgpu_graph *g);
void truss_release(kombgpu_graph *g);   // truss.cu
void format_release(kombgpu_graph *g);  // format.cu

}  // namespace kg
