/*
 * oracle/igraph_shim/igraph.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A minimal, header-only stand-in for the subset of the igraph >= 0.10 C API
 * that the KOMB reference (komb2) calls.  igraph itself is an un-vendored
 * dependency of the reference (komb.yml:7 `igraph>=0.10.0`, src/Makefile.am:2
 * `-ligraph`) and is not installed in this image, so the reference's own
 * sources (src/gfa.cpp, src/graph.cpp, src/komb2.cpp) are compiled UNMODIFIED
 * against this header to obtain `oracle/_ref/komb2_ref`.
 *
 * Every function here has a mathematically unique result on the inputs komb2
 * feeds it (simple-graph edge set, degree, coreness), so any correct
 * implementation is an exact stand-in:
 *   igraph_simplify  -> canonicalise (min,max), sort, unique, drop loops
 *   igraph_degree    -> count of non-loop incidences
 *   igraph_coreness  -> Batagelj-Zaversnik O(n+m) bucket peel (the algorithm
 *                       igraph 0.10 documents for igraph_coreness)
 * Call sites in the reference: graph.cpp:379-387,413-419,423-425,438,443-444,
 * 462-466,474,495-529,645,647; komb2.cpp:78,111,118; CombineCoreA.h:74-75,150.
 * Functions only reached from dead code (runTruss, runMerge) are provided so
 * the translation units link; they abort if ever called.
 */
#ifndef KOMB_ORACLE_IGRAPH_SHIM_H
#define KOMB_ORACLE_IGRAPH_SHIM_H

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

typedef int64_t igraph_integer_t;
typedef int igraph_bool_t;
typedef int igraph_error_t;
typedef double igraph_real_t;

enum { IGRAPH_SUCCESS = 0 };
enum igraph_neimode_t { IGRAPH_OUT = 1, IGRAPH_IN = 2, IGRAPH_ALL = 3 };
enum igraph_loops_t { IGRAPH_NO_LOOPS = 0, IGRAPH_LOOPS_TWICE = 1, IGRAPH_LOOPS_ONCE = 2 };
#define IGRAPH_LOOPS IGRAPH_LOOPS_TWICE
enum igraph_multiple_t { IGRAPH_NO_MULTIPLE = 0, IGRAPH_MULTIPLE = 1 };
enum igraph_subgraph_implementation_t { IGRAPH_SUBGRAPH_AUTO = 0 };
enum { IGRAPH_UNDIRECTED = 0, IGRAPH_DIRECTED = 1 };

/* ---- integer vector ---------------------------------------------------- */
struct igraph_vector_int_t {
    igraph_integer_t *stor_begin;
    igraph_integer_t *stor_end;
    igraph_integer_t *end;
};
#define VECTOR(v) ((v).stor_begin)

inline igraph_error_t igraph_vector_int_init(igraph_vector_int_t *v, igraph_integer_t size) {
    igraph_integer_t cap = size > 0 ? size : 1;
    v->stor_begin = (igraph_integer_t *)calloc((size_t)cap, sizeof(igraph_integer_t));
    if (!v->stor_begin) { fprintf(stderr, "igraph shim: out of memory\n"); abort(); }
    v->stor_end = v->stor_begin + cap;
    v->end = v->stor_begin + size;
    return IGRAPH_SUCCESS;
}
inline void igraph_vector_int_destroy(igraph_vector_int_t *v) {
    free(v->stor_begin);
    v->stor_begin = v->stor_end = v->end = nullptr;
}
inline igraph_integer_t igraph_vector_int_size(const igraph_vector_int_t *v) { return v->end - v->stor_begin; }
inline igraph_error_t igraph_vector_int_reserve(igraph_vector_int_t *v, igraph_integer_t cap) {
    igraph_integer_t cur = v->stor_end - v->stor_begin;
    if (cap <= cur) return IGRAPH_SUCCESS;
    igraph_integer_t sz = igraph_vector_int_size(v);
    igraph_integer_t *p = (igraph_integer_t *)realloc(v->stor_begin, (size_t)cap * sizeof(igraph_integer_t));
    if (!p) { fprintf(stderr, "igraph shim: out of memory\n"); abort(); }
    v->stor_begin = p; v->stor_end = p + cap; v->end = p + sz;
    return IGRAPH_SUCCESS;
}
inline igraph_error_t igraph_vector_int_resize(igraph_vector_int_t *v, igraph_integer_t n) {
    igraph_vector_int_reserve(v, n);
    v->end = v->stor_begin + n;
    return IGRAPH_SUCCESS;
}
inline void igraph_vector_int_resize_min(igraph_vector_int_t *v) {
    igraph_integer_t sz = igraph_vector_int_size(v);
    igraph_integer_t cap = sz > 0 ? sz : 1;
    igraph_integer_t *p = (igraph_integer_t *)realloc(v->stor_begin, (size_t)cap * sizeof(igraph_integer_t));
    if (p) { v->stor_begin = p; v->stor_end = p + cap; v->end = p + sz; }
}
inline void igraph_vector_int_set(igraph_vector_int_t *v, igraph_integer_t pos, igraph_integer_t val) { v->stor_begin[pos] = val; }
inline igraph_integer_t igraph_vector_int_get(const igraph_vector_int_t *v, igraph_integer_t pos) { return v->stor_begin[pos]; }
inline igraph_integer_t igraph_vector_int_max(const igraph_vector_int_t *v) {
    igraph_integer_t m = v->stor_begin[0];
    for (igraph_integer_t *p = v->stor_begin; p < v->end; ++p) if (*p > m) m = *p;
    return m;
}
inline igraph_error_t igraph_vector_int_push_back(igraph_vector_int_t *v, igraph_integer_t e) {
    if (v->end == v->stor_end) {
        igraph_integer_t cap = v->stor_end - v->stor_begin;
        igraph_vector_int_reserve(v, cap ? 2 * cap : 1);
    }
    *v->end++ = e;
    return IGRAPH_SUCCESS;
}

/* ---- string vector ----------------------------------------------------- */
struct igraph_strvector_t { std::vector<std::string> *s; };
inline igraph_error_t igraph_strvector_init(igraph_strvector_t *sv, igraph_integer_t n) {
    sv->s = new std::vector<std::string>((size_t)n);
    return IGRAPH_SUCCESS;
}
inline igraph_error_t igraph_strvector_set_len(igraph_strvector_t *sv, igraph_integer_t idx, const char *value, size_t len) {
    (*sv->s)[(size_t)idx].assign(value, len);
    return IGRAPH_SUCCESS;
}

/* ---- graph ------------------------------------------------------------- */
struct igraph_t {
    igraph_integer_t n;
    igraph_bool_t directed;
    std::vector<igraph_integer_t> *from, *to;                   /* edge list */
    std::map<std::string, std::vector<std::string>> *vattr_str; /* C attribute table, string vertex attrs */
};
struct igraph_attribute_table_t { int unused; };
static const igraph_attribute_table_t igraph_cattribute_table = {0};
inline igraph_attribute_table_t *igraph_set_attribute_table(const igraph_attribute_table_t *) { return nullptr; }

inline igraph_error_t igraph_create(igraph_t *g, const igraph_vector_int_t *edges, igraph_integer_t n, igraph_bool_t directed) {
    igraph_integer_t m2 = igraph_vector_int_size(edges);
    g->directed = directed;
    g->from = new std::vector<igraph_integer_t>();
    g->to = new std::vector<igraph_integer_t>();
    g->vattr_str = new std::map<std::string, std::vector<std::string>>();
    g->from->reserve((size_t)(m2 / 2));
    g->to->reserve((size_t)(m2 / 2));
    igraph_integer_t maxv = -1;
    for (igraph_integer_t i = 0; i + 1 < m2; i += 2) {
        igraph_integer_t a = VECTOR(*edges)[i], b = VECTOR(*edges)[i + 1];
        g->from->push_back(a); g->to->push_back(b);
        maxv = std::max(maxv, std::max(a, b));
    }
    g->n = std::max(n, maxv + 1);   /* igraph_create grows the vertex set to fit the ids */
    return IGRAPH_SUCCESS;
}
inline void igraph_destroy(igraph_t *g) {
    delete g->from; delete g->to; delete g->vattr_str;
    g->from = g->to = nullptr; g->vattr_str = nullptr;
}
inline igraph_integer_t igraph_vcount(const igraph_t *g) { return g->n; }
inline igraph_integer_t igraph_ecount(const igraph_t *g) { return (igraph_integer_t)g->from->size(); }

struct igraph_attribute_combination_t;
inline igraph_error_t igraph_simplify(igraph_t *g, igraph_bool_t multiple, igraph_bool_t loops, const igraph_attribute_combination_t *) {
    std::vector<std::pair<igraph_integer_t, igraph_integer_t>> e;
    e.reserve(g->from->size());
    for (size_t i = 0; i < g->from->size(); ++i) {
        igraph_integer_t a = (*g->from)[i], b = (*g->to)[i];
        if (loops && a == b) continue;
        if (!g->directed && a > b) std::swap(a, b);
        e.emplace_back(a, b);
    }
    if (multiple) {
        std::sort(e.begin(), e.end());
        e.erase(std::unique(e.begin(), e.end()), e.end());
    }
    g->from->clear(); g->to->clear();
    for (auto &p : e) { g->from->push_back(p.first); g->to->push_back(p.second); }
    return IGRAPH_SUCCESS;
}

struct igraph_vs_t { int type; const igraph_vector_int_t *vec; };
inline igraph_vs_t igraph_vss_all(void) { igraph_vs_t v = {0, nullptr}; return v; }
inline igraph_error_t igraph_vs_vector(igraph_vs_t *vs, const igraph_vector_int_t *v) { vs->type = 1; vs->vec = v; return IGRAPH_SUCCESS; }

inline igraph_error_t igraph_degree(const igraph_t *g, igraph_vector_int_t *res, igraph_vs_t vs, igraph_neimode_t, igraph_bool_t loops) {
    if (vs.type != 0) { fprintf(stderr, "igraph shim: igraph_degree supports igraph_vss_all() only\n"); abort(); }
    igraph_vector_int_resize(res, g->n);
    for (igraph_integer_t i = 0; i < g->n; ++i) VECTOR(*res)[i] = 0;
    for (size_t i = 0; i < g->from->size(); ++i) {
        igraph_integer_t a = (*g->from)[i], b = (*g->to)[i];
        if (a == b) { if (loops) VECTOR(*res)[a] += 2; continue; }
        VECTOR(*res)[a]++; VECTOR(*res)[b]++;
    }
    return IGRAPH_SUCCESS;
}

/* Batagelj & Zaversnik, "An O(m) algorithm for cores decomposition of
 * networks" (2003): bin-sort vertices by degree, sweep in increasing order. */
inline igraph_error_t igraph_coreness(const igraph_t *g, igraph_vector_int_t *cores, igraph_neimode_t) {
    const igraph_integer_t n = g->n;
    const size_t m = g->from->size();
    igraph_vector_int_resize(cores, n);
    if (n == 0) return IGRAPH_SUCCESS;
    std::vector<igraph_integer_t> deg((size_t)n, 0), off((size_t)n + 1, 0);
    for (size_t i = 0; i < m; ++i) {
        igraph_integer_t a = (*g->from)[i], b = (*g->to)[i];
        if (a == b) continue;
        deg[a]++; deg[b]++;
    }
    for (igraph_integer_t v = 0; v < n; ++v) off[v + 1] = off[v] + deg[v];
    std::vector<igraph_integer_t> adj((size_t)off[n]), cur(off.begin(), off.end() - 1);
    for (size_t i = 0; i < m; ++i) {
        igraph_integer_t a = (*g->from)[i], b = (*g->to)[i];
        if (a == b) continue;
        adj[cur[a]++] = b; adj[cur[b]++] = a;
    }
    igraph_integer_t md = 0;
    for (igraph_integer_t v = 0; v < n; ++v) md = std::max(md, deg[v]);
    std::vector<igraph_integer_t> bin((size_t)md + 1, 0), pos((size_t)n), vert((size_t)n);
    for (igraph_integer_t v = 0; v < n; ++v) bin[deg[v]]++;
    igraph_integer_t start = 0;
    for (igraph_integer_t d = 0; d <= md; ++d) { igraph_integer_t c = bin[d]; bin[d] = start; start += c; }
    for (igraph_integer_t v = 0; v < n; ++v) { pos[v] = bin[deg[v]]; vert[pos[v]] = v; bin[deg[v]]++; }
    for (igraph_integer_t d = md; d > 0; --d) bin[d] = bin[d - 1];
    bin[0] = 0;
    for (igraph_integer_t i = 0; i < n; ++i) {
        igraph_integer_t v = vert[i];
        for (igraph_integer_t j = off[v]; j < off[v + 1]; ++j) {
            igraph_integer_t u = adj[j];
            if (deg[u] > deg[v]) {
                igraph_integer_t du = deg[u], pu = pos[u], pw = bin[du], w = vert[pw];
                if (u != w) { pos[u] = pw; vert[pu] = w; pos[w] = pu; vert[pw] = u; }
                bin[du]++; deg[u]--;
            }
        }
    }
    for (igraph_integer_t v = 0; v < n; ++v) VECTOR(*cores)[v] = deg[v];
    return IGRAPH_SUCCESS;
}

/* ---- C attribute table (string vertex attributes only) ------------------ */
inline igraph_error_t igraph_cattribute_VAS_setv(igraph_t *g, const char *name, const igraph_strvector_t *sv) {
    (*g->vattr_str)[name] = *sv->s;
    return IGRAPH_SUCCESS;
}
#define SETVASV(graph, n, v) (igraph_cattribute_VAS_setv((graph), (n), (v)))
inline const char *igraph_cattribute_VAS(const igraph_t *g, const char *name, igraph_integer_t vid) {
    return (*g->vattr_str)[name][(size_t)vid].c_str();
}

/* ---- lazy adjacency list: initialised by anomalyDetection, never read ---- */
struct igraph_lazy_adjlist_t { const igraph_t *graph; };
inline igraph_error_t igraph_lazy_adjlist_init(const igraph_t *g, igraph_lazy_adjlist_t *al, igraph_neimode_t, igraph_loops_t, igraph_multiple_t) {
    al->graph = g;
    return IGRAPH_SUCCESS;
}

/* ---- reached only from dead code (runTruss / runMerge): link, never run -- */
[[noreturn]] inline void igraph_shim_dead(const char *fn) {
    fprintf(stderr, "igraph shim: %s is only reachable from code komb2 never calls\n", fn);
    abort();
}
inline igraph_vector_int_t *igraph_lazy_adjlist_get(igraph_lazy_adjlist_t *, igraph_integer_t) { igraph_shim_dead("igraph_lazy_adjlist_get"); }
inline igraph_error_t igraph_induced_subgraph_map(const igraph_t *, igraph_t *, igraph_vs_t, igraph_subgraph_implementation_t, igraph_vector_int_t *, igraph_vector_int_t *) { igraph_shim_dead("igraph_induced_subgraph_map"); }
inline igraph_error_t igraph_trussness(const igraph_t *, igraph_vector_int_t *) { igraph_shim_dead("igraph_trussness"); }
inline igraph_error_t igraph_edge(const igraph_t *, igraph_integer_t, igraph_integer_t *, igraph_integer_t *) { igraph_shim_dead("igraph_edge"); }

#endif /* KOMB_ORACLE_IGRAPH_SHIM_H */
