/*
 * oracle/komb_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of KOMB's graph-analysis hot path
 * (hits -> unitig adjacency graph -> k-core -> CORE-A), used ONLY as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Nothing under komb_b200/ may import, link or call this file.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle.py)
 * against (1) the outputs of the reference itself, `oracle/_ref/komb2_ref`
 * (reference sources compiled unmodified against oracle/igraph_shim) at -t 1,
 * committed as fixtures under tests/golden/, (2) CoreA::getAnomalyScore
 * compiled straight from the reference's src/CoreA.h (oracle/corea_ref.cpp),
 * and (3) networkx.core_number / scipy.stats.rankdata.  The reference ships no
 * tests or golden vectors of its own.
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_OK 0
#define KO_ENOMEM (-1)
#define KO_EINVAL (-2)

/* ---- helpers ------------------------------------------------------------ */

/* LSD radix sort of 64-bit keys, 16 bits per pass, skipping constant digits. */
static int ko_sort_u64(uint64_t *a, uint64_t n)
{
    if (n < 2) return KO_OK;
    uint64_t *b = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *cnt = (uint64_t *)malloc(65536 * sizeof(uint64_t));
    if (!b || !cnt) { free(b); free(cnt); return KO_ENOMEM; }
    uint64_t *src = a, *dst = b;
    for (int pass = 0; pass < 4; ++pass) {
        const int sh = pass * 16;
        memset(cnt, 0, 65536 * sizeof(uint64_t));
        for (uint64_t i = 0; i < n; ++i) cnt[(src[i] >> sh) & 0xffff]++;
        if (cnt[(src[0] >> sh) & 0xffff] == n) continue;      /* all equal in this digit */
        uint64_t run = 0;
        for (int d = 0; d < 65536; ++d) { uint64_t c = cnt[d]; cnt[d] = run; run += c; }
        for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> sh) & 0xffff]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(b); free(cnt);
    return KO_OK;
}

static uint64_t ko_unique_u64(uint64_t *a, uint64_t n)
{
    if (n == 0) return 0;
    uint64_t w = 1;
    for (uint64_t i = 1; i < n; ++i) if (a[i] != a[w - 1]) a[w++] = a[i];
    return w;
}

void ko_free(void *p) { free(p); }

/* ---- stage 1: hits -> simple edge set ----------------------------------- */

/*
 * Hits (read_key[i], unitig[i]) of BOTH mate files, concatenated.
 *
 * Follows:
 *   src/graph.cpp:235      each hit inserts RNAME into the SET of its read key
 *   src/graph.cpp:259-285  getEdgeInfo: per key, mate-1 set U= mate-2 set; keys
 *                          only in mate 2 are kept (=> set union over the
 *                          concatenated hits of both files)
 *   src/graph.cpp:332-347  generateGraph: every set is a clique, all i<j pairs
 *   src/graph.cpp:438      igraph_simplify(multiple, loops): undirected simple
 *                          edge set (a set has no repeated unitig, so no loops
 *                          arise; duplicates across reads collapse)
 * Output: packed edges (u << 32 | v) with u < v, ascending; *P_out = number of
 * pairs emitted before dedup, *S_out = number of distinct (read, unitig) hits.
 */
int ko_build_edges(const uint32_t *read_key, const uint32_t *unitig, uint64_t H,
                   uint64_t **edges_out, uint64_t *E_out, uint64_t *P_out, uint64_t *S_out)
{
    *edges_out = NULL; *E_out = 0; *P_out = 0; *S_out = 0;
    if (H == 0) return KO_OK;
    uint64_t *hk = (uint64_t *)malloc(H * sizeof(uint64_t));
    if (!hk) return KO_ENOMEM;
    for (uint64_t i = 0; i < H; ++i) hk[i] = ((uint64_t)read_key[i] << 32) | unitig[i];
    if (ko_sort_u64(hk, H) != KO_OK) { free(hk); return KO_ENOMEM; }
    uint64_t S = ko_unique_u64(hk, H);
    *S_out = S;
    /* count pairs */
    uint64_t P = 0;
    for (uint64_t s = 0; s < S;) {
        uint64_t e = s + 1;
        while (e < S && (hk[e] >> 32) == (hk[s] >> 32)) ++e;
        uint64_t k = e - s;
        P += k * (k - 1) / 2;
        s = e;
    }
    *P_out = P;
    if (P == 0) { free(hk); return KO_OK; }
    uint64_t *pr = (uint64_t *)malloc(P * sizeof(uint64_t));
    if (!pr) { free(hk); return KO_ENOMEM; }
    uint64_t w = 0;
    for (uint64_t s = 0; s < S;) {
        uint64_t e = s + 1;
        while (e < S && (hk[e] >> 32) == (hk[s] >> 32)) ++e;
        for (uint64_t i = s; i < e; ++i)
            for (uint64_t j = i + 1; j < e; ++j) {
                uint32_t a = (uint32_t)hk[i], b = (uint32_t)hk[j];   /* a < b: sorted, unique */
                pr[w++] = ((uint64_t)a << 32) | b;
            }
        s = e;
    }
    free(hk);
    if (ko_sort_u64(pr, P) != KO_OK) { free(pr); return KO_ENOMEM; }
    *E_out = ko_unique_u64(pr, P);
    *edges_out = pr;
    return KO_OK;
}

/*
 * Arbitrary (u, v) pairs -> simple undirected edge set.
 * Follows src/graph.cpp:418 (igraph_create, undirected) + :438 (igraph_simplify
 * with multiple=true, loops=true): canonical (min,max), no loops, no repeats.
 */
int ko_simplify(const uint32_t *u, const uint32_t *v, uint64_t m,
                uint64_t **edges_out, uint64_t *E_out)
{
    *edges_out = NULL; *E_out = 0;
    if (m == 0) return KO_OK;
    uint64_t *e = (uint64_t *)malloc(m * sizeof(uint64_t));
    if (!e) return KO_ENOMEM;
    uint64_t w = 0;
    for (uint64_t i = 0; i < m; ++i) {
        uint32_t a = u[i], b = v[i];
        if (a == b) continue;
        if (a > b) { uint32_t t = a; a = b; b = t; }
        e[w++] = ((uint64_t)a << 32) | b;
    }
    if (ko_sort_u64(e, w) != KO_OK) { free(e); return KO_ENOMEM; }
    *E_out = ko_unique_u64(e, w);
    *edges_out = e;
    return KO_OK;
}

/* ---- stage 2: degree + coreness ----------------------------------------- */

/*
 * Follows src/graph.cpp:462 (igraph_degree, ALL, NO_LOOPS) and :463
 * (igraph_coreness, ALL).  igraph documents igraph_coreness as the
 * Batagelj-Zaversnik O(m) bucket algorithm; coreness is a unique function of
 * the simple graph, so the result is implementation-independent.
 * edges: packed (u << 32 | v), simple (no loops, no repeats).
 */
int ko_coreness(uint32_t n, const uint64_t *edges, uint64_t E, int32_t *deg_out, int32_t *core_out)
{
    if (n == 0) return KO_OK;
    int64_t *off = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    int32_t *deg = (int32_t *)calloc(n, sizeof(int32_t));
    if (!off || !deg) { free(off); free(deg); return KO_ENOMEM; }
    for (uint64_t i = 0; i < E; ++i) {
        uint32_t a = (uint32_t)(edges[i] >> 32), b = (uint32_t)edges[i];
        if (a >= n || b >= n || a == b) { free(off); free(deg); return KO_EINVAL; }
        deg[a]++; deg[b]++;
    }
    int32_t md = 0;
    for (uint32_t v = 0; v < n; ++v) { off[v + 1] = off[v] + deg[v]; if (deg[v] > md) md = deg[v]; }
    if (deg_out) memcpy(deg_out, deg, n * sizeof(int32_t));
    uint32_t *adj = (uint32_t *)malloc((size_t)(off[n] ? off[n] : 1) * sizeof(uint32_t));
    int64_t *cur = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t *bin = (int64_t *)calloc((size_t)md + 2, sizeof(int64_t));
    int64_t *pos = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    uint32_t *vert = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    if (!adj || !cur || !bin || !pos || !vert) {
        free(off); free(deg); free(adj); free(cur); free(bin); free(pos); free(vert);
        return KO_ENOMEM;
    }
    memcpy(cur, off, n * sizeof(int64_t));
    for (uint64_t i = 0; i < E; ++i) {
        uint32_t a = (uint32_t)(edges[i] >> 32), b = (uint32_t)edges[i];
        adj[cur[a]++] = b; adj[cur[b]++] = a;
    }
    for (uint32_t v = 0; v < n; ++v) bin[deg[v]]++;
    int64_t start = 0;
    for (int32_t d = 0; d <= md; ++d) { int64_t c = bin[d]; bin[d] = start; start += c; }
    for (uint32_t v = 0; v < n; ++v) { pos[v] = bin[deg[v]]; vert[pos[v]] = v; bin[deg[v]]++; }
    for (int32_t d = md; d > 0; --d) bin[d] = bin[d - 1];
    bin[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t v = vert[i];
        for (int64_t j = off[v]; j < off[v + 1]; ++j) {
            uint32_t u = adj[j];
            if (deg[u] > deg[v]) {
                int32_t du = deg[u];
                int64_t pu = pos[u], pw = bin[du];
                uint32_t w = vert[pw];
                if (u != w) { pos[u] = pw; vert[pu] = w; pos[w] = pu; vert[pw] = u; }
                bin[du]++; deg[u]--;
            }
        }
    }
    memcpy(core_out, deg, n * sizeof(int32_t));
    free(off); free(deg); free(adj); free(cur); free(bin); free(pos); free(vert);
    return KO_OK;
}

/* ---- stage 3: CORE-A ----------------------------------------------------- */

typedef struct { int64_t key; uint32_t idx; } ko_kv;

static int ko_kv_desc(const void *a, const void *b)
{
    int64_t x = ((const ko_kv *)a)->key, y = ((const ko_kv *)b)->key;
    return (x < y) - (x > y);
}

/*
 * Descending average-tie ranks.  Follows src/CoreA.h:142-187 (fractionalRank):
 * distinct values visited in DESCENDING order (:154), a running `rank` counter
 * incremented once per member (:169), all members of a tie class receive the
 * mean of the ranks they span (:175-183).  For a class starting after `start`
 * earlier elements with `cnt` members the mean is start + (cnt+1)/2, exact in
 * double for n < 2^26.  O(n log n) instead of O(n * distinct).
 */
static int ko_frac_rank(const int64_t *key, uint32_t n, double *rank)
{
    ko_kv *kv = (ko_kv *)malloc((size_t)(n ? n : 1) * sizeof(ko_kv));
    if (!kv) return KO_ENOMEM;
    for (uint32_t i = 0; i < n; ++i) { kv[i].key = key[i]; kv[i].idx = i; }
    qsort(kv, n, sizeof(ko_kv), ko_kv_desc);
    for (uint32_t s = 0; s < n;) {
        uint32_t e = s + 1;
        while (e < n && kv[e].key == kv[s].key) ++e;
        /* ranks s+1 .. e, summed exactly like the reference's running `avg += rank` */
        double sum = 0.0;
        for (uint32_t r = s + 1; r <= e; ++r) sum += (double)r;
        double avg = sum / (double)(e - s);
        for (uint32_t i = s; i < e; ++i) rank[kv[i].idx] = avg;
        s = e;
    }
    free(kv);
    return KO_OK;
}

/*
 * key_mode 0 = "ref32": key = (int32)(coreness * n + degree) with two's
 *              complement wrap, exactly what src/CoreA.h:122 computes in `int`
 *              on every mainstream ABI;
 * key_mode 1 = "exact64": the same expression in int64 (no overflow).
 * score[i] = | ln rank_desc(degree)[i] - ln rank_desc(key)[i] |   (CoreA.h:125-132)
 */
int ko_corea(uint32_t n, const int32_t *core, const int32_t *deg, int key_mode, double *score)
{
    if (n == 0) return KO_OK;
    int64_t *key = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    double *rk = (double *)malloc((size_t)n * sizeof(double));
    double *rd = (double *)malloc((size_t)n * sizeof(double));
    if (!key || !rk || !rd) { free(key); free(rk); free(rd); return KO_ENOMEM; }
    for (uint32_t i = 0; i < n; ++i) {
        if (key_mode == 0) {
            uint32_t w = (uint32_t)core[i] * (uint32_t)n + (uint32_t)deg[i];
            key[i] = (int64_t)(int32_t)w;
        } else {
            key[i] = (int64_t)core[i] * (int64_t)n + (int64_t)deg[i];
        }
    }
    int rc = ko_frac_rank(key, n, rk);
    for (uint32_t i = 0; i < n; ++i) key[i] = deg[i];
    if (rc == KO_OK) rc = ko_frac_rank(key, n, rd);
    if (rc == KO_OK)
        for (uint32_t i = 0; i < n; ++i) score[i] = fabs(log(rd[i]) - log(rk[i]));
    free(key); free(rk); free(rd);
    return rc;
}

/* Intermediate ranks, for tests that want exact (half-integer) comparisons. */
int ko_corea_ranks(uint32_t n, const int32_t *core, const int32_t *deg, int key_mode,
                   double *deg_rank, double *key_rank)
{
    if (n == 0) return KO_OK;
    int64_t *key = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    if (!key) return KO_ENOMEM;
    for (uint32_t i = 0; i < n; ++i) {
        if (key_mode == 0) {
            uint32_t w = (uint32_t)core[i] * (uint32_t)n + (uint32_t)deg[i];
            key[i] = (int64_t)(int32_t)w;
        } else {
            key[i] = (int64_t)core[i] * (int64_t)n + (int64_t)deg[i];
        }
    }
    int rc = ko_frac_rank(key, n, key_rank);
    for (uint32_t i = 0; i < n; ++i) key[i] = deg[i];
    if (rc == KO_OK) rc = ko_frac_rank(key, n, deg_rank);
    free(key);
    return rc;
}
