/*
 * oracle/densest_ref.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The serial greedy densest-block peel of CombineCoreA::runMerge (/root/reference/src/CombineCoreA.h:45-219) driven
 * over the reference's OWN priority queue, HashIndexedMinHeap (src/HashIndexedMinHeap.h, included where it lies, never
 * copied): the removal order -- ties included -- is the reference's.  runMerge itself cannot be called: it is
 * unreachable from main, reads an uninitialised `removed[]` (:105-109) and sizes `cols` with the number of rows (:192);
 * the loop below restates :56-186 with `removed` cleared and the two result lists sized by their own counts.
 * Built only in the authoring container into oracle/_ref/densest_ref.
 *
 * usage: densest_ref <in.bin> <out.bin>
 *   in.bin  = int32 n, int32 weighted, int64 m, double w[n] (if weighted), int32 u[m], int32 v[m]   (simple edges)
 *   out.bin = double density, int32 n_rows, int32 n_cols, int32 rows[n_rows], int32 cols[n_cols]
 */
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "HashIndexedMinHeap.h"

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int32_t n = 0, weighted = 0;
    int64_t m = 0;
    if (fread(&n, 4, 1, f) != 1 || fread(&weighted, 4, 1, f) != 1 || fread(&m, 8, 1, f) != 1 || n < 0 || m < 0) return 1;
    std::vector<double> w(n, 0.0);
    if (weighted && fread(w.data(), 8, n, f) != (size_t)n) return 1;
    std::vector<int32_t> eu(m), ev(m);
    if (fread(eu.data(), 4, m, f) != (size_t)m || fread(ev.data(), 4, m, f) != (size_t)m) return 1;
    fclose(f);
    std::vector<std::vector<int>> adj(n);            // what igraph_lazy_adjlist_get(ALL) hands runMerge: every neighbour once
    for (int64_t i = 0; i < m; ++i) { adj[eu[i]].push_back(ev[i]); adj[ev[i]].push_back(eu[i]); }

    // :56-93  priorities = suspiciousness + degree; suspiciousSum = 2 * sum(w) + directed entries
    std::vector<double> rowDegree(w), colDegree(w);
    double suspiciousSum = 0;
    for (int i = 0; i < n; ++i) suspiciousSum += 2 * w[i];
    long edgeNum = 0;
    for (int src = 0; src < n; ++src)
        for (int dst : adj[src]) { rowDegree[src] += 1; colDegree[dst] += 1; edgeNum += 1; }
    suspiciousSum += edgeNum;
    HashIndexedMinHeap rowHeap(n > 0 ? n : 1), colHeap(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) rowHeap.insert(i, rowDegree[i]);
    for (int j = 0; j < n; ++j) colHeap.insert(j, colDegree[j]);
    std::vector<int> modes(2 * (size_t)n), order(2 * (size_t)n);
    std::vector<char> removed[2] = {std::vector<char>(n, 0), std::vector<char>(n, 0)};
    double maxDensity = 0;
    int maxDensityNodesNum = 0, numOfNodesBelong = 2 * n;
    // :112-173
    while (numOfNodesBelong >= 1) {
        std::pair<int, double> rowPair = rowHeap.peek(), colPair = colHeap.peek(), pair;
        int modeToRemove;
        if ((rowPair.first != INT_MIN && rowPair.second != INT_MIN) &&
            ((colPair.first == INT_MIN && colPair.second == INT_MIN) || rowPair.second < colPair.second)) {
            pair = rowHeap.poll();
            modeToRemove = 0;
        } else {
            pair = colHeap.poll();
            modeToRemove = 1;
        }
        suspiciousSum -= pair.second;
        const int node = pair.first;
        order[--numOfNodesBelong] = node;
        modes[numOfNodesBelong] = modeToRemove;
        const double density = suspiciousSum / numOfNodesBelong;
        if (numOfNodesBelong >= 1 && density > maxDensity) { maxDensity = density; maxDensityNodesNum = numOfNodesBelong; }
        removed[modeToRemove][node] = 1;
        HashIndexedMinHeap &other = modeToRemove == 0 ? colHeap : rowHeap;
        for (int dst : adj[node])
            if (!removed[1 - modeToRemove][dst]) other.refreshPriority(dst, other.getPriority(dst) - 1);
    }
    std::vector<int32_t> rows, cols;
    for (int i = 0; i < maxDensityNodesNum; ++i) (modes[i] == 0 ? rows : cols).push_back(order[i]);
    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 1; }
    const int32_t nr = (int32_t)rows.size(), nc = (int32_t)cols.size();
    fwrite(&maxDensity, 8, 1, o);
    fwrite(&nr, 4, 1, o);
    fwrite(&nc, 4, 1, o);
    fwrite(rows.data(), 4, rows.size(), o);
    fwrite(cols.data(), 4, cols.size(), o);
    fclose(o);
    return 0;
}
