/*
 * oracle/corea_ref.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin driver around the reference's own CoreA::getAnomalyScore, compiled
 * straight from /root/reference/src/CoreA.h (included where it lies, never
 * copied).  Built only in the authoring container into oracle/_ref/corea_ref.
 *
 * usage: corea_ref <in.bin> <out.bin>
 *   in.bin  = int32 n, int32 coreness[n], int32 degree[n]
 *   out.bin = double score[n]
 * Note the reference is O(n * distinct keys) (CoreA.h:142-187); keep n small.
 */
#include "CoreA.h"
#include <cstdio>
#include <vector>

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int n = 0;
    if (fread(&n, sizeof(int), 1, f) != 1 || n < 0) { fprintf(stderr, "bad header\n"); return 1; }
    std::vector<int> core(n), deg(n);
    if (fread(core.data(), sizeof(int), n, f) != (size_t)n || fread(deg.data(), sizeof(int), n, f) != (size_t)n) {
        fprintf(stderr, "short read\n"); return 1;
    }
    fclose(f);
    CoreA ca;
    double *score = ca.getAnomalyScore(deg, core);
    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 1; }
    fwrite(score, sizeof(double), n, o);
    fclose(o);
    free(score);
    return 0;
}
