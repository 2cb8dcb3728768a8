// komb2_tokenize — test helper: runs the komb2 host tokeniser + interner on two
// SAM files and prints the integer hits (no GPU needed).
//   usage: komb2_tokenize <threads> <r1.sam> <r2.sam>
//   stdout: "N <n_vertices>" then one "name" line per vid, then "H <n_hits>" and
//           one "<read_key>\t<vid>" line per hit (file order).
#include <omp.h>

#include <cstdio>
#include <cstdlib>

#include "sam_tokenizer.hpp"

int main(int argc, char **argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s threads r1.sam r2.sam\n", argv[0]); return 2; }
    omp_set_dynamic(0);
    const int threads = atoi(argv[1]);
    komb::MappedFile f1, f2;
    komb::SamTokens tok;
    try {
        if (!f1.open(argv[2]) || !f2.open(argv[3])) { fprintf(stderr, "cannot open input\n"); return 1; }
        komb::tokenise_sam(f1, threads, tok, argv[2]);
        komb::tokenise_sam(f2, threads, tok, argv[3]);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    komb::InternResult keys = komb::intern_spans(tok.keys, threads, false);
    std::vector<komb::Span> items(tok.sq);
    items.insert(items.end(), tok.rnames.begin(), tok.rnames.end());
    komb::InternResult names = komb::intern_spans(items, threads, true);
    std::vector<int> has(names.n_distinct, 0);
    for (size_t i = tok.sq.size(); i < items.size(); ++i) has[names.ids[i]] = 1;
    std::vector<uint32_t> vid(names.n_distinct, 0);
    uint32_t n = 0;
    for (uint32_t d = 0; d < names.n_distinct; ++d) if (has[d]) vid[d] = n++;
    printf("N %u\n", n);
    for (uint32_t d = 0; d < names.n_distinct; ++d)
        if (has[d]) printf("%.*s\n", (int)items[names.first_index[d]].len, items[names.first_index[d]].p);
    printf("H %zu\n", tok.keys.size());
    for (size_t i = 0; i < tok.keys.size(); ++i) printf("%u\t%u\n", keys.ids[i], vid[names.ids[tok.sq.size() + i]]);
    return 0;
}
