// sam_tokenizer.hpp — host side of the drop-in boundary: SAM text -> integer hits.
//
// Produces exactly the information the reference keeps after tokenising a SAM
// file (reference src/graph.cpp:197-239, at -t 1):
//   * lines starting with '@' are skipped (:220)            [@SQ names are noted, see below]
//   * token 0 (QNAME) and token 2 (RNAME) under strtok("\t") rules, i.e. runs
//     of tabs collapse (:222-231, quirk Q3)
//   * RNAME "*" (unmapped) is skipped (:232)
//   * read key = QNAME.substr(1, QNAME.find('/')) (:235, quirk Q2)
//   * a final line without '\n' is not processed (the -t 1 behaviour of :206-218)
// Unlike the reference there is no thread-boundary line loss (quirk Q1) and the
// numbering is deterministic (quirk Q4): unitig ids follow @SQ header order,
// then first appearance, over unitigs with at least one hit; read keys get
// arbitrary dense ids.  Lines with fewer than three tokens or empty lines are
// undefined behaviour in the reference and are rejected here.
#pragma once

#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace komb {

struct Span {
    const char *p;
    uint32_t len;
};

class MappedFile {
   public:
    const char *data = nullptr;
    size_t size = 0;
    bool open(const std::string &path) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) return false;
        struct stat st;
        if (fstat(fd_, &st) != 0) return false;
        size = (size_t)st.st_size;
        if (size == 0) return true;
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd_, 0);   // pre-fault: parallel first-touch faults serialise on the mm lock
        if (m == MAP_FAILED) return false;
        madvise(m, size, MADV_SEQUENTIAL);
        data = static_cast<const char *>(m);
        return true;
    }
    ~MappedFile() {
        if (data) munmap(const_cast<char *>(data), size);
        if (fd_ >= 0) close(fd_);
    }

   private:
    int fd_ = -1;
};

struct SamTokens {
    std::vector<Span> keys;    // read key of every hit, file order
    std::vector<Span> rnames;  // unitig name of every hit, file order
    std::vector<Span> sq;      // @SQ SN: names, header order
};

// next token under strtok(.., "\t") rules inside [p, end): skips leading tabs
inline bool next_token(const char *&p, const char *end, Span &tok) {
    while (p < end && *p == '\t') ++p;
    if (p >= end) return false;
    const char *s = p;
    const char *t = static_cast<const char *>(memchr(p, '\t', (size_t)(end - p)));
    p = t ? t : end;
    tok.p = s;
    tok.len = (uint32_t)(p - s);
    return true;
}

inline void tokenise_sam(const MappedFile &f, int threads, SamTokens &out, const std::string &path) {
    const char *base = f.data;
    const size_t n = f.size;
    if (n == 0) return;
    threads = std::max(1, threads);
    // chunk boundaries on line starts
    std::vector<size_t> cut(threads + 1, n);
    cut[0] = 0;
    for (int t = 1; t < threads; ++t) {
        size_t pos = n / threads * t;
        if (pos <= cut[t - 1]) { cut[t] = cut[t - 1]; continue; }
        const char *nl = static_cast<const char *>(memchr(base + pos, '\n', n - pos));
        cut[t] = nl ? (size_t)(nl - base) + 1 : n;
    }
    std::vector<SamTokens> part(threads);
    std::vector<std::string> err(threads);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
        SamTokens &loc = part[t];
        const char *p = base + cut[t];
        const char *stop = base + cut[t + 1];
        // alignment records are rarely shorter than ~48 bytes: one allocation instead of a dozen doublings
        loc.keys.reserve((size_t)(stop - p) / 48 + 16);
        loc.rnames.reserve((size_t)(stop - p) / 48 + 16);
        while (p < stop) {
            const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(base + n - p)));
            if (!nl) break;  // unterminated final line: not processed (reference -t 1 behaviour)
            const char *line = p, *end = nl;
            p = nl + 1;
            if (end > line && end[-1] == '\r') { /* keep '\r' inside the last field like the reference */ }
            if (line == end) { if (err[t].empty()) err[t] = "empty line"; continue; }
            if (*line == '@') {
                if (end - line > 4 && line[1] == 'S' && line[2] == 'Q' && line[3] == '\t') {
                    const char *q = line + 4;
                    Span tok;
                    while (next_token(q, end, tok))
                        if (tok.len > 3 && tok.p[0] == 'S' && tok.p[1] == 'N' && tok.p[2] == ':') {
                            loc.sq.push_back(Span{tok.p + 3, tok.len - 3});
                            break;
                        }
                }
                continue;
            }
            const char *q = line;
            Span qname, skip, rname;
            if (!next_token(q, end, qname) || !next_token(q, end, skip) || !next_token(q, end, rname)) {
                if (err[t].empty()) err[t] = "line with fewer than 3 tab-separated fields";
                continue;
            }
            if (rname.len == 1 && rname.p[0] == '*') continue;
            // key = qname.substr(1, qname.find('/')): from index 1, as many chars as the index of '/'
            Span key{qname.p + 1, qname.len - 1};
            const char *sl = static_cast<const char *>(memchr(qname.p, '/', qname.len));
            if (sl) key.len = std::min<uint32_t>(key.len, (uint32_t)(sl - qname.p));
            loc.keys.push_back(key);
            loc.rnames.push_back(rname);
        }
    }
    for (int t = 0; t < threads; ++t)
        if (!err[t].empty()) throw std::runtime_error("malformed SAM " + path + ": " + err[t]);
    // concatenate the per-thread parts in file order; every thread copies its own part to its final offset
    std::vector<size_t> hit_off(threads + 1, out.keys.size()), sq_off(threads + 1, out.sq.size());
    for (int t = 0; t < threads; ++t) {
        hit_off[t + 1] = hit_off[t] + part[t].keys.size();
        sq_off[t + 1] = sq_off[t] + part[t].sq.size();
    }
    out.keys.resize(hit_off[threads]);
    out.rnames.resize(hit_off[threads]);
    out.sq.resize(sq_off[threads]);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
        std::copy(part[t].keys.begin(), part[t].keys.end(), out.keys.begin() + hit_off[t]);
        std::copy(part[t].rnames.begin(), part[t].rnames.end(), out.rnames.begin() + hit_off[t]);
        std::copy(part[t].sq.begin(), part[t].sq.end(), out.sq.begin() + sq_off[t]);
    }
}

// ---------------------------------------------------------------------------
// string interning: spans -> dense ids, sharded by hash so threads never share a
// table.  With `ordered`, ids follow first appearance in `items`.
// ---------------------------------------------------------------------------
inline uint64_t hash_bytes(const char *p, uint32_t len) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ ((uint64_t)len * 0xff51afd7ed558ccdull);
    while (len >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0xc2b2ae3d27d4eb4full;
        h ^= h >> 29;
        p += 8;
        len -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, p, len);
    h = (h ^ w) * 0x165667b19e3779f9ull;
    h ^= h >> 32;
    h *= 0xd6e8feb86659fd93ull;
    h ^= h >> 32;
    return h;
}

struct InternResult {
    std::vector<uint32_t> ids;          // id of every item
    std::vector<uint64_t> first_index;  // per id: index of the first item carrying it
    uint32_t n_distinct = 0;
};

inline InternResult intern_spans(const std::vector<Span> &items, int threads, bool ordered) {
    constexpr int kShardBits = 8, kShards = 1 << kShardBits;
    const size_t n = items.size();
    InternResult res;
    res.ids.resize(n);
    if (n == 0) return res;
    threads = std::max(1, threads);
    std::vector<uint64_t> hash(n);
    std::vector<std::vector<uint32_t>> count(threads, std::vector<uint32_t>(kShards, 0));
#pragma omp parallel num_threads(threads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
        for (size_t i = lo; i < hi; ++i) {
            hash[i] = hash_bytes(items[i].p, items[i].len);
            count[t][hash[i] >> (64 - kShardBits)]++;
        }
    }
    // shard-major layout, threads in order inside a shard => ascending item index inside every shard
    std::vector<size_t> shard_begin(kShards + 1, 0);
    std::vector<std::vector<size_t>> cursor(threads, std::vector<size_t>(kShards, 0));
    {
        size_t run = 0;
        for (int s = 0; s < kShards; ++s) {
            shard_begin[s] = run;
            for (int t = 0; t < threads; ++t) { cursor[t][s] = run; run += count[t][s]; }
        }
        shard_begin[kShards] = run;
    }
    std::vector<uint32_t> order(n);  // item indices grouped by shard (n < 2^32 items per call)
    if (n >= (1ull << 32)) throw std::runtime_error("more than 2^32 hits in one komb2 run are not supported");
#pragma omp parallel num_threads(threads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
        // note: count[] was filled with the same (t, range) mapping only if nt == threads
        for (size_t i = lo; i < hi; ++i) order[cursor[t][hash[i] >> (64 - kShardBits)]++] = (uint32_t)i;
    }
    // per shard: open addressing over (hash, bytes); local ids in order of first appearance
    std::vector<std::vector<uint32_t>> shard_first(kShards);  // local id -> first item index
    std::vector<uint32_t> local_id(n);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int s = 0; s < kShards; ++s) {
        const size_t b = shard_begin[s], e = shard_begin[s + 1];
        if (b == e) continue;
        size_t cap = 16;
        while (cap < 2 * (e - b)) cap <<= 1;
        std::vector<uint32_t> slot(cap, UINT32_MAX);  // holds local id
        std::vector<uint32_t> &first = shard_first[s];
        for (size_t k = b; k < e; ++k) {
            const uint32_t i = order[k];
            size_t pos = (size_t)(hash[i] * 0x9e3779b97f4a7c15ull >> 20) & (cap - 1);
            while (true) {
                const uint32_t id = slot[pos];
                if (id == UINT32_MAX) {
                    slot[pos] = (uint32_t)first.size();
                    local_id[i] = (uint32_t)first.size();
                    first.push_back(i);
                    break;
                }
                const uint32_t j = first[id];
                if (hash[j] == hash[i] && items[j].len == items[i].len && memcmp(items[j].p, items[i].p, items[i].len) == 0) {
                    local_id[i] = id;
                    break;
                }
                pos = (pos + 1) & (cap - 1);
            }
        }
    }
    // global numbering
    std::vector<uint32_t> shard_base(kShards + 1, 0);
    for (int s = 0; s < kShards; ++s) shard_base[s + 1] = shard_base[s] + (uint32_t)shard_first[s].size();
    const uint32_t d = shard_base[kShards];
    res.n_distinct = d;
    std::vector<uint32_t> remap;  // provisional id (shard_base + local) -> final id
    res.first_index.resize(d);
    if (ordered) {
        // final id = rank of the id's first item among all first items: flag the first items, prefix-sum the flags
        // over the item array in parallel (no sort of the distinct ids)
        std::vector<uint8_t> is_first(n, 0);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int s = 0; s < kShards; ++s)
            for (uint32_t i : shard_first[s]) is_first[i] = 1;
        std::vector<uint32_t> chunk_base(threads + 1, 0);
        std::vector<uint32_t> rank_of_item(n);   // valid where is_first
#pragma omp parallel num_threads(threads)
        {
            const int t = omp_get_thread_num(), nt = omp_get_num_threads();
            const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
            uint32_t c = 0;
            for (size_t i = lo; i < hi; ++i) c += is_first[i];
            chunk_base[t + 1] = c;
#pragma omp barrier
#pragma omp single
            for (int k = 0; k < nt; ++k) chunk_base[k + 1] += chunk_base[k];
            uint32_t r = chunk_base[t];
            for (size_t i = lo; i < hi; ++i)
                if (is_first[i]) rank_of_item[i] = r++;
        }
        remap.resize(d);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int s = 0; s < kShards; ++s)
            for (size_t l = 0; l < shard_first[s].size(); ++l) {
                const uint32_t r = rank_of_item[shard_first[s][l]];
                remap[shard_base[s] + l] = r;
                res.first_index[r] = shard_first[s][l];
            }
    } else {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int s = 0; s < kShards; ++s)
            for (size_t l = 0; l < shard_first[s].size(); ++l) res.first_index[shard_base[s] + l] = shard_first[s][l];
    }
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const uint32_t prov = shard_base[hash[i] >> (64 - kShardBits)] + local_id[i];
        res.ids[i] = ordered ? remap[prov] : prov;
    }
    return res;
}

}  // namespace komb
