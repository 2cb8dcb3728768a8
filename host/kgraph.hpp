// kgraph.hpp — komb::Kgraph for the B200 path: the reference's class and method
// names (reference src/graph.h:41-63) kept as the host-side interface, every
// compute step forwarded to libkombgpu through the C ABI (include/kombgpu.h).
//
//   reference method (src/graph.cpp)          here
//   readSAM            :166-257               tokenise on the host -> integer hits + name table
//   getEdgeInfo        :259-285               nothing left to do: the mate union happens on the
//                                             device when both files' hits are sorted together
//   generateGraph      :287-393               kombgpu_build_graph (pairs, dedup, CSR)
//   readEdgeList       :395-453               writes edgelist.txt (simple edges, quirk Q7), prints
//                                             GraphInfo, calls runCore
//   runCore            :455-484               degree / coreness (kombgpu_graph_results) -> kcore.tsv
//   anomalyDetection   :637-648               CORE-A scores (kombgpu_graph_results) -> CoreA_anomaly.txt
//                       (CombineCoreA::run, src/CombineCoreA.h:16-43)
// Errors follow the reference: a message on stderr and exit(EXIT_FAILURE)
// (src/graph.cpp:58-61).
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "kombgpu.h"
#include "sam_tokenizer.hpp"

namespace komb {

struct HitTable {
    SamTokens tokens;                 // spans into the mapped SAM files
    std::vector<uint32_t> read_key;   // per hit
    std::vector<uint32_t> unitig;     // per hit (vid)
    std::vector<std::string> names;   // vid -> unitig name
};

// KOMB_TIMING=1: fine-grained wall-clock marks on stderr (where the drop-in's time goes outside the stage lines)
inline void timing_mark(const char *what) {
    static const bool on = getenv("KOMB_TIMING") != nullptr;
    static const auto t0 = std::chrono::steady_clock::now();
    static auto last = t0;
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[komb2 timing] %-28s +%.3f s (at %.3f s)\n", what,
            std::chrono::duration_cast<std::chrono::microseconds>(now - last).count() / 1e6,
            std::chrono::duration_cast<std::chrono::microseconds>(now - t0).count() / 1e6);
    last = now;
}

inline double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1000000.0;
}

// Page-locked host array for the large result buffers (DMA at full PCIe rate).
template <typename T>
class PinnedArray {
    kombgpu_ctx *ctx_;
    T *p_ = nullptr;
    size_t n_ = 0;

   public:
    PinnedArray(kombgpu_ctx *ctx, size_t n) : ctx_(ctx), n_(n) {
        void *q = nullptr;
        if (kombgpu_pinned_alloc(ctx, (uint64_t)(n ? n : 1) * sizeof(T), &q) != KOMBGPU_OK) {
            std::cerr << "komb2: " << kombgpu_last_error(ctx) << std::endl;
            exit(EXIT_FAILURE);
        }
        p_ = static_cast<T *>(q);
    }
    PinnedArray(const PinnedArray &) = delete;
    PinnedArray &operator=(const PinnedArray &) = delete;
    ~PinnedArray() { kombgpu_pinned_free(ctx_, p_); }
    T *data() { return p_; }
    T &operator[](size_t i) { return p_[i]; }
    size_t size() const { return n_; }
};

// ---- fast text formatting ------------------------------------------------------
inline char *put_u64(char *p, uint64_t v) {
    char tmp[24];
    int k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) *p++ = tmp[--k];
    return p;
}

// format blocks of rows [lo, hi) with `fmt_block(p, lo, hi)` into per-thread buffers, then write them in order
template <typename BlockFn>
inline void write_row_blocks(FILE *f, size_t n_rows, size_t max_row_bytes, int threads, BlockFn fmt_block) {
    threads = std::max(1, threads);
    const size_t block = 1 << 16;
    std::vector<std::vector<char>> buf(threads);
    for (size_t base = 0; base < n_rows; base += block * threads) {
        std::vector<size_t> used(threads, 0);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
        for (int t = 0; t < threads; ++t) {
            size_t lo = base + block * t, hi = std::min(n_rows, lo + block);
            if (lo >= hi) continue;
            buf[t].resize((hi - lo) * max_row_bytes);
            char *p = fmt_block(buf[t].data(), lo, hi);
            used[t] = (size_t)(p - buf[t].data());
        }
        for (int t = 0; t < threads; ++t)
            if (used[t]) fwrite(buf[t].data(), 1, used[t], f);
    }
}
template <typename RowFn>
inline void write_rows(FILE *f, size_t n_rows, size_t max_row_bytes, int threads, RowFn fmt_row) {
    write_row_blocks(f, n_rows, max_row_bytes, threads, [&](char *p, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) p = fmt_row(p, i);
        return p;
    });
}

class Kgraph {
    uint32_t _threads;
    uint64_t _readlength;
    kombgpu_ctx *_ctx = nullptr;
    kombgpu_graph *_graph = nullptr;
    kombgpu_hits *_hits = nullptr;       // device-resident hits when the SAM text is tokenised on the GPU
    bool _gpu_tokenise = true;           // KOMB_TOKENIZE=host keeps the host tokeniser (always used with several GPUs)
    bool _device_output = true;          // KOMB_OUTPUT=host formats the three files on the host (always without device hits)
    std::unique_ptr<PinnedArray<char>> _text[3];   // the files' bytes, formatted on the device
    uint64_t _text_bytes[3] = {0, 0, 0};
    std::vector<const MappedFile *> _sam_files;
    int _key_mode = KOMBGPU_KEY_REF32;
    int _device = 0;
    std::future<int> _ctx_ready;
    std::string _ctx_error;
    // results of the whole path, fetched in one overlapped call (kombgpu_graph_results) by readEdgeList
    std::unique_ptr<PinnedArray<uint64_t>> _efp;   // edge list in CSR form: forward offsets per source + targets
    std::unique_ptr<PinnedArray<uint32_t>> _ev;
    // multi-GPU (KOMB_GPU_DEVICES=0,1,...): one rank per device, a host thread per rank; each rank's share of the results
    struct RankOut {
        uint32_t v_lo = 0, n_local = 0;
        uint64_t n_fwd = 0;
        std::vector<uint64_t> fwd_ptr;
        std::vector<uint32_t> v;
        std::vector<int32_t> deg, core;
        std::vector<double> score;
        int rc = KOMBGPU_OK;
        std::string err;
    };
    std::vector<int> _devices;
    std::vector<RankOut> _ranks;
    uint64_t _m_multi = 0;
    int32_t _max_core_multi = 0;
    double _max_score_multi = 0.0;
    bool multi() const { return _devices.size() > 1; }
    std::unique_ptr<PinnedArray<int32_t>> _deg, _core;
    std::unique_ptr<PinnedArray<double>> _score;

    // exit() must not run CUDA's teardown under a context that is still being created
    void settleDevice() { if (_ctx_ready.valid()) _ctx_ready.wait(); }
    [[noreturn]] void fileNotFoundError(const std::string &path) {
        std::cerr << "File " << path << " could not be opened. Exiting..." << std::endl;
        settleDevice();
        exit(EXIT_FAILURE);
    }
    [[noreturn]] void gpuError(const char *what, int rc) {
        std::cerr << "komb2: " << what << " failed (" << rc << "): " << kombgpu_last_error(_ctx) << std::endl;
        exit(EXIT_FAILURE);
    }

   public:
    Kgraph(uint32_t threads, uint64_t readlength, int device, int key_mode, std::vector<int> devices = {})
        : _threads(threads ? threads : 1), _readlength(readlength), _key_mode(key_mode), _device(device), _devices(std::move(devices)) {
        const char *tok_env = getenv("KOMB_TOKENIZE");
        _gpu_tokenise = !multi() && !(tok_env && strcmp(tok_env, "host") == 0);
        const char *out_env = getenv("KOMB_OUTPUT");
        _device_output = _gpu_tokenise && !(out_env && strcmp(out_env, "host") == 0);
        if (multi()) { timing_mark("start"); return; }   // every rank's thread creates its own context (runMulti)
        // Creating the CUDA context costs seconds on a box without the persistence daemon (1.9 - 4.1 s measured on
        // the B200 boxes, against 0.75 s for everything else on a 2.5 M-hit SAM pair): do it on a helper thread
        // while the SAM files are tokenised and interned on the host; generateGraph is the first to need it.
        timing_mark("start");
        _ctx_ready = std::async(std::launch::async, [this]() {
            int rc = kombgpu_ctx_create(_device, &_ctx);
            if (rc != KOMBGPU_OK) _ctx_error = kombgpu_last_error(nullptr);   // thread-local message: fetch it here
            return rc;
        });
        (void)_readlength;  // stored and never read, like the reference (src/graph.cpp:55)
    }
    ~Kgraph() {
        if (_ctx_ready.valid()) _ctx_ready.wait();
        _efp.reset(); _ev.reset(); _deg.reset(); _core.reset(); _score.reset();  // pinned buffers go before the context
        for (auto &t : _text) t.reset();
        if (_graph) kombgpu_graph_destroy(_graph);
        if (_hits) kombgpu_hits_destroy(_hits);
        if (_ctx) kombgpu_ctx_destroy(_ctx);
    }

    // joins the helper thread that creates the CUDA context
    void waitForDevice() {
        if (!_ctx_ready.valid()) return;
        const int rc = _ctx_ready.get();
        timing_mark("kombgpu_ctx_create (joined)");
        if (rc != KOMBGPU_OK) {
            std::cerr << "komb2: cannot use CUDA device " << _device << ": " << _ctx_error << std::endl;
            exit(EXIT_FAILURE);
        }
    }

    // Tokenise one SAM file; the mapped file must outlive `hits` (spans point into it).
    void readSAM(const std::string &samfile, MappedFile &file, HitTable &hits, bool /*fulgor: ignored like the reference*/) {
        if (!file.open(samfile)) fileNotFoundError(samfile);
        if (_gpu_tokenise) { _sam_files.push_back(&file); return; }   // the bytes go to the device as they are (getEdgeInfo)
        try {
            tokenise_sam(file, (int)_threads, hits.tokens, samfile);
        } catch (const std::exception &e) {
            std::cerr << e.what() << std::endl;
            settleDevice();
            exit(EXIT_FAILURE);
        }
    }

    // Both mates are in `hits.tokens` now: intern read keys and unitig names.  The union of the two
    // mates' unitig sets per read (reference getEdgeInfo) needs no host work: equal keys get equal ids.
    void getEdgeInfo(HitTable &hits) {
        if (_gpu_tokenise) { tokeniseOnDevice(hits); return; }
        timing_mark("tokenise SAMs");
        const size_t h = hits.tokens.keys.size();
        // ordered: ids follow first appearance, so the two mate files arrive as (at most) two runs of non-decreasing
        // read ids and the device build merges them instead of radix-sorting the hits (build.cu, merge path)
        InternResult keys = intern_spans(hits.tokens.keys, (int)_threads, true);
        hits.read_key.swap(keys.ids);
        // unitig ids: @SQ order first, then first appearance; only unitigs with a hit become vertices
        std::vector<Span> items;
        items.reserve(hits.tokens.sq.size() + h);
        items.insert(items.end(), hits.tokens.sq.begin(), hits.tokens.sq.end());
        items.insert(items.end(), hits.tokens.rnames.begin(), hits.tokens.rnames.end());
        const size_t n_sq = hits.tokens.sq.size();
        InternResult names = intern_spans(items, (int)_threads, true);
        std::vector<uint8_t> has_hit(names.n_distinct, 0);
        for (size_t i = n_sq; i < items.size(); ++i) has_hit[names.ids[i]] = 1;
        std::vector<uint32_t> vid(names.n_distinct, UINT32_MAX);
        uint32_t next = 0;
        for (uint32_t d = 0; d < names.n_distinct; ++d)
            if (has_hit[d]) {
                vid[d] = next++;
                const Span &s = items[names.first_index[d]];
                hits.names.emplace_back(s.p, s.len);
            }
        hits.unitig.resize(h);
#pragma omp parallel for num_threads(_threads) schedule(static)
        for (size_t i = 0; i < h; ++i) hits.unitig[i] = vid[names.ids[n_sq + i]];
        timing_mark("intern keys + names");
    }

    // readSAM's tokeniser and the interning of read keys and unitig names, on the device (kombgpu_sam_parse): the host
    // keeps only the unitig names, as spans of the mapped files.
    void tokeniseOnDevice(HitTable &hits) {
        timing_mark("map SAMs");
        waitForDevice();
        std::vector<const char *> texts;
        std::vector<uint64_t> sizes;
        for (const MappedFile *f : _sam_files) { texts.push_back(f->data); sizes.push_back(f->size); }
        int rc = kombgpu_sam_parse(_ctx, texts.data(), sizes.data(), (int)texts.size(), &_hits);
        if (rc == KOMBGPU_EINVAL) {   // malformed input: the host tokeniser's message and exit code
            std::cerr << kombgpu_last_error(_ctx) << std::endl;
            exit(EXIT_FAILURE);
        }
        if (rc != KOMBGPU_OK) gpuError("kombgpu_sam_parse", rc);
        uint32_t n = 0;
        kombgpu_hits_counts(_hits, nullptr, nullptr, &n, nullptr);
        if (getenv("KOMB_TIMING")) {
            float up = 0.f, parse = 0.f;
            uint64_t launches = 0;
            kombgpu_hits_timing(_hits, &up, &parse, &launches, nullptr);
            fprintf(stderr, "[komb2 timing]   device tokeniser: upload %.3f ms, parse + intern %.3f ms, %llu kernels\n", up, parse,
                    (unsigned long long)launches);
        }
        timing_mark("kombgpu_sam_parse");
        if (_device_output) return;   // kcore.tsv is formatted on the device from its copy of the text: no names on the host
        std::vector<uint32_t> file(n), len(n);
        std::vector<uint64_t> off(n);
        rc = kombgpu_hits_names(_hits, file.data(), off.data(), len.data());
        if (rc != KOMBGPU_OK) gpuError("kombgpu_hits_names", rc);
        hits.names.resize(n);
#pragma omp parallel for num_threads(_threads) schedule(static)
        for (size_t i = 0; i < (size_t)n; ++i) hits.names[i].assign(_sam_files[file[i]]->data + off[i], len[i]);
        timing_mark("unitig names to the host");
    }

    // The whole device phase over several GPUs: hits split by read range (a read's hits stay together), one host
    // thread per rank; the ranks talk through peer memory inside libkombgpu (include/kombgpu.h, peer-memory path).
    void runMulti(HitTable &hits) {
        const int world = (int)_devices.size();
        const uint32_t n = (uint32_t)hits.names.size();
        const size_t h = hits.read_key.size();
        uint32_t n_reads = 0;
        for (size_t i = 0; i < h; ++i) n_reads = std::max(n_reads, hits.read_key[i] + 1);
        _ranks.assign(world, RankOut());
        void *group_slot = nullptr;
        // phase 1: the contexts (seconds each on a cold driver, so in parallel); nobody enters a collective before all exist
        std::vector<kombgpu_ctx *> ctxs(world, nullptr);
        {
            std::vector<std::thread> tc;
            for (int r = 0; r < world; ++r)
                tc.emplace_back([&, r]() {
                    const int rc = kombgpu_ctx_create(_devices[r], &ctxs[r]);
                    if (rc != KOMBGPU_OK) { _ranks[r].rc = rc; _ranks[r].err = std::string("kombgpu_ctx_create: ") + kombgpu_last_error(nullptr); }
                });
            for (auto &t : tc) t.join();
            bool ok = true;
            for (int r = 0; r < world; ++r)
                if (_ranks[r].rc != KOMBGPU_OK) {
                    std::cerr << "komb2: cannot use CUDA device " << _devices[r] << ": " << _ranks[r].err << std::endl;
                    ok = false;
                }
            if (!ok) {
                for (auto *c : ctxs) if (c) kombgpu_ctx_destroy(c);
                exit(EXIT_FAILURE);
            }
        }
        timing_mark("kombgpu_ctx_create (all devices)");
        std::vector<std::thread> th;
        for (int r = 0; r < world; ++r)
            th.emplace_back([&, r]() {
                RankOut &o = _ranks[r];
                kombgpu_ctx *ctx = ctxs[r];
                kombgpu_comm *comm = nullptr;
                kombgpu_dist_graph *g = nullptr;
                auto fail = [&](int rc, const char *what) {
                    o.rc = rc;
                    o.err = std::string(what) + ": " + kombgpu_last_error(ctx);
                    if (comm) kombgpu_comm_abort(comm);
                };
                int rc = kombgpu_comm_create_local(ctx, r, world, &group_slot, 0, &comm);
                if (rc != KOMBGPU_OK) fail(rc, "kombgpu_comm_create_local");
                if (rc == KOMBGPU_OK) {
                    // this rank's reads: ids [lo, hi)
                    const uint32_t lo = (uint32_t)((uint64_t)n_reads * r / world), hi = (uint32_t)((uint64_t)n_reads * (r + 1) / world);
                    std::vector<uint32_t> rk, ut;
                    for (size_t i = 0; i < h; ++i)
                        if (hits.read_key[i] >= lo && hits.read_key[i] < hi) { rk.push_back(hits.read_key[i]); ut.push_back(hits.unitig[i]); }
                    rc = kombgpu_dist_build_hits(comm, rk.data(), ut.data(), rk.size(), n, &g);
                    if (rc != KOMBGPU_OK) fail(rc, "kombgpu_dist_build_hits");
                }
                if (rc == KOMBGPU_OK && (rc = kombgpu_dist_coreness(g)) != KOMBGPU_OK) fail(rc, "kombgpu_dist_coreness");
                if (rc == KOMBGPU_OK && (rc = kombgpu_dist_corea(g, _key_mode)) != KOMBGPU_OK) fail(rc, "kombgpu_dist_corea");
                if (rc == KOMBGPU_OK) {
                    kombgpu_dist_stats st;
                    kombgpu_dist_graph_stats(g, &st);
                    o.v_lo = st.v_lo; o.n_local = st.n_local; o.n_fwd = st.n_fwd_local;
                    o.fwd_ptr.resize((size_t)o.n_local + 1); o.v.resize(o.n_fwd);
                    o.deg.resize(o.n_local); o.core.resize(o.n_local); o.score.resize(o.n_local);
                    rc = kombgpu_dist_graph_results(g, o.deg.data(), o.core.data(), o.score.data());
                    if (rc == KOMBGPU_OK) rc = kombgpu_dist_graph_edges_csr(g, o.fwd_ptr.data(), o.v.data());
                    if (rc != KOMBGPU_OK) fail(rc, "kombgpu_dist_graph_results");
                    if (r == 0) {
                        _m_multi = st.n_edges_global;
                        kombgpu_dist_graph_summary(g, &_max_core_multi, &_max_score_multi);
                    }
                }
                if (g) kombgpu_dist_graph_destroy(g);
                if (comm) kombgpu_comm_destroy(comm);
                if (ctx) kombgpu_ctx_destroy(ctx);
            });
        for (auto &t : th) t.join();
        for (int r = 0; r < world; ++r)
            if (_ranks[r].rc != KOMBGPU_OK) {
                std::cerr << "komb2: rank " << r << " (device " << _devices[r] << "): " << _ranks[r].err << " (" << _ranks[r].rc << ")" << std::endl;
                exit(EXIT_FAILURE);
            }
        timing_mark("multi-GPU build + peel + CORE-A + results");
    }

    void generateGraph(HitTable &hits) {
        if (multi()) { runMulti(hits); return; }
        waitForDevice();
        int rc = _hits ? kombgpu_build_graph_hits(_hits, &_graph)
                       : kombgpu_build_graph(_ctx, hits.read_key.data(), hits.unitig.data(), hits.read_key.size(),
                                             (uint32_t)hits.names.size(), &_graph);
        if (rc != KOMBGPU_OK) gpuError(_hits ? "kombgpu_build_graph_hits" : "kombgpu_build_graph", rc);
        timing_mark("kombgpu_build_graph");
    }

    void readEdgeList(const std::string &dir, const std::string &inputUnitigs, HitTable &hits) {
        const std::string edgelist_file = dir + "/edgelist.txt";
        FILE *ef = fopen(edgelist_file.c_str(), "w");
        if (ef == nullptr) fileNotFoundError(edgelist_file);
        auto begin_graph = std::chrono::steady_clock::now();
        uint32_t n = 0;
        uint64_t m = 0;
        if (multi()) {
            // the ranks' slices of the canonical edge list, in rank order, ARE the sorted list
            n = (uint32_t)hits.names.size();
            m = _m_multi;
            for (const RankOut &o : _ranks) {
                const uint64_t *fp = o.fwd_ptr.data();
                write_row_blocks(ef, o.n_fwd, 24, (int)_threads, [&](char *p, size_t lo, size_t hi) {
                    size_t x = (size_t)(std::upper_bound(fp, fp + o.n_local + 1, (uint64_t)lo) - fp) - 1;
                    for (size_t i = lo; i < hi; ++i) {
                        while (fp[x + 1] <= i) ++x;
                        p = put_u64(p, (uint64_t)o.v_lo + x); *p++ = '\t';
                        p = put_u64(p, o.v[i]); *p++ = '\n';
                    }
                    return p;
                });
            }
            fclose(ef);
            timing_mark("write edgelist.txt");
            fprintf(stdout, "\nTime elapsed for initializing igraph graph: %.3f s\n", seconds_since(begin_graph));
            fprintf(stdout, "\nTime elapsed for simplifying graph: %.3f s\n", 0.0);
            fprintf(stdout, "GraphInfo...\n\tNumber of vertices: %d\n", (int)n);
            fprintf(stdout, "\tNumber of edges: %d\n", (int)m);
            FILE *uf2 = fopen(inputUnitigs.c_str(), "r");
            if (uf2 == nullptr) fileNotFoundError(inputUnitigs);
            fclose(uf2);
            auto begin_kcore2 = std::chrono::steady_clock::now();
            runCore(dir, hits);
            fprintf(stdout, "\nTime elapsed doing K-core decomposition: %.3f s\n", seconds_since(begin_kcore2));
            return;
        }
        kombgpu_graph_counts(_graph, &n, &m);
        if (_device_output) {
            // the files are formatted on the device (kombgpu_graph_format): the edge list first, its download runs on the
            // copy stream under the peel and CORE-A; the host issues one write per file
            auto fetch = [&](int which, bool async) {
                int rc = kombgpu_graph_format(_graph, which, _hits, &_text_bytes[which]);
                if (rc != KOMBGPU_OK) gpuError("kombgpu_graph_format", rc);
                _text[which].reset(new PinnedArray<char>(_ctx, _text_bytes[which]));
                rc = kombgpu_graph_format_fetch(_graph, which, _text[which]->data(), async ? 1 : 0);
                if (rc != KOMBGPU_OK) gpuError("kombgpu_graph_format_fetch", rc);
            };
            fetch(KOMBGPU_FILE_EDGELIST, true);
            int rc = kombgpu_graph_analyse(_graph, _key_mode);
            if (rc != KOMBGPU_OK) gpuError("kombgpu_graph_analyse", rc);
            fetch(KOMBGPU_FILE_KCORE, true);
            fetch(KOMBGPU_FILE_COREA, true);
            if ((rc = kombgpu_graph_format_wait(_graph)) != KOMBGPU_OK) gpuError("kombgpu_graph_format_wait", rc);
            timing_mark("format on the device + download");
            if (_text_bytes[0] && fwrite(_text[0]->data(), 1, _text_bytes[0], ef) != _text_bytes[0]) fileNotFoundError(edgelist_file);
            fclose(ef);
            _text[0].reset();
            timing_mark("write edgelist.txt");
            fprintf(stdout, "\nTime elapsed for initializing igraph graph: %.3f s\n", seconds_since(begin_graph));
            fprintf(stdout, "\nTime elapsed for simplifying graph: %.3f s\n", 0.0);
            fprintf(stdout, "GraphInfo...\n\tNumber of vertices: %d\n", (int)n);
            fprintf(stdout, "\tNumber of edges: %d\n", (int)m);
            FILE *uf3 = fopen(inputUnitigs.c_str(), "r");
            if (uf3 == nullptr) fileNotFoundError(inputUnitigs);
            fclose(uf3);
            auto begin_kcore3 = std::chrono::steady_clock::now();
            runCore(dir, hits);
            fprintf(stdout, "\nTime elapsed doing K-core decomposition: %.3f s\n", seconds_since(begin_kcore3));
            return;
        }
        // one call fetches everything the three output files need; the edge-list download (CSR form: offsets per
        // source + targets, half the bytes of two id arrays) overlaps the peel
        _efp.reset(new PinnedArray<uint64_t>(_ctx, (size_t)n + 1));
        _ev.reset(new PinnedArray<uint32_t>(_ctx, m));
        _deg.reset(new PinnedArray<int32_t>(_ctx, n));
        _core.reset(new PinnedArray<int32_t>(_ctx, n));
        _score.reset(new PinnedArray<double>(_ctx, n));
        int rc = kombgpu_graph_results_csr(_graph, _key_mode, _efp->data(), _ev->data(), _deg->data(), _core->data(), _score->data());
        if (rc != KOMBGPU_OK) gpuError("kombgpu_graph_results_csr", rc);
        timing_mark("pinned alloc + results");
        const uint64_t *fp = _efp->data();
        PinnedArray<uint32_t> &v = *_ev;
        write_row_blocks(ef, m, 24, (int)_threads, [&](char *p, size_t lo, size_t hi) {
            size_t u = (size_t)(std::upper_bound(fp, fp + n + 1, (uint64_t)lo) - fp) - 1;   // source of edge lo
            for (size_t i = lo; i < hi; ++i) {
                while (fp[u + 1] <= i) ++u;
                p = put_u64(p, u); *p++ = '\t';
                p = put_u64(p, v[i]); *p++ = '\n';
            }
            return p;
        });
        fclose(ef);
        timing_mark("write edgelist.txt");
        fprintf(stdout, "\nTime elapsed for initializing igraph graph: %.3f s\n", seconds_since(begin_graph));
        fprintf(stdout, "\nTime elapsed for simplifying graph: %.3f s\n", 0.0);  // dedup is part of the device build
        fprintf(stdout, "GraphInfo...\n\tNumber of vertices: %d\n", (int)n);
        fprintf(stdout, "\tNumber of edges: %d\n", (int)m);
        // the reference opens and parses the unitig FASTA here and discards the result
        // (src/graph.cpp:446,565-589): keep its one observable effect, the open check
        FILE *uf = fopen(inputUnitigs.c_str(), "r");
        if (uf == nullptr) fileNotFoundError(inputUnitigs);
        fclose(uf);
        auto begin_kcore = std::chrono::steady_clock::now();
        runCore(dir, hits);
        fprintf(stdout, "\nTime elapsed doing K-core decomposition: %.3f s\n", seconds_since(begin_kcore));
    }

    void runCore(const std::string &dir, HitTable &hits) {
        const std::string kcore_file = dir + "/kcore.tsv";
        const uint32_t n = (uint32_t)hits.names.size();
        if (_device_output && !multi()) {
            FILE *kf = fopen(kcore_file.c_str(), "w+");
            if (kf == nullptr) fileNotFoundError(kcore_file);
            if (fwrite(_text[1]->data(), 1, _text_bytes[1], kf) != _text_bytes[1]) fileNotFoundError(kcore_file);
            fclose(kf);
            _text[1].reset();
            timing_mark("write kcore.tsv");
            return;
        }
        // fetched by readEdgeList (one GPU) or runMulti (every rank's slice)
        std::vector<int32_t> deg_all, core_all;
        if (multi()) {
            deg_all.resize(n); core_all.resize(n);
            for (const RankOut &o : _ranks) {
                std::copy(o.deg.begin(), o.deg.end(), deg_all.begin() + o.v_lo);
                std::copy(o.core.begin(), o.core.end(), core_all.begin() + o.v_lo);
            }
        }
        const int32_t *deg = multi() ? deg_all.data() : _deg->data(), *core = multi() ? core_all.data() : _core->data();
        FILE *kcf = fopen(kcore_file.c_str(), "w+");
        if (kcf == nullptr) fileNotFoundError(kcore_file);
        fprintf(kcf, "#VID\tName\tCoreness\tDegree\n");
        size_t max_name = 0;
        for (auto &s : hits.names) max_name = std::max(max_name, s.size());
        write_rows(kcf, n, max_name + 40, (int)_threads, [&](char *p, size_t i) {
            p = put_u64(p, i); *p++ = '\t';
            memcpy(p, hits.names[i].data(), hits.names[i].size()); p += hits.names[i].size(); *p++ = '\t';
            p = put_u64(p, (uint64_t)core[i]); *p++ = '\t';
            p = put_u64(p, (uint64_t)deg[i]); *p++ = '\n';
            return p;
        });
        fclose(kcf);
        timing_mark("write kcore.tsv");
    }

    void anomalyDetection(const std::string &dir, bool weight) {
        uint32_t n = 0;
        std::vector<double> score_all;
        int32_t max_core = 0;
        double max_score = 0.0;
        if (multi()) {
            for (const RankOut &o : _ranks) n += o.n_local;
            score_all.resize(n);
            for (const RankOut &o : _ranks) std::copy(o.score.begin(), o.score.end(), score_all.begin() + o.v_lo);
            max_core = _max_core_multi;
            max_score = _max_score_multi;
        } else {
            kombgpu_graph_counts(_graph, &n, nullptr);
            kombgpu_graph_summary(_graph, &max_core, &max_score);
        }
        const double *score = multi() ? score_all.data() : (_score ? _score->data() : nullptr);  // fetched by readEdgeList / runMulti
        const double dense_ratio = (double)(max_core / 2);  // integer division, like CombineCoreA.h:24
        fprintf(stdout, "Dense Ratio: %f\n", n ? dense_ratio : 0.0);
        if (weight) {
            fprintf(stdout, "Max CoreA score: %f\n", max_score);
            const std::string anomaly_output = dir + "/CoreA_anomaly.txt";
            FILE *fp = fopen(anomaly_output.c_str(), "w+");
            if (fp == nullptr) fileNotFoundError(anomaly_output);
            if (_device_output && !multi()) {
                if (_text_bytes[2] && fwrite(_text[2]->data(), 1, _text_bytes[2], fp) != _text_bytes[2]) fileNotFoundError(anomaly_output);
            } else
            write_rows(fp, n, 64, (int)_threads, [&](char *p, size_t i) {
                p = put_u64(p, i); *p++ = '\t';
                p += snprintf(p, 48, "%f\n", score[i]);
                return p;
            });
            fclose(fp);
            timing_mark("write CoreA_anomaly.txt");
        }
        _efp.reset(); _ev.reset(); _deg.reset(); _core.reset(); _score.reset();
        for (auto &t : _text) t.reset();
        _ranks.clear();
        if (_graph) kombgpu_graph_destroy(_graph);
        _graph = nullptr;
        timing_mark("free pinned + graph");
    }

    const kombgpu_graph *graph() const { return _graph; }
};

}  // namespace komb
