// ppeel.cu — stage 2 over the ranks of a communicator: k-core peel of a graph partitioned by unitig-id range, driven
// entirely from the device.
//
// Replaces igraph_coreness (src/graph.cpp:463) like peel.cu does on one GPU; the result is the same unique
// function of the graph, so it is bit-exact whatever the partition.
//
// One persistent kernel per GPU.  Every rank owns the working degrees of its own unitigs and nobody else touches
// them: a decrement of a unitig another rank owns travels as a MESSAGE -- its 32-bit id, stored straight into
// that rank's mailbox (peer memory over NVLink; one mailbox per ordered pair of ranks, sized by the number of
// CSR entries that cross that way, so it can never overflow and is never reused).  The peel advances in
// SUB-ROUNDS, one per cascade generation:
//
//   [scan]     first sub-round of a level k: local unitigs with degree == k form the frontier, the alive list is
//              compacted, the smallest surviving degree is noted (empty levels are skipped with it)
//   process    walk the rows of the frontier: local neighbours are decremented in place (the decrement that takes
//              a degree to k discovers that unitig for the next sub-round), remote neighbours become messages
//   exchange   ONE meeting of all ranks: each publishes, in every peer's control words, how many messages it has
//              sent there so far and how much it did this sub-round, tagged with the sub-round number, and waits
//              for the same from every peer (flags in peer memory; no host, no collective library)
//   apply      decrement the targets of the messages that arrived; discoveries join the next frontier
//
// A level ends when no rank discovered, sliced or sent anything in a sub-round.  Two things keep a sub-round
// cheap when cascades are thin (hundreds of dependent generations of a few unitigs each, the usual shape of a
// collapsing core): a rank whose share of the sub-round is small runs it SOLO, on CTA 0 alone, while its other
// CTAs wait on a local word -- no grid-wide barrier on the critical path -- and rows longer than kSliceLen are
// cut into slices that the whole grid shares in the next sub-round.
//
// Ranks that share one device (tests) run inside ONE cooperative grid, a group of CTAs per rank: kernels of
// different ranks must never wait for one another on the same GPU.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dgraph.cuh"

namespace kg {
namespace {

constexpr int kPThreads = 512;
constexpr int kPWarps = kPThreads / 32;
constexpr int kPU = 4;                       // independent edge chains per lane
constexpr uint32_t kSliceLen = 2048;         // rows longer than this are cut into slices of this many edges
constexpr uint32_t kSoloFront = 512;         // a sub-round with at most this many frontier unitigs ...
constexpr unsigned long long kSoloEdges = 16384;   // ... and this many edges to walk runs on CTA 0 alone
constexpr unsigned long long kSoloInbox = 8192;    // same for the messages to apply
constexpr int kScanItems = 4;
constexpr unsigned long long kPeelWatchdogNs = 20ull * 1000000000ull;
constexpr unsigned long long kTagShift = 40, kValMask = (1ull << 40) - 1;
// The leader CTA of a rank may run ahead of its other CTAs while sub-rounds are SOLO; its decisions are kept in a
// ring indexed by the sub-round number, and every kPlanSync sub-rounds all CTAs of the rank meet, so the lead
// stays below the ring size.
constexpr int kPlanRing = 256;
constexpr uint32_t kPlanSync = 64;

enum : uint32_t { kModeFull = 1, kModeSolo = 2, kFlagLevelOver = 4, kFlagDone = 8 };

// control words, one block per rank in symmetric memory: w[src][i] is written by rank src
//   0: messages src has sent here so far   1: work src did this sub-round (discoveries + slices + messages)
//   2: smallest surviving degree at src    3: frontier size at src
struct PeelCtl {
    unsigned long long w[kMaxRanks][4];
};

struct PRankState {
    unsigned long long bar_count, bar_gen;      // barrier of this rank's CTAs
    unsigned long long plan_a[kPlanRing], plan_b[kPlanRing];   // leader -> CTAs: (sub-round << 8) | mode / flags, slot = sub-round % ring
    int32_t k_next[kPlanRing];                  // level of the next sub-round when plan_b says the level is over
    unsigned long long sent[kMaxRanks];         // messages sent to rank p so far
    unsigned long long published[kMaxRanks];    // ... as of the last exchange
    unsigned long long recv_hi[kMaxRanks];      // messages from rank q that have arrived
    unsigned long long applied[kMaxRanks];      // ... and that have been applied
    unsigned long long front_edges[2];          // sum of the row lengths of the frontier lists
    unsigned long long work;                    // discoveries + slices + messages of the sub-round in progress
    unsigned long long n_peeled;
    unsigned long long msg_sent_total, msg_recv_total;
    uint32_t front_cnt[2], slice_cnt[2];
    uint32_t alive_out[2];                      // survivors written by a scan (slot = index of the list it wrote)
    int32_t local_min;
    int32_t max_core;
    uint32_t levels, subrounds, solo_subrounds;
    uint32_t error;                             // 1 watchdog (CTAs), 2 watchdog (peers), 3 mailbox overflow, 4 bad message, 5 list overflow
};

struct PRank {
    uint32_t n_local, v_lo, step;
    int world, rank;
    uint32_t ctas;                              // CTAs that work for this rank
    const uint64_t *row_ptr;
    const uint32_t *col;
    int32_t *deg;                               // working degrees (never clamped: only ever decremented)
    int32_t *core;
    uint32_t *alive[2];
    uint32_t *front[2];
    uint64_t *slices[2];                        // first_edge << 12 | length (length <= kSliceLen)
    uint32_t slice_cap;
    PeelCtl *ctl_local;
    PeelCtl *ctl_peer[kMaxRanks];
    const uint32_t *mbox_in[kMaxRanks];         // messages from rank q (local memory)
    uint32_t *mbox_out[kMaxRanks];              // rank p's mailbox for this rank (peer memory)
    unsigned long long mbox_out_cap[kMaxRanks];
    PRankState *st;
};

__device__ __forceinline__ unsigned long long ld_acq_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acq_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pglobal_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// barrier of one rank's CTAs; the fence is system-wide because messages stored into peer memory before it must be
// visible at the peer before the leader publishes this rank's counts after it
__device__ __forceinline__ void rank_barrier(PRankState *st, uint32_t ctas, unsigned long long &gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        ++gen;
        const unsigned long long arrived = atomicAdd(&st->bar_count, 1ull) + 1ull;
        if (arrived == gen * ctas) {
            st_rel_gpu(&st->bar_gen, gen);
        } else {
            const unsigned long long t0 = pglobal_ns();
            uint32_t spins = 0;
            while (ld_acq_gpu(&st->bar_gen) < gen) {
                if ((++spins & 4095u) == 0 && (*(volatile uint32_t *)&st->error || pglobal_ns() - t0 > kPeelWatchdogNs)) {
                    atomicCAS(&st->error, 0u, 1u);
                    break;
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// every CTA learns what the leader decided for sub-round t
__device__ __forceinline__ uint32_t wait_plan(unsigned long long *word, uint32_t t, PRankState *st, uint32_t *s_bcast) {
    if (threadIdx.x == 0) {
        const unsigned long long t0 = pglobal_ns();
        uint32_t spins = 0, res = 0;
        while (true) {
            const unsigned long long w = ld_acq_gpu(word);
            if ((uint32_t)(w >> 8) == t) { res = (uint32_t)(w & 0xffu); break; }
            if ((++spins & 4095u) == 0 && (*(volatile uint32_t *)&st->error || pglobal_ns() - t0 > kPeelWatchdogNs)) {
                atomicCAS(&st->error, 0u, 1u);
                res = kFlagDone | kModeSolo;
                break;
            }
        }
        *s_bcast = res;
    }
    __syncthreads();
    const uint32_t r = *s_bcast;
    __syncthreads();
    return r;
}

// a newly discovered unitig (its degree just reached k): coreness k, member of the next frontier
// (count_work: discoveries of the process stage keep the level open; those of the apply stage are implied by the
// messages that caused them, which were counted by their sender)
__device__ __forceinline__ void discover(const PRank &R, PRankState *st, bool found, uint32_t loc, int32_t k, uint32_t nxt,
                                         bool count_work) {
    const uint32_t lane = lane_id();
    const uint32_t fm = __ballot_sync(kFullMask, found);
    if (fm == 0) return;
    uint32_t len = 0;
    if (found) {
        R.core[loc] = k;
        len = (uint32_t)(R.row_ptr[loc + 1] - R.row_ptr[loc]);
    }
    const uint32_t tot_len = warp_reduce_add(len);
    uint32_t pos = 0;
    if (lane == 0) {
        pos = atomicAdd(&st->front_cnt[nxt], (uint32_t)__popc(fm));
        atomicAdd(&st->front_edges[nxt], (unsigned long long)tot_len);
        if (count_work) atomicAdd(&st->work, (unsigned long long)__popc(fm));
    }
    pos = __shfl_sync(kFullMask, pos, 0) + __popc(fm & lanemask_lt());
    if (found) {
        if (pos < R.n_local) R.front[nxt][pos] = loc;
        else atomicCAS(&st->error, 0u, 5u);
    }
}

// one neighbour per lane: decrement it here, or send the decrement to its owner
__device__ __forceinline__ void visit(const PRank &R, PRankState *st, bool valid, uint32_t u, int32_t k, uint32_t nxt, bool look_first) {
    const uint32_t lane = lane_id();
    const uint32_t owner = valid ? min(u / R.step, (uint32_t)R.world - 1u) : 0xffffffffu;
    const bool local = valid && owner == (uint32_t)R.rank;
    const bool remote = valid && !local;
    // ---- remote: one message per neighbour, a warp's messages to one rank are stored as one run
    if (__ballot_sync(kFullMask, remote)) {
        const uint32_t same = __match_any_sync(kFullMask, remote ? owner : 0xffffffffu);
        const uint32_t lead = (uint32_t)__ffs(same) - 1u;
        unsigned long long base = 0;
        if (remote && lane == lead) base = atomicAdd(&st->sent[owner], (unsigned long long)__popc(same));
        base = __shfl_sync(kFullMask, base, lead);
        if (remote) {
            const unsigned long long pos = base + (unsigned long long)__popc(same & lanemask_lt());
            if (pos < R.mbox_out_cap[owner]) R.mbox_out[owner][pos] = u;
            else atomicCAS(&st->error, 0u, 3u);
        }
    }
    // ---- local
    bool found = false;
    const uint32_t loc = u - R.v_lo;
    if (local) {
        int32_t d = look_first ? __ldcg(&R.deg[loc]) : INT32_MAX;
        if (d > k) d = atomicSub(&R.deg[loc], 1);
        found = d == k + 1;
    }
    discover(R, st, found, loc, k, nxt, true);
}

// walk edges [0, total) of a batch of rows, one row per lane (row_begin, excl prefix of the lengths); all lanes call it
__device__ __forceinline__ void walk_rows(const PRank &R, PRankState *st, uint64_t row_begin, uint32_t excl, uint32_t total, int32_t k,
                                          uint32_t nxt) {
    const uint32_t lane = lane_id();
    const uint32_t row_lo = (uint32_t)row_begin, row_hi = (uint32_t)(row_begin >> 32);
    const bool look_first = total > 32u * kPU;
    for (uint32_t base = 0; base < total; base += 32u * kPU) {
        uint32_t u[kPU];
        bool valid[kPU];
#pragma unroll
        for (int t = 0; t < kPU; ++t) {
            const uint32_t e = base + t * 32u + lane;
            valid[t] = false;
            u[t] = 0;
            if (base + t * 32u >= total) continue;   // warp-uniform
            uint32_t j = 0;   // owner row: the last lane j with excl[j] <= e
#pragma unroll
            for (uint32_t s = 16; s > 0; s >>= 1) {
                const uint32_t x = __shfl_sync(kFullMask, excl, (j + s) & 31u);
                if (j + s < 32u && x <= e) j += s;
            }
            const uint32_t ex_j = __shfl_sync(kFullMask, excl, j);
            const uint32_t lo = __shfl_sync(kFullMask, row_lo, j), hi = __shfl_sync(kFullMask, row_hi, j);
            if (e < total) { u[t] = R.col[(((uint64_t)hi << 32) | lo) + (e - ex_j)]; valid[t] = true; }
        }
#pragma unroll
        for (int t = 0; t < kPU; ++t) {
            if (base + t * 32u >= total) continue;
            visit(R, st, valid[t], u[t], k, nxt, look_first);
        }
    }
}

// PROCESS stage for the warps [w0, w0 + nw) of this rank (global warp index gw): slices first, then the frontier
__device__ __forceinline__ void process_stage(const PRank &R, PRankState *st, uint32_t gw, uint32_t nw, uint32_t cur, int32_t k) {
    const uint32_t lane = lane_id();
    const uint32_t nxt = cur ^ 1u;
    const uint32_t n_sl = __ldcg(&st->slice_cnt[cur]);
    for (uint32_t i = gw; i < n_sl; i += nw) {
        const uint64_t sl = __ldcg(&R.slices[cur][i]);
        const uint64_t first = sl >> 12;
        const uint32_t len = (uint32_t)(sl & 0xfffu) + 1u;
        walk_rows(R, st, first, lane == 0 ? 0u : len, len, k, nxt);
    }
    const uint32_t n_f = __ldcg(&st->front_cnt[cur]);
    uint32_t peeled = 0;
    for (uint32_t c = gw; (uint64_t)c * 32u < n_f; c += nw) {
        const uint32_t i = c * 32u + lane;
        uint64_t row = 0;
        uint32_t len = 0;
        if (i < n_f) {
            const uint32_t v = __ldcg(&R.front[cur][i]);
            row = R.row_ptr[v];
            len = (uint32_t)(R.row_ptr[v + 1] - row);
            ++peeled;
        }
        // long rows are cut into slices that the whole grid walks in the next sub-round
        const uint32_t n_cut = len > kSliceLen ? (len + kSliceLen - 1) / kSliceLen : 0u;
        if (__ballot_sync(kFullMask, n_cut != 0)) {
            const uint32_t inc = warp_incl_scan_add(n_cut);
            const uint32_t tot = __shfl_sync(kFullMask, inc, 31);
            uint32_t pos = 0;
            if (lane == 0) {
                pos = atomicAdd(&st->slice_cnt[nxt], tot);
                atomicAdd(&st->work, (unsigned long long)tot);
            }
            pos = __shfl_sync(kFullMask, pos, 0) + (inc - n_cut);
            for (uint32_t s = 0; s < n_cut; ++s) {
                const uint32_t l = min(kSliceLen, len - s * kSliceLen);
                if (pos + s < R.slice_cap) R.slices[nxt][pos + s] = ((row + (uint64_t)s * kSliceLen) << 12) | (uint64_t)(l - 1u);
                else atomicCAS(&st->error, 0u, 5u);
            }
            if (n_cut) len = 0;
        }
        const uint32_t incl = warp_incl_scan_add(len);
        walk_rows(R, st, row, incl - len, __shfl_sync(kFullMask, incl, 31), k, nxt);
    }
    peeled = warp_reduce_add(peeled);
    if (lane == 0 && peeled) atomicAdd(&st->n_peeled, (unsigned long long)peeled);
}

// APPLY stage: the messages that arrived since the last sub-round, spread over the warps [.., nw) of this rank
__device__ __forceinline__ void apply_stage(const PRank &R, PRankState *st, uint32_t gw, uint32_t nw, uint32_t cur, int32_t k) {
    const uint32_t lane = lane_id();
    const uint32_t nxt = cur ^ 1u;
    for (int q = 0; q < R.world; ++q) {
        const unsigned long long lo = __ldcg(&st->applied[q]), hi = __ldcg(&st->recv_hi[q]);
        for (unsigned long long base = lo + (unsigned long long)gw * 32ull; base < hi; base += (unsigned long long)nw * 32ull) {
            const unsigned long long i = base + lane;
            bool found = false;
            uint32_t loc = 0;
            if (i < hi) {
                loc = __ldcg(&R.mbox_in[q][i]) - R.v_lo;
                if (loc >= R.n_local) {
                    atomicCAS(&st->error, 0u, 4u);
                } else {
                    int32_t d = __ldcg(&R.deg[loc]);
                    if (d > k) d = atomicSub(&R.deg[loc], 1);
                    found = d == k + 1;
                }
            }
            discover(R, st, found, loc, k, nxt, false);
        }
    }
}

// SCAN stage of level k (all CTAs of the rank): frontier = alive unitigs at degree k, survivors compacted
__device__ __forceinline__ void scan_stage(const PRank &R, PRankState *st, uint32_t cta, const uint32_t *alive_src, uint32_t n_alive,
                                           uint32_t *alive_dst, uint32_t *alive_out, uint32_t cur, int32_t k, uint32_t *s_scan,
                                           uint32_t *s_base) {
    const uint32_t tid = threadIdx.x;
    int32_t local_min = INT32_MAX;
    unsigned long long edges = 0;
    uint32_t zero_deg = 0;
    const uint32_t tile = kPThreads * kScanItems;
    for (uint64_t t0 = (uint64_t)cta * tile; t0 < n_alive; t0 += (uint64_t)R.ctas * tile) {
        uint32_t v[kScanItems], flag[kScanItems];
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint64_t i = t0 + (uint64_t)j * kPThreads + tid;
            flag[j] = 0;
            v[j] = 0;
            if (i < n_alive) {
                v[j] = alive_src ? __ldcg(&alive_src[i]) : (uint32_t)i;
                const int32_t d = __ldcg(&R.deg[v[j]]);
                if (d == k) {
                    R.core[v[j]] = k;
                    if (k > 0) { flag[j] = 1u; edges += (unsigned long long)(R.row_ptr[v[j] + 1] - R.row_ptr[v[j]]); }
                    else ++zero_deg;   // no row to walk
                } else if (d > k) {
                    flag[j] = 0x10000u;
                    local_min = min(local_min, d);
                }
            }
            mine += flag[j];
        }
        uint32_t total = 0;
        const uint32_t ex = block_excl_scan_add<uint32_t, kPThreads>(mine, s_scan, &total);
        if (tid == 0) {
            const uint32_t nf = total & 0xffffu, ns = total >> 16;
            s_base[0] = nf ? atomicAdd(&st->front_cnt[cur], nf) : 0;
            s_base[1] = ns ? atomicAdd(alive_out, ns) : 0;
        }
        __syncthreads();
        uint32_t fpos = s_base[0] + (ex & 0xffffu), spos = s_base[1] + (ex >> 16);
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            if (flag[j] == 1u) R.front[cur][fpos++] = v[j];
            else if (flag[j]) alive_dst[spos++] = v[j];
        }
        __syncthreads();
    }
    local_min = warp_reduce_min(local_min);
    zero_deg = warp_reduce_add(zero_deg);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) edges += __shfl_xor_sync(kFullMask, edges, o);
    if (lane_id() == 0) {
        if (local_min != INT32_MAX) atomicMin(&st->local_min, local_min);
        if (edges) atomicAdd(&st->front_edges[cur], edges);
        if (zero_deg) atomicAdd(&st->n_peeled, (unsigned long long)zero_deg);
    }
}

// The leader (warp 0 of the rank's CTA 0) meets the other ranks: publish, wait, decide.
__device__ __forceinline__ void leader_exchange(const PRank &R, PRankState *st, uint32_t t, uint32_t cur, int32_t k, bool scanned,
                                                uint32_t front_now) {
    const uint32_t lane = lane_id();
    const int world = R.world;
    const unsigned long long tag = (unsigned long long)t << kTagShift;
    unsigned long long work = 0, sent_before = 0;
    int32_t lmin = INT32_MAX;
    if (lane == 0) {
        work = __ldcg(&st->work);
        lmin = __ldcg(&st->local_min);
    }
    work = __shfl_sync(kFullMask, work, 0);
    lmin = __shfl_sync(kFullMask, lmin, 0);
    // messages sent this sub-round count as work: a level is over only when nothing was discovered, sliced or sent
    unsigned long long sent_q = 0;
    if ((int)lane < world) {
        sent_q = __ldcg(&st->sent[lane]);
        sent_before = st->published[lane];
        st->published[lane] = sent_q;
    }
    unsigned long long sent_now = (int)lane < world ? sent_q - sent_before : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sent_now += __shfl_xor_sync(kFullMask, sent_now, o);
    work += sent_now;
    if (work > kValMask) work = kValMask;
    __threadfence_system();
    if ((int)lane < world) {
        PeelCtl *dst = R.ctl_peer[lane];
        st_relaxed_sys(&dst->w[R.rank][1], tag | work);
        st_relaxed_sys(&dst->w[R.rank][2], tag | (unsigned long long)(uint32_t)lmin);
        st_relaxed_sys(&dst->w[R.rank][3], tag | (unsigned long long)front_now);
        st_relaxed_sys(&dst->w[R.rank][0], tag | sent_q);
    }
    // wait for every rank's words of this sub-round
    unsigned long long w0 = 0, w1 = 0, w2 = (unsigned long long)(uint32_t)INT32_MAX, w3 = 0;
    bool failed = false;
    if ((int)lane < world) {
        const PeelCtl *mine = R.ctl_local;
        const unsigned long long t0 = pglobal_ns();
        uint32_t spins = 0;
        while (true) {
            w0 = ld_acq_sys(&mine->w[lane][0]);
            w1 = ld_acq_sys(&mine->w[lane][1]);
            w2 = ld_acq_sys(&mine->w[lane][2]);
            w3 = ld_acq_sys(&mine->w[lane][3]);
            if ((w0 >> kTagShift) == t && (w1 >> kTagShift) == t && (w2 >> kTagShift) == t && (w3 >> kTagShift) == t) break;
            if ((++spins & 1023u) == 0 && (*(volatile uint32_t *)&st->error || pglobal_ns() - t0 > kPeelWatchdogNs)) { failed = true; break; }
        }
        w0 &= kValMask; w1 &= kValMask; w2 &= kValMask; w3 &= kValMask;
    }
    if (__ballot_sync(kFullMask, failed)) {
        if (lane == 0) {
            atomicCAS(&st->error, 0u, 2u);
            st_rel_gpu(&st->plan_b[t % kPlanRing], ((unsigned long long)t << 8) | kFlagDone | kModeSolo);
        }
        return;
    }
    unsigned long long inbox = 0;
    if ((int)lane < world) {
        const unsigned long long before = __ldcg(&st->recv_hi[lane]);
        st->applied[lane] = before;
        st->recv_hi[lane] = w0;
        inbox = w0 - before;
    }
    unsigned long long g_work = (int)lane < world ? w1 : 0ull, g_front = (int)lane < world ? w3 : 0ull, g_inbox = inbox;
    uint32_t g_min = (int)lane < world ? (uint32_t)w2 : (uint32_t)INT32_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        g_work += __shfl_xor_sync(kFullMask, g_work, o);
        g_front += __shfl_xor_sync(kFullMask, g_front, o);
        g_inbox += __shfl_xor_sync(kFullMask, g_inbox, o);
        g_min = min(g_min, __shfl_xor_sync(kFullMask, g_min, o));
    }
    __threadfence();   // every lane's stores to applied / recv_hi are ordered before lane 0's release of the plan
    __syncwarp();
    if (lane == 0) {
        uint32_t flags = g_inbox <= kSoloInbox ? kModeSolo : kModeFull;
        st->msg_recv_total += g_inbox;
        st->subrounds = t;
        if (scanned && g_front) { st->levels += 1; st->max_core = k; }
        if (g_work == 0) {
            // nothing was discovered, sliced or sent anywhere: the level is over
            flags |= kFlagLevelOver;
            if (scanned && g_front == 0) {
                if (g_min == (uint32_t)INT32_MAX) flags |= kFlagDone;    // nothing alive anywhere
                st->k_next[t % kPlanRing] = (int32_t)g_min;               // the level was empty: skip to the smallest degree
            } else {
                st->k_next[t % kPlanRing] = k + 1;
            }
            st->local_min = INT32_MAX;
        }
        // the list that was walked this sub-round becomes the one the next sub-round appends to
        st->front_cnt[cur] = 0;
        st->slice_cnt[cur] = 0;
        st->front_edges[cur] = 0;
        st->work = 0;
        __threadfence();
        st_rel_gpu(&st->plan_b[t % kPlanRing], ((unsigned long long)t << 8) | flags);
    }
}

__global__ void __launch_bounds__(kPThreads) ppeel_kernel(const PRank *ranks, uint32_t ctas_per_rank) {
    __shared__ uint32_t s_scan[kPWarps + 1];
    __shared__ uint32_t s_base[2];
    __shared__ uint32_t s_bcast;
    const PRank R = ranks[blockIdx.x / ctas_per_rank];
    PRankState *st = R.st;
    const uint32_t cta = blockIdx.x % ctas_per_rank;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const bool leader_cta = cta == 0;
    unsigned long long bar_gen = 0;
    uint32_t t = 1;          // sub-round number (tags of zero-initialised control words are 0)
    int32_t k = 0;
    bool scan = true;
    uint32_t cur = 0;
    const uint32_t *alive_src = nullptr;   // nullptr: every local unitig
    uint32_t alive_i = 0;
    uint32_t n_alive = R.n_local;

    while (true) {
        if (scan) {
            scan_stage(R, st, cta, alive_src, n_alive, R.alive[alive_i], &st->alive_out[alive_i], cur, k, s_scan, s_base);
            rank_barrier(st, R.ctas, bar_gen);
            n_alive = __ldcg(&st->alive_out[alive_i]);
            alive_src = R.alive[alive_i];
            alive_i ^= 1u;
            // the other slot was last read two scans ago (every CTA has passed a barrier since): re-arm it for the next scan
            if (leader_cta && tid == 0) st->alive_out[alive_i] = 0;
        } else if ((t % kPlanSync) == 0) {
            rank_barrier(st, R.ctas, bar_gen);   // bounds the leader's lead over the other CTAs (plan ring)
        }
        // ---- plan A: who walks the frontier
        uint32_t front_now = 0;
        if (leader_cta && tid == 0) {
            const uint32_t nf = __ldcg(&st->front_cnt[cur]), ns = __ldcg(&st->slice_cnt[cur]);
            const unsigned long long ne = __ldcg(&st->front_edges[cur]);
            const uint32_t mode = (ns == 0 && nf <= kSoloFront && ne <= kSoloEdges) ? kModeSolo : kModeFull;
            if (mode == kModeSolo) st->solo_subrounds += 1;
            s_base[0] = nf;
            __threadfence();
            st_rel_gpu(&st->plan_a[t % kPlanRing], ((unsigned long long)t << 8) | mode);
        }
        const uint32_t mode_a = wait_plan(&st->plan_a[t % kPlanRing], t, st, &s_bcast);
        if (mode_a & kFlagDone) break;   // watchdog
        if (leader_cta) front_now = s_base[0];
        // ---- process
        if (mode_a & kModeFull) {
            process_stage(R, st, cta * kPWarps + warp, R.ctas * kPWarps, cur, k);
            rank_barrier(st, R.ctas, bar_gen);
        } else if (leader_cta) {
            process_stage(R, st, warp, kPWarps, cur, k);
            __syncthreads();
            if (tid == 0) __threadfence_system();
            __syncthreads();
        }
        // ---- exchange
        if (leader_cta && warp == 0) leader_exchange(R, st, t, cur, k, scan, front_now);
        const uint32_t flags = wait_plan(&st->plan_b[t % kPlanRing], t, st, &s_bcast);
        // ---- apply
        if (flags & kModeFull) {
            apply_stage(R, st, cta * kPWarps + warp, R.ctas * kPWarps, cur, k);
            rank_barrier(st, R.ctas, bar_gen);
        } else if (leader_cta) {
            apply_stage(R, st, warp, kPWarps, cur, k);
            __syncthreads();
            if (tid == 0) __threadfence();
            __syncthreads();
        }
        if ((flags & kFlagDone) || *(volatile uint32_t *)&st->error) break;
        cur ^= 1u;
        ++t;
        if (t >= (1u << 24) - 2u) { if (tid == 0) atomicCAS(&st->error, 0u, 1u); break; }
        if (flags & kFlagLevelOver) {
            k = __ldcg(&st->k_next[(t - 1u) % kPlanRing]);
            scan = true;
        } else {
            scan = false;
        }
    }
}

// owners of the CSR entries of the local rows: cross[p] = entries whose target rank p owns
__global__ void __launch_bounds__(256) cross_count_kernel(const uint32_t *__restrict__ col, uint64_t n, uint32_t step, int world,
                                                          unsigned long long *__restrict__ cross) {
    __shared__ uint32_t s_cnt[kMaxRanks];
    if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t o = i < n ? min(col[i] / step, (uint32_t)world - 1u) : 0xffffffffu;
        const uint32_t same = __match_any_sync(kFullMask, o);
        if (o != 0xffffffffu && lane == (uint32_t)__ffs(same) - 1u) atomicAdd(&s_cnt[o], (uint32_t)__popc(same));
    }
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(&cross[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

}  // namespace

int dist_peel(kombgpu_dist_graph *g) {
    kombgpu_comm *c = g->comm;
    kombgpu_ctx *ctx = g->ctx;
    const int world = c->world;
    const uint32_t n_local = g->n_local;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    KG_CUDA(ctx, cudaEventCreate(&ev0));
    KG_CUDA(ctx, cudaEventCreate(&ev1));
    KG_CUDA(ctx, cudaEventRecord(ev0, ctx->stream));
    if (!g->core) {
        g->core = static_cast<int32_t *>(ws_alloc(ctx, (n_local ? n_local : 1) * sizeof(int32_t)));
        if (!g->core) return ctx_fail(ctx, KOMBGPU_ENOMEM, "coreness array");
    }
    KG_CUDA(ctx, cudaMemsetAsync(g->core, 0, (size_t)(n_local ? n_local : 1) * sizeof(int32_t), ctx->stream));

    // mailbox sizes: cross[q][p] = CSR entries of rank q whose target rank p owns
    DevBuf<unsigned long long> d_cross(ctx, kMaxRanks);
    if (!d_cross) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemsetAsync(d_cross.p, 0, kMaxRanks * sizeof(unsigned long long), ctx->stream));
    if (g->n_directed)
        KG_LAUNCH(ctx, cross_count_kernel, min(ceil_div_u64(g->n_directed, 256), (uint32_t)ctx->sm_count * 8u), 256, 0, g->col,
                  g->n_directed, g->step, world, d_cross.p);
    unsigned long long h_cross[kMaxRanks] = {}, matrix[kMaxRanks * kMaxRanks];
    KG_TRY(read_back(ctx, d_cross.p, h_cross, kMaxRanks));
    KG_TRY(comm_exchange(c, h_cross, world, matrix));   // matrix[q * world + p]
    unsigned long long in_off[kMaxRanks + 1] = {}, out_off[kMaxRanks] = {}, max_in = 0;
    for (int p = 0; p < world; ++p) {
        unsigned long long tot = 0;
        for (int q = 0; q < world; ++q) {
            if (p == c->rank) in_off[q] = tot;
            if (q == c->rank) out_off[p] = tot;    // where this rank's messages start in rank p's mailbox block
            tot += (q == p) ? 0ull : matrix[q * world + p];
        }
        if (p == c->rank) in_off[world] = tot;
        if (tot > max_in) max_in = tot;
    }
    const SymMark mark = sym_mark(c);
    uint32_t *mbox = nullptr;
    PeerPtrs<uint32_t> mbox_peers{};
    PeelCtl *ctl = nullptr;
    PeerPtrs<PeelCtl> ctl_peers{};
    KG_TRY(sym_alloc(c, (size_t)max_in, &mbox, &mbox_peers));
    KG_TRY(sym_alloc(c, 1, &ctl, &ctl_peers));
    KG_CUDA(ctx, cudaMemsetAsync(ctl, 0, sizeof(PeelCtl), ctx->stream));

    // per-rank lists and state
    DevBuf<int32_t> work;
    DevBuf<uint32_t> alive_a, alive_b, front_a, front_b;
    DevBuf<uint64_t> slices_a, slices_b;
    DevBuf<PRankState> state(ctx, 1);
    DevBuf<PRank> desc(ctx, kMaxRanks);
    const uint64_t slice_cap64 = g->n_directed / kSliceLen + (uint64_t)n_local + 64;
    if (slice_cap64 >= 0xffffffffull || g->n_directed >= (1ull << 51)) return ctx_fail(ctx, KOMBGPU_EINVAL, "partition too large for the slice encoding");
    KG_ALLOC(ctx, work, n_local);
    KG_ALLOC(ctx, alive_a, n_local);
    KG_ALLOC(ctx, alive_b, n_local);
    KG_ALLOC(ctx, front_a, n_local);
    KG_ALLOC(ctx, front_b, n_local);
    KG_ALLOC(ctx, slices_a, slice_cap64);
    KG_ALLOC(ctx, slices_b, slice_cap64);
    if (!state || !desc) return ctx_fail(ctx, KOMBGPU_ENOMEM, "workspace");
    KG_CUDA(ctx, cudaMemcpyAsync(work.p, g->deg, (size_t)n_local * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    PRankState init{};
    init.local_min = INT32_MAX;
    KG_CUDA(ctx, cudaMemcpyAsync(state.p, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));

    int per_sm = 0;
    KG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppeel_kernel, kPThreads, 0));
    if (per_sm < 1) return ctx_fail(ctx, KOMBGPU_ECUDA, "partitioned peel kernel does not fit on an SM");
    const int resident = per_sm * ctx->sm_count;
    const uint32_t ctas_per_rank = c->same_device ? (uint32_t)(resident / world) : (uint32_t)resident;
    if (ctas_per_rank < 1) return ctx_fail(ctx, KOMBGPU_EINVAL, "too many ranks on one device");

    PRank R{};
    R.n_local = n_local; R.v_lo = g->v_lo; R.step = g->step; R.world = world; R.rank = c->rank; R.ctas = ctas_per_rank;
    R.row_ptr = g->row_ptr; R.col = g->col; R.deg = work.p; R.core = g->core;
    R.alive[0] = alive_a.p; R.alive[1] = alive_b.p; R.front[0] = front_a.p; R.front[1] = front_b.p;
    R.slices[0] = slices_a.p; R.slices[1] = slices_b.p; R.slice_cap = (uint32_t)slice_cap64;
    R.ctl_local = ctl;
    for (int q = 0; q < world; ++q) {
        R.ctl_peer[q] = ctl_peers.p[q];
        R.mbox_in[q] = mbox + in_off[q];
        R.mbox_out[q] = mbox_peers.p[q] + out_off[q];
        R.mbox_out_cap[q] = q == c->rank ? 0ull : matrix[c->rank * world + q];
    }
    R.st = state.p;

    // every rank's control words are cleared before anyone publishes into them
    unsigned long long token = 1, tokens[kMaxRanks];
    KG_TRY(comm_exchange(c, &token, 1, tokens));

    cudaError_t le = cudaSuccess;
    if (!c->same_device) {
        KG_CUDA(ctx, cudaMemcpyAsync(desc.p, &R, sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
        const PRank *dp = desc.p;
        uint32_t cpr = ctas_per_rank;
        void *args[] = {(void *)&dp, (void *)&cpr};
        le = cudaLaunchCooperativeKernel((void *)ppeel_kernel, dim3(ctas_per_rank), dim3(kPThreads), args, 0, ctx->stream);
        ctx->launches++;
        if (le == cudaSuccess) le = cudaStreamSynchronize(ctx->stream);
    } else {
        // emulation: one cooperative grid holds every rank (a group of CTAs each); rank 0's thread launches it
        LocalGroup *grp = c->group;
        grp->slot[c->rank] = &R;
        KG_TRY(comm_group_barrier(c));
        if (c->rank == 0) {
            std::vector<PRank> all(world);
            for (int q = 0; q < world; ++q) all[q] = *static_cast<PRank *>(grp->slot[q]);
            le = cudaMemcpyAsync(desc.p, all.data(), sizeof(PRank) * world, cudaMemcpyHostToDevice, ctx->stream);
            const PRank *dp = desc.p;
            uint32_t cpr = ctas_per_rank;
            void *args[] = {(void *)&dp, (void *)&cpr};
            if (le == cudaSuccess)
                le = cudaLaunchCooperativeKernel((void *)ppeel_kernel, dim3(ctas_per_rank * world), dim3(kPThreads), args, 0, ctx->stream);
            ctx->launches++;
            if (le == cudaSuccess) le = cudaStreamSynchronize(ctx->stream);
            grp->slot_rc = le == cudaSuccess ? 0 : 1;
        }
        KG_TRY(comm_group_barrier(c));
        if (grp->slot_rc) le = cudaErrorLaunchFailure;
        KG_TRY(comm_group_barrier(c));   // slot_rc was read by everyone before the next use
    }
    if (le != cudaSuccess) return ctx_fail(ctx, KOMBGPU_ECUDA, "partitioned peel kernel: %s", cudaGetErrorString(le));

    PRankState fin{};
    KG_TRY(read_back(ctx, state.p, &fin, 1));
    unsigned long long sent_total = 0;
    for (int q = 0; q < world; ++q) sent_total += fin.sent[q];
    // global figures; also: nobody releases its mailbox while a peer may still be writing
    unsigned long long mine[4] = {fin.error, fin.n_peeled, (unsigned long long)(uint32_t)fin.max_core, sent_total}, all[kMaxRanks * 4];
    KG_TRY(comm_exchange(c, mine, 4, all));
    sym_release(c, mark);
    uint64_t peeled = 0;
    for (int q = 0; q < world; ++q) {
        if (all[q * 4]) return ctx_fail(ctx, KOMBGPU_EINTERNAL, "partitioned peel: rank %d reports error %llu (1/2 watchdog, 3 mailbox, 4 bad message, 5 list)", q, all[q * 4]);
        peeled += all[q * 4 + 1];
    }
    if (peeled != g->n_global) return ctx_fail(ctx, KOMBGPU_EINTERNAL, "partitioned peel ended with %llu of %u unitigs peeled", (unsigned long long)peeled, g->n_global);
    g->st.max_coreness = fin.max_core;
    g->st.peel_levels = fin.levels;
    g->st.peel_subrounds = fin.subrounds;
    g->st.peel_solo_subrounds = fin.solo_subrounds;
    g->st.n_messages_sent = sent_total;
    g->st.n_messages_recv = fin.msg_recv_total;
    g->has_core = true;
    KG_CUDA(ctx, cudaEventRecord(ev1, ctx->stream));
    KG_CUDA(ctx, cudaEventSynchronize(ev1));
    cudaEventElapsedTime(&g->st.ms_peel, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return KOMBGPU_OK;
}

}  // namespace kg
