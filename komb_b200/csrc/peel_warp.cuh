// peel_warp.cuh — PROCESS phase of one peel level with warp-autonomous workers.
//
// The CTA-wide version in peel_device.cuh moves a whole CTA through every cascade step (block scan, five CTA
// barriers per batch: ~3.6 us per dependent step measured on B200).  Here the 16 warps of a CTA never meet at a
// barrier inside a level.  They share one shared-memory TICKET RING:
//
//   r_tail  tickets pushed (slot written right after the fetch-add that reserves it)
//   r_head  tickets claimed or reserved by consumers
//   r_done  tickets fully processed, credited after the consumer's own pushes
//
// A warp claims tickets with one shared-memory fetch-add; if the ring is empty the claim is a reservation and the
// warp polls the slot's own word, so the warp that discovers the next vertex hands it over with one store.  A
// cascade step is  atomicSub -> ballot -> ring store -> (any parked warp) -> row_ptr -> col -> atomicSub  with no
// barrier and no block-wide scan.  Edges of a batch (up to 32 rows, one per lane) are spread over the lanes with a
// shuffle-based search of the row prefix, four independent chains per lane.
//
// The global pool protocol of peel_device.cuh is unchanged, but it is spoken by ONE warp per CTA at a time, the
// AGENT: the parked warp that holds the LAST reservation (its range ends at r_head), so that discoveries of the
// CTA's own warps are handed to plain parked warps first (they poll shared memory only; the agent may be waiting
// for a global load) and work the agent brings in reaches every other parked warp before itself.  When the CTA is locally
// quiescent (r_done == r_tail) the agent settles the CTA's credit with q_done, detects the end of the level, or
// claims / reserves pool slots.  Pool entries known to be below q_tail are handed to the workers as RANGE
// descriptors (one ring ticket = up to 32 pool entries, loaded by the worker itself), reserved pool slots are
// polled by the agent and copied into the ring when they are written.  Discoveries beyond what the CTA's warps can
// start on at once (kRingKeep queued tickets) go to the pool, where idle CTAs' agents are parked.
#pragma once

#include "peel_device.cuh"

namespace kg {
namespace peel {

constexpr int kRing = 4096;               // ring slots (power of two)
constexpr uint32_t kRingMask = kRing - 1;
constexpr int32_t kRingKeep = 16;         // default of WarpTune::keep
constexpr uint64_t kRangeBit = 1ull << 62;  // ring-only entry: kRangeBit | first_pool_slot << 8 | count
constexpr uint32_t kRangeLen = 32;
constexpr uint32_t kWarpSplit = 128;       // default of WarpTune::wsplit
constexpr uint32_t kWarpSplitMin = 128;    // lower bound of the knob (sizes the pool)
constexpr uint32_t kStaticMax = 16 * kRangeLen;  // frontier entries dealt to a CTA without a claim
constexpr uint32_t kSpinCheck = 1u << 16;  // polls between two looks at the watchdog clock

// run-time knobs (PeelState::tune; defaults below, overridable through KOMBGPU_PEEL_* for measurements)
struct WarpTune {
    int32_t keep;      // queued ring tickets above which discoveries are shared through the pool
    uint32_t wsplit;   // edges of a row one warp keeps
    uint32_t park_ns;  // sleep between two polls of a parked warp (0 = spin)
    uint32_t thin;     // a batch up to this many edges may hand its discoveries on in registers
    uint32_t hub_slice;  // edges per pool slice of a row longer than kSplit
};

struct WarpShared {
    uint64_t ring[kRing];
    uint32_t r_head, r_tail, r_done;
    uint32_t over;        // the level has ended (or an invariant broke): every warp leaves
    uint32_t lock;        // agent mutual exclusion
    uint32_t credit;      // pool tasks this CTA took and has not credited to q_done yet
    uint32_t gb, ge;      // reserved pool slots [gb, ge) this CTA polls
    uint32_t polls;
    uint32_t removed;
    uint32_t shared_cnt;
    uint32_t batches;
    uint32_t carried, ring_pushed;
    uint32_t spare;
    long long t[6];       // CTA 0, trace only: entry, init done, first batch, last batch end, level over, exit
};

__device__ __forceinline__ uint32_t lds_u32(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ uint64_t lds_u64(const uint64_t *p) { return *reinterpret_cast<const volatile uint64_t *>(p); }
__device__ __forceinline__ void sts_u32(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
__device__ __forceinline__ void sts_u64(uint64_t *p, uint64_t v) { *reinterpret_cast<volatile uint64_t *>(p) = v; }

// store one ring ticket; the slot's previous lap has been consumed long ago (the ring never holds more than
// kRingKeep + 16 warps x 128 tickets), the wait only makes that explicit
__device__ __forceinline__ void ring_put(WarpShared &sh, uint32_t ticket, uint64_t e) {
    uint64_t *slot = &sh.ring[ticket & kRingMask];
    uint32_t tries = 0;
    while (lds_u64(slot) != kEmpty && ++tries < (1u << 24)) {}
    sts_u64(slot, e);
}

// One batch of a warp: lanes [0, m) hold tasks `ent` (a vertex, a slice of a long row, or kEmpty).  Returns how
// many vertices were peeled.  If the batch fits one traversal iteration and discovers at most kCarry vertices,
// they are NOT queued: they come back in `carry` (lanes [0, n_carry)) and the caller runs them as its next batch.
// A dependent step of a thin cascade then costs row_ptr -> col -> atomic plus a few shuffles, instead of a ring
// hand-off, a parked warp's wake-up and the 32-row bookkeeping (2.2 us per step measured on a path graph, of
// which the three memory round trips are 0.45 us).
constexpr uint32_t kCarry = 2;
constexpr uint32_t kThinEdges = 512;   // a batch up to this size may hand its discoveries on in registers
template <bool kDist, int kU>
__device__ __forceinline__ uint32_t warp_batch(const uint64_t ent, const uint32_t m, const int32_t k,
                                               const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, int32_t *deg,
                                               int32_t *core_out, uint64_t *Q, const uint32_t cap, PeelState *st, WarpShared &sh,
                                               const PartView &part, const WarpTune &tn, uint32_t &n_shared, uint64_t &carry,
                                               uint32_t &n_carry) {
    const uint32_t lane = lane_id();
    n_carry = 0;
    carry = kEmpty;
    uint32_t my_len = 0;
    uint64_t my_row = 0;
    bool is_vertex = false;
    if (ent != kEmpty) {
        if (ent & kSliceBit) {
            my_row = (ent & ~kSliceBit) >> kSliceLenBits;
            my_len = (uint32_t)(ent & ((1u << kSliceLenBits) - 1));
        } else {
            const uint32_t v = (uint32_t)ent;
            is_vertex = true;
            core_out[v] = k;   // peeled at level k
            my_row = row_ptr[v];
            my_len = (uint32_t)(row_ptr[v + 1] - my_row);
            if (my_len > kSplit) {
                // hub row: hand it to the whole grid as slices
                const uint32_t n_sl = (my_len + tn.hub_slice - 1) / tn.hub_slice;
                const uint32_t s0 = atomicAdd(&st->q_tail, n_sl);
                for (uint32_t i = 0; i < n_sl; ++i) {
                    const uint64_t e = kSliceBit | ((my_row + (uint64_t)i * tn.hub_slice) << kSliceLenBits) |
                                       min(tn.hub_slice, my_len - i * tn.hub_slice);
                    if (s0 + i < cap) st_volatile_u64(&Q[s0 + i], e);
                    else atomicExch(&st->error, 3u);
                }
                atomicAdd(&st->sliced, (unsigned long long)n_sl);
                my_len = 0;
            }
        }
    }
    // rows longer than kWarpSplit: keep the first piece, queue the rest as slices any warp (or, when the ring is
    // busy, any CTA) can take -- a long row must not serialise on one warp
    {
        const uint32_t extra = my_len > tn.wsplit ? (my_len - 1) / tn.wsplit : 0u;
        if (__ballot_sync(kFullMask, extra != 0)) {
            const uint32_t xin = warp_incl_scan_add(extra);
            const uint32_t xtot = __shfl_sync(kFullMask, xin, 31);
            uint32_t pos = 0, to_pool = 0;
            if (lane == 0) {
                const int32_t queued = (int32_t)(lds_u32(&sh.r_tail) - lds_u32(&sh.r_head));
                to_pool = (queued >= tn.keep || xtot > 128u) ? 1u : 0u;
                pos = to_pool ? atomicAdd(&st->q_tail, xtot) : atomicAdd(&sh.r_tail, xtot);
            }
            pos = __shfl_sync(kFullMask, pos, 0);
            to_pool = __shfl_sync(kFullMask, to_pool, 0);
            // the xtot tickets are written by all lanes together (ticket t belongs to the first row j with
            // xin[j] > t): one lane writing the 30 pieces of a 4096-edge row costs more than walking a piece
            const uint32_t rlo = (uint32_t)my_row, rhi = (uint32_t)(my_row >> 32);
            for (uint32_t tb = 0; tb < xtot; tb += 32) {
                const uint32_t t = tb + lane;
                uint32_t j = 0;
#pragma unroll
                for (uint32_t sft = 16; sft > 0; sft >>= 1) {
                    const uint32_t x = __shfl_sync(kFullMask, xin, (j + sft - 1) & 31u);
                    if (x <= t) j += sft;
                }
                j &= 31u;
                const uint32_t j_xin = __shfl_sync(kFullMask, xin, j), j_extra = __shfl_sync(kFullMask, extra, j);
                const uint32_t j_len = __shfl_sync(kFullMask, my_len, j);
                const uint64_t j_row = ((uint64_t)__shfl_sync(kFullMask, rhi, j) << 32) | __shfl_sync(kFullMask, rlo, j);
                if (t < xtot) {
                    const uint32_t i = t - (j_xin - j_extra) + 1;   // piece 1..extra of row j (piece 0 stays here)
                    const uint64_t e = kSliceBit | ((j_row + (uint64_t)i * tn.wsplit) << kSliceLenBits) |
                                       min(tn.wsplit, j_len - i * tn.wsplit);
                    if (to_pool) {
                        if (pos + t < cap) st_volatile_u64(&Q[pos + t], e);
                        else atomicExch(&st->error, 3u);
                    } else {
                        ring_put(sh, pos + t, e);
                    }
                }
            }
            if (extra) my_len = tn.wsplit;
        }
    }
    uint32_t excl, total;
    if (m <= 2) {   // the batch of a thin cascade: no scan
        const uint32_t len0 = __shfl_sync(kFullMask, my_len, 0), len1 = __shfl_sync(kFullMask, my_len, 1);
        total = len0 + (m == 2 ? len1 : 0u);
        excl = lane == 0 ? 0u : ((lane == 1 && m == 2) ? len0 : total);
    } else {
        const uint32_t incl = warp_incl_scan_add(my_len);
        excl = incl - my_len;
        total = __shfl_sync(kFullMask, incl, 31);
    }
    const uint32_t row_lo = (uint32_t)my_row, row_hi = (uint32_t)(my_row >> 32);
    // latency-bound batches decrement without looking first: one iteration's worth of edges, or the (<= 2 row)
    // batch of a cascade being followed, whose rows may be a few hundred entries long
    const bool thin = total <= tn.thin;
    const bool direct = total <= 32u * kU || (m <= kCarry && thin);
    // binary search over the rows of the batch: log2 steps for m rows (none for one row)
    const uint32_t top = m <= 1 ? 0u : (1u << (31 - __clz((int)(m - 1))));

    for (uint32_t base = 0; base < total; base += 32u * kU) {
        uint32_t u[kU];
        int32_t d[kU];
        bool push[kU];
#pragma unroll
        for (int t = 0; t < kU; ++t) {
            const uint32_t e = base + t * 32u + lane;
            u[t] = kFullMask;
            if (base + t * 32u >= total) continue;   // warp-uniform: no edge in this group of 32 slots
            // owner row: the last lane j with excl[j] <= e (rows of length 0 are skipped by "last")
            uint32_t j = 0;
            for (uint32_t s = top; s > 0; s >>= 1) {
                const uint32_t x = __shfl_sync(kFullMask, excl, (j + s) & 31u);
                if (x <= e) j += s;
            }
            const uint32_t ex_j = __shfl_sync(kFullMask, excl, j);
            const uint32_t lo = __shfl_sync(kFullMask, row_lo, j), hi = __shfl_sync(kFullMask, row_hi, j);
            if (e < total) u[t] = col[(((uint64_t)hi << 32) | lo) + (e - ex_j)];
        }
        if (kDist) {
            // split off the neighbours other ranks own: ship their ids, keep local ones as local ids
#pragma unroll
            for (int t = 0; t < kU; ++t) {
                const bool valid = u[t] != kFullMask;
                const uint32_t loc = u[t] - part.v_lo;
                const bool remote = valid && loc >= part.n_local;
                const uint32_t rm = __ballot_sync(kFullMask, remote);
                if (rm) {
                    uint32_t pos = 0;
                    if (lane == 0) pos = atomicAdd(part.outbox_cnt, (uint32_t)__popc(rm));
                    pos = __shfl_sync(kFullMask, pos, 0) + __popc(rm & lanemask_lt());
                    if (remote) part.outbox[pos] = u[t];
                }
                u[t] = (valid && !remote) ? loc : kFullMask;
            }
        }
        // deg[] is NOT clamped in this mode (the coreness goes to core_out when a vertex is taken): a decrement is
        // one fire-and-look atomic, never undone, never preceded by a load.  deg[u] passes k + 1 -> k exactly once,
        // and the thread that sees it owns u.  (Decrements of vertices already peeled are wasted, not wrong: a
        // vertex receives at most one per neighbour, so deg stays >= 0.)
        if (direct) {
            // latency-bound batch (one iteration): no look before the decrement
#pragma unroll
            for (int t = 0; t < kU; ++t) d[t] = (u[t] != kFullMask) ? atomicSub(&deg[u[t]], 1) : INT32_MAX;
        } else {
            // throughput-bound batch: a load is cheaper than an atomic on a vertex that is already gone
#pragma unroll
            for (int t = 0; t < kU; ++t) d[t] = (u[t] != kFullMask) ? __ldcg(&deg[u[t]]) : INT32_MIN;
#pragma unroll
            for (int t = 0; t < kU; ++t) d[t] = (d[t] > k) ? atomicSub(&deg[u[t]], 1) : INT32_MAX;
        }
#pragma unroll
        for (int t = 0; t < kU; ++t) push[t] = d[t] == k + 1;
        uint32_t c = 0;
#pragma unroll
        for (int t = 0; t < kU; ++t) c += push[t] ? 1u : 0u;
        const uint32_t any = __ballot_sync(kFullMask, c != 0);
        if (any == 0) continue;
        if (thin && n_carry + (uint32_t)__popc(any) <= kCarry && __ballot_sync(kFullMask, c > 1) == 0) {
            // a thin cascade: keep the discoveries in registers, they are this warp's next batch
            uint32_t val = 0;
#pragma unroll
            for (int t = 0; t < kU; ++t) if (push[t]) val = u[t];
            uint32_t left = any;
#pragma unroll
            for (uint32_t i = 0; i < kCarry; ++i) {   // the i-th discovery of this iteration goes to lane n_carry + i
                const uint32_t vi = __shfl_sync(kFullMask, val, left ? __ffs(left) - 1 : 0);
                if (lane == n_carry + i && left) carry = (uint64_t)vi;
                left &= left - 1;
            }
            n_carry += (uint32_t)__popc(any);
            continue;
        }
        const uint32_t inc = warp_incl_scan_add(c);
        const uint32_t tot = __shfl_sync(kFullMask, inc, 31);
        uint32_t pos = 0, to_pool = 0;
        if (lane == 0) {
            const int32_t queued = (int32_t)(lds_u32(&sh.r_tail) - lds_u32(&sh.r_head));
            to_pool = queued >= tn.keep ? 1u : 0u;
            pos = to_pool ? atomicAdd(&st->q_tail, tot) : atomicAdd(&sh.r_tail, tot);
        }
        pos = __shfl_sync(kFullMask, pos, 0) + (inc - c);
        to_pool = __shfl_sync(kFullMask, to_pool, 0);
        if (to_pool) n_shared += c; else n_shared += c << 16;   // low half: to the pool, high half: to the ring (per batch < 65536)
#pragma unroll
        for (int t = 0; t < kU; ++t) {
            if (!push[t]) continue;
            if (to_pool) {
                if (pos < cap) st_volatile_u64(&Q[pos], (uint64_t)u[t]);
                else atomicExch(&st->error, 3u);
            } else {
                ring_put(sh, pos, (uint64_t)u[t]);
            }
            ++pos;
        }
    }
    return (uint32_t)__popc(__ballot_sync(kFullMask, is_vertex));
}

// One step of the agent (whole warp, sh.lock held by lane 0).  Either delivers work into the ring, ends the level
// (sh.over), or does nothing because other warps of the CTA are still working.
__device__ __forceinline__ void agent_step(const uint64_t token, uint64_t *Q, const uint32_t cap, PeelState *st, WarpShared &sh,
                                           const uint32_t dealt, unsigned long long &idle_since) {
    const uint32_t lane = lane_id();
    const uint32_t gb = lds_u32(&sh.gb), ge = lds_u32(&sh.ge);
    // 1. reserved pool slots: copy the leading run of written ones into the ring
    if (gb < ge) {
        uint64_t e = kEmpty;
        bool tok = false;
        if (lane < ge - gb) {
            e = ld_volatile_u64(&Q[gb + lane]);
            if (e == token) { tok = true; st_volatile_u64(&Q[gb + lane], kEmpty); }  // leave the slot clean
        }
        const uint32_t rm = __ballot_sync(kFullMask, entry_is_task(e));
        const uint32_t tk = __ballot_sync(kFullMask, tok);
        const uint32_t m = (rm == kFullMask) ? 32u : (uint32_t)__ffs(~rm) - 1u;
        if (m) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&sh.r_tail, m);
            pos = __shfl_sync(kFullMask, pos, 0);
            if (lane < m) ring_put(sh, pos + lane, e);
            if (lane == 0) { sts_u32(&sh.gb, gb + m); sts_u32(&sh.credit, lds_u32(&sh.credit) + m); sts_u32(&sh.polls, 0); }
            idle_since = 0;
            return;
        }
        if (tk) {  // level over: nothing can be pending anywhere
            if (lane == 0) sts_u32(&sh.over, 1u);
            return;
        }
    }
    // 2. anything still running in this CTA?  (r_done is read first: both counters are monotone and
    //    r_done <= r_tail, so equality of this pair means quiescence at the time r_done was read)
    uint32_t quiet = 0;
    if (lane == 0) {
        const uint32_t dn = lds_u32(&sh.r_done);
        const uint32_t tl = lds_u32(&sh.r_tail);
        quiet = dn == tl ? 1u : 0u;
    }
    quiet = __shfl_sync(kFullMask, quiet, 0);
    // a busy CTA with an idle warp listens on reserved pool slots too (work goes where idle warps are); it
    // cannot end the level and has nothing to settle
    if (!quiet && gb < ge) return;
    // 3. settle the credit of a drained CTA; then end the level, keep waiting, or claim / reserve pool slots
    uint32_t over = 0, fin_tail = 0, fin_head = 0, begin = 0, sure = 0, end = 0;
    if (lane == 0) {
        const uint32_t credit = quiet ? lds_u32(&sh.credit) : 0u;
        uint4 a = make_uint4(0, 0, 0, 0);
        bool have_a = false;
        if (credit) {
            __threadfence();
            const uint32_t old = atomicAdd(&st->q_done, credit);
            a = ld_volatile_u4(st);  // q_head, q_tail, q_done, error
            a.x += dealt;            // q_head counts claims only: the statically dealt entries come on top
            have_a = true;
            sts_u32(&sh.credit, 0);
            if (old + credit == a.y) { fin_tail = a.y; fin_head = min(a.x, cap); over = 1; }  // quiescent, final
        }
        if (!over) {
            if (gb < ge) {
                // drained and parked on reserved slots: only rarely look at the shared counters
                const uint32_t polls = lds_u32(&sh.polls) + 1;
                sts_u32(&sh.polls, polls);
                if (have_a || polls == 1 || (polls & 7u) == 0) {
                    if (!have_a) { a = ld_volatile_u4(st); a.x += dealt; }
                    if (a.z == a.y || a.w) over = 1;  // q_done == q_tail: final, nobody can append any more
                    else if (idle_since == 0) idle_since = global_ns();
                    else if (global_ns() - idle_since > kWatchdogNs) { atomicExch(&st->error, 2u); over = 1; }
                }
                if (!over) __nanosleep(100);
            } else {
                if (!have_a) { a = ld_volatile_u4(st); a.x += dealt; }
                if (a.w || (quiet && a.z == a.y)) {
                    over = 1;
                } else {
                    uint32_t take = 1;
                    if (a.x < a.y) take = min(max((a.y - a.x + gridDim.x - 1) / gridDim.x, 1u), kClaimMax);
                    begin = atomicAdd(&st->q_head, take) + dealt;
                    end = min(begin + take, cap);
                    if (begin >= cap) { atomicExch(&st->error, 4u); over = 1; end = begin; }
                    sure = min(max(a.y, begin), end);  // slots below the tail seen before the claim: written, or about to be
                    sts_u32(&sh.polls, 0);
                }
            }
        }
    }
    over = __shfl_sync(kFullMask, over, 0);
    fin_tail = __shfl_sync(kFullMask, fin_tail, 0);
    fin_head = __shfl_sync(kFullMask, fin_head, 0);
    begin = __shfl_sync(kFullMask, begin, 0);
    sure = __shfl_sync(kFullMask, sure, 0);
    end = __shfl_sync(kFullMask, end, 0);
    // this CTA ended the level: wake every agent parked on a reserved slot
    for (uint32_t i = fin_tail + lane; i < fin_head; i += 32) st_volatile_u64(&Q[i], token);
    if (over) {
        if (lane == 0) sts_u32(&sh.over, 1u);
        return;
    }
    if (end > begin) {
        // spread the claimed entries over all warps of the CTA: ranges of ceil(count / warps), at most kRangeLen
        const uint32_t rl = min(max((sure - begin + kPeelWarps - 1) / kPeelWarps, 1u), kRangeLen);
        const uint32_t n_desc = (sure - begin + rl - 1) / rl;
        if (n_desc) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&sh.r_tail, n_desc);
            pos = __shfl_sync(kFullMask, pos, 0);
            for (uint32_t i = lane; i < n_desc; i += 32) {
                const uint32_t s = begin + i * rl;
                ring_put(sh, pos + i, kRangeBit | ((uint64_t)s << 8) | (uint64_t)min(rl, sure - s));
            }
        }
        if (lane == 0) {
            sts_u32(&sh.credit, lds_u32(&sh.credit) + (sure - begin));
            sts_u32(&sh.gb, sure);
            sts_u32(&sh.ge, end);
        }
        idle_since = 0;
    }
}

// PROCESS phase of one level for one CTA (all kPeelThreads threads call it; returns the vertices the CTA peeled).
template <bool kDist, int kU = kUnroll>
__device__ __forceinline__ uint32_t process_level_warp(const int32_t k, const uint32_t round, uint64_t *Q, const uint32_t cap,
                                                       const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                                       int32_t *deg, int32_t *core_out, PeelState *st, WarpShared &sh,
                                                       const PartView &part, const uint32_t front_base = 0,
                                                       const uint32_t front_cnt = 0) {
    const uint32_t tid = threadIdx.x, lane = lane_id();
    const uint64_t token = ((uint64_t)kTokenHi << 32) | round;
    for (uint32_t i = tid; i < (uint32_t)kRing; i += kPeelThreads) sh.ring[i] = kEmpty;
    // The level's frontier is pool[front_base, front_base + front_cnt), complete before this phase (grid barrier).
    // Its first grid x S entries are dealt statically, S = min(ceil(front_cnt / grid), kStaticMax) per CTA: no
    // claim, no look at the shared counters on the way to the first edge.  q_head keeps counting claims only (it
    // starts every level at front_base), so every slot index derived from it is offset by `dealt`.
    const uint32_t S = min((front_cnt + gridDim.x - 1) / gridDim.x, kStaticMax);
    const uint32_t dealt = min(front_cnt, gridDim.x * S);
    const uint32_t my_lo = front_base + min(blockIdx.x * S, dealt), my_hi = front_base + min((blockIdx.x + 1) * S, dealt);
    const bool prof = blockIdx.x == 0 && st->trace != nullptr;
    if (tid == 0) {
        if (prof) { sh.t[0] = clock64(); sh.t[2] = 0; sh.t[3] = 0; sh.t[4] = 0; }
        sh.r_head = 0; sh.r_tail = 0; sh.r_done = 0;
        sh.over = 0; sh.lock = 0; sh.credit = my_hi - my_lo; sh.gb = 0; sh.ge = 0; sh.polls = 0;
        sh.removed = 0; sh.shared_cnt = 0; sh.batches = 0; sh.carried = 0; sh.ring_pushed = 0;
    }
    __syncthreads();
    if (my_hi > my_lo && tid < 32) {
        const uint32_t cnt = my_hi - my_lo;
        const uint32_t rl = min(max((cnt + kPeelWarps - 1) / kPeelWarps, 1u), kRangeLen);
        const uint32_t n_desc = (cnt + rl - 1) / rl;   // <= 16 (kStaticMax = 16 x kRangeLen)
        if (lane < n_desc) sh.ring[lane] = kRangeBit | ((uint64_t)(my_lo + lane * rl) << 8) | (uint64_t)min(rl, cnt - lane * rl);
        if (lane == 0) sh.r_tail = n_desc;
    }
    __syncthreads();

    if (prof && tid == 0) sh.t[1] = clock64();
    WarpTune tn;
    tn.keep = (int32_t)st->tune[0];
    tn.wsplit = st->tune[1];
    tn.park_ns = st->tune[2];
    tn.thin = st->tune[3];
    tn.hub_slice = st->tune[4];
    uint32_t rb = 0, re = 0;   // ring tickets this warp owns
    uint32_t removed = 0, n_shared = 0, batches = 0, carried = 0;
    uint32_t pool_cnt = 0, ring_cnt = 0;   // lane-local
    uint32_t spins = 0;
    unsigned long long idle_since = 0;  // lane 0

    while (true) {
        if (rb == re) {
            if (lane == 0) {
                const int32_t avail = (int32_t)(lds_u32(&sh.r_tail) - lds_u32(&sh.r_head));
                const uint32_t take = avail > 0 ? (uint32_t)min(max(avail / (2 * kPeelWarps), 1), 32) : 1u;
                rb = atomicAdd(&sh.r_head, take);
                re = rb + take;
            }
            rb = __shfl_sync(kFullMask, rb, 0);
            re = __shfl_sync(kFullMask, re, 0);
        }
        // the leading run of written tickets in the owned range
        uint64_t ent = kEmpty;
        if (lane < re - rb) ent = lds_u64(&sh.ring[(rb + lane) & kRingMask]);
        const uint32_t rm = __ballot_sync(kFullMask, ent != kEmpty);
        uint32_t m = (rm == kFullMask) ? 32u : (uint32_t)__ffs(~rm) - 1u;
        if (m == 0) {
            // parked on ticket rb
            uint32_t act = 0;  // 1: leave, 2: act as the agent
            if (lane == 0) {
                if (lds_u32(&sh.over)) act = 1;
                else if (lds_u32(&sh.r_head) == re && atomicCAS(&sh.lock, 0u, 1u) == 0u)
                    // re-check under the lock: another warp may have parked behind this one, the level may have ended
                    act = (lds_u32(&sh.r_head) == re && !lds_u32(&sh.over)) ? 2u : 3u;
            }
            act = __shfl_sync(kFullMask, act, 0);
            if (act == 1) break;
            if (act >= 2) {
                if (act == 2) agent_step(token, Q, cap, st, sh, dealt, idle_since);
                __syncwarp();
                if (lane == 0) { __threadfence_block(); atomicExch(&sh.lock, 0u); }
            } else {
                if ((++spins & (kSpinCheck - 1)) == 0 && lane == 0) {
                    if (idle_since == 0) idle_since = global_ns();
                    else if (global_ns() - idle_since > kWatchdogNs) { atomicExch(&st->error, 2u); sts_u32(&sh.over, 1u); }
                }
                if (tn.park_ns) __nanosleep(tn.park_ns);
            }
            continue;
        }
        idle_since = 0;
        if (prof && lane == 0 && sh.t[2] == 0) sh.t[2] = clock64();
        // a range descriptor is a batch of its own
        const uint32_t dm = __ballot_sync(kFullMask, ent != kEmpty && (ent & kSliceBit) == 0 && (ent & kRangeBit) != 0) & ((m == 32u) ? kFullMask : ((1u << m) - 1u));
        bool is_range = false;
        if (dm & 1u) { m = 1; is_range = true; }
        else if (dm) m = (uint32_t)__ffs(dm) - 1u;
        if (lane < m) sts_u64(&sh.ring[(rb + lane) & kRingMask], kEmpty);  // free the slots for the next lap
        if (lane >= m) ent = kEmpty;
        rb += m;
        if (is_range) {
            const uint64_t desc = __shfl_sync(kFullMask, ent, 0);
            const uint32_t s = (uint32_t)(desc >> 8), cnt = (uint32_t)(desc & 0xffu);
            ent = kEmpty;
            if (lane < cnt) {
                uint32_t tries = 0;
                while (true) {
                    ent = ld_volatile_u64(&Q[s + lane]);
                    if (entry_is_task(ent)) break;
                    if ((++tries & (kSpinCheck - 1)) == 0 && __ldcg(&st->error)) { ent = kEmpty; break; }
                    if (tries > (1u << 26)) { atomicExch(&st->error, 6u); ent = kEmpty; break; }
                }
            }
            __syncwarp();
        }
        uint64_t carry;
        uint32_t n_carry;
        removed += warp_batch<kDist, kU>(ent, is_range ? kRangeLen : m, k, row_ptr, col, deg, core_out, Q, cap, st, sh, part, tn, n_shared,
                                         carry, n_carry);
        ++batches;
        while (n_carry) {   // follow the cascade this warp started; the tickets stay open (r_done) until it ends
            const uint64_t next = carry;
            const uint32_t n_next = n_carry;
            carried += n_next;
            removed += warp_batch<kDist, kU>(next, n_next, k, row_ptr, col, deg, core_out, Q, cap, st, sh, part, tn, n_shared, carry, n_carry);
            pool_cnt += n_shared & 0xffffu; ring_cnt += n_shared >> 16; n_shared = 0;
            ++batches;
        }
        pool_cnt += n_shared & 0xffffu; ring_cnt += n_shared >> 16; n_shared = 0;
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&sh.r_done, m); if (prof) sh.t[3] = clock64(); }
    }
    if (prof && lane == 0 && sh.t[4] == 0) sh.t[4] = clock64();
    pool_cnt = warp_reduce_add(pool_cnt);
    ring_cnt = warp_reduce_add(ring_cnt);
    if (lane == 0) {
        if (removed) atomicAdd(&sh.removed, removed);
        if (pool_cnt) atomicAdd(&sh.shared_cnt, pool_cnt);
        if (ring_cnt) atomicAdd(&sh.ring_pushed, ring_cnt);
        if (carried) atomicAdd(&sh.carried, carried);
        atomicAdd(&sh.batches, batches);
    }
    __syncthreads();
    if (tid == 0) {
        if (sh.shared_cnt) atomicAdd(&st->shared, (unsigned long long)sh.shared_cnt);
        if (sh.batches) atomicAdd(&st->batches, (unsigned long long)sh.batches);
        if (sh.carried) atomicAdd(&st->carried, (unsigned long long)sh.carried);
        if (sh.ring_pushed) atomicAdd(&st->ring_pushed, (unsigned long long)sh.ring_pushed);
    }
    const uint32_t total_removed = sh.removed;
    if (prof && tid == 0) {
        sh.t[5] = clock64();
        unsigned long long *tr = st->trace + 20ull * (round - 1);   // peel.cu: kTraceWords = 20, row = round - 1
        if (round - 1 < st->trace_cap) {
            tr[10] = sh.t[1] - sh.t[0]; tr[11] = sh.t[2] ? sh.t[2] - sh.t[1] : 0; tr[12] = sh.t[3] ? sh.t[3] - sh.t[1] : 0;
            tr[13] = sh.t[4] - sh.t[1]; tr[14] = sh.t[5] - sh.t[1];
        }
    }
    __syncthreads();  // sh is re-initialised by the next call
    return total_removed;
}

}  // namespace peel
}  // namespace kg
