// stage 2: greedy non-overlapping k-mer search on both strands (search_reads.h:34-87, bloom_filter.h:124-131)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"

namespace commet {

// ------------------------------------------------------ stage 2: search ----
// BloomFilter::is_found (bloom_filter.h:124-131): b, c, d after a passed,
// short-circuit in the reference's order.
__device__ __forceinline__ bool probe_bcd(const uint32_t *__restrict__ filter, const Keys &q, unsigned int &tests)
{
    tests++;
    if (!(ld_probe_u32(filter + key_word(q.b)) & key_bit(q.b, 1))) return false;
    tests++;
    if (!(ld_probe_u32(filter + key_word(q.c)) & key_bit(q.c, 2))) return false;
    tests++;
    return (ld_probe_u32(filter + key_word(q.d)) & key_bit(q.d, 3)) != 0;
}

// One strand of search_reads (search_reads.h:46-64 forward, :66-83 reverse):
// left-to-right greedy scan; on a hit seen++ and, unless seen >= t, the next
// candidate is k positions later (hash.clear()).  The lane keeps a 96-bit
// register window of the H/L/W planes and issues kSearchBatch a-probes at once.
// COUNT adds the number of filter byte tests (`tests`) and k-mer lookups the
// REFERENCE performs on this strand: a-probes issued speculatively past a hit
// are not counted, so the totals equal the oracle's (SURVEY 8d N_probes).
template <bool COUNT>
__device__ __forceinline__ bool scan_strand(const uint32_t *__restrict__ filter,
                                            const uint4 *__restrict__ planes, uint64_t o,
                                            uint32_t npos, int k, int t, uint64_t mask, bool rev,
                                            unsigned int &tests, unsigned int &lookups)
{
    constexpr int U = kSearchBatch;
    uint64_t wi = o >> 5;
    uint4 q0 = planes[wi], q1 = planes[wi + 1], q2 = planes[wi + 2];
    int seen = 0;
    uint32_t p = 0;
    while (p < npos) {
        uint64_t b = o + p;
        uint64_t need = b >> 5;
        if (need != wi) {
            if (need - wi >= 3) {
                wi = need;
                q0 = planes[wi]; q1 = planes[wi + 1]; q2 = planes[wi + 2];
            } else {
                do {
                    q0 = q1; q1 = q2; q2 = planes[wi + 3]; wi++;
                } while (wi != need);
            }
        }
        uint32_t sh = (uint32_t)b & 31u;
        uint32_t rem = npos - p;
        uint32_t wv = __funnelshift_r(q0.w, q1.w, sh);
        uint32_t m = wv & ((rem >= (uint32_t)U) ? ((1u << U) - 1u) : ((1u << rem) - 1u));
        if (m == 0) {
            // no k-mer starts in this batch: jump to the next W bit among the 32 visible ones
            uint32_t vis = rem < 32u ? rem : 32u;
            uint32_t mv = (vis >= 32u) ? wv : (wv & ((1u << vis) - 1u));
            p += mv ? (uint32_t)(__ffs(mv) - 1) : vis;
            continue;
        }
        uint64_t hv = window64(q0.x, q1.x, q2.x, sh);
        uint64_t lv = window64(q0.y, q1.y, q2.y, sh);
        uint32_t av[U];
        uint64_t ka[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rev) ka[u] = ~(hv >> u) & mask;
            else ka[u] = __brevll(hv >> u) >> (64 - k);
            av[u] = 0;
            if ((m >> u) & 1u) av[u] = ld_probe_u32(filter + key_word(ka[u]));
        }
        bool hit = false;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!hit && ((m >> u) & 1u)) {
                unsigned int tt = 1;
                if (av[u] & key_bit(ka[u], 0)) {
                    Keys q = make_keys(hv >> u, lv >> u, k, mask, rev);
                    if (probe_bcd(filter, q, tt)) {
                        hit = true;
                        seen++;
                        p += (uint32_t)u + (uint32_t)k;
                    }
                }
                if (COUNT) { tests += tt; lookups++; }
            }
        }
        if (hit) {
            if (seen >= t) return true;
        } else {
            p += U;
        }
    }
    return false;
}

// Both strands of search_reads in ONE left-to-right pass.  The reference scans the forward strand to the end
// before it looks at the reverse-complement keys (search_reads.h:46-83); the tag it sets is
// "forward greedy count >= t OR reverse greedy count >= t", which does not depend on the order the two scans are
// evaluated in.  Both scans walk the same windows of the same planes (rv_add also goes left to right,
// hash_key.h:99-125), so as long as neither strand has a hit the lane probes a window's forward AND reverse
// a-keys together (2 x kSearchBatch independent DRAM probes in flight).  The first hit FOCUSES the scan on its
// strand: that strand alone follows its k-jumps (search_reads.h:53-60) to the end of the read; only if it ends
// below t hits does the other strand resume where it stopped.  A reverse-complement copy is then found after a
// few batches instead of after a full fruitless forward scan, and a forward copy wastes one batch of reverse
// probes.  Each strand keeps its own hit count and its own next position, so every strand's greedy count is
// exactly the reference's.
// K32: keys of at most 30 bits (filters of at most 512 MiB, among them the L2-resident ones): the plane windows, the keys
// and the mask are 32-bit values and two plane words are enough -- a third fewer registers, one more resident block.
template <bool K32> struct KeyType { typedef uint64_t type; };
template <> struct KeyType<true> { typedef uint32_t type; };
__device__ __forceinline__ uint64_t fwd_key(uint64_t v, int k) { return __brevll(v) >> (64 - k); }
__device__ __forceinline__ uint32_t fwd_key(uint32_t v, int k) { return __brev(v) >> (32 - k); }

// b, c, d of one position after its a-bit was found set, in the reference's order (bloom_filter.h:124-131)
template <class KT>
__device__ __forceinline__ bool probe_bcd_of(const uint32_t *__restrict__ filter, KT a, KT lw, int k, KT mask, bool rev)
{
    const KT b = rev ? (KT)(~lw & mask) : fwd_key(lw, k);
    if (!(ld_probe_u32(filter + key_word(b)) & key_bit(b, 1))) return false;
    const KT c = a ^ b;
    if (!(ld_probe_u32(filter + key_word(c)) & key_bit(c, 2))) return false;
    const KT d = a | b;
    return (ld_probe_u32(filter + key_word(d)) & key_bit(d, 3)) != 0;
}

template <int U, bool K32>
__device__ __forceinline__ bool scan_both(const uint32_t *__restrict__ filter, const uint4 *__restrict__ planes,
                                          uint64_t o, uint32_t npos, int k, int t, uint64_t mask64)
{
    typedef typename KeyType<K32>::type KT;
    const KT mask = (KT)mask64;
    uint64_t wi = o >> 5;
    uint4 q0 = planes[wi], q1 = planes[wi + 1], q2 = make_uint4(0u, 0u, 0u, 0u);
    if (!K32) q2 = planes[wi + 2];
    int seen_f = 0, seen_r = 0;
    uint32_t nf = 0, nr = 0;                         // next position of each strand (>= npos: strand finished)
    int focus = 0;                                   // 0: both strands, 1: forward only, 2: reverse only
    while (true) {
        const bool use_f = focus != 2 && nf < npos, use_r = focus != 1 && nr < npos;
        if (!use_f && !use_r) {
            if (focus == 0) return false;            // both strands scanned to the end
            focus = 0;                               // the focused strand ended below t: the other one resumes
            continue;
        }
        const uint32_t p = use_f && use_r ? (nf < nr ? nf : nr) : (use_f ? nf : nr);
        uint64_t b = o + p;
        uint64_t need = b >> 5;
        if (need != wi) {
            if (K32) {
                if (need == wi + 1) { q0 = q1; q1 = planes[wi + 2]; }
                else { q0 = planes[need]; q1 = planes[need + 1]; }      // a jump, or a resumed strand behind the window
                wi = need;
            } else if (need < wi || need - wi >= 3) {                   // a resumed strand may be behind the window
                wi = need;
                q0 = planes[wi]; q1 = planes[wi + 1]; q2 = planes[wi + 2];
            } else {
                do {
                    q0 = q1; q1 = q2; q2 = planes[wi + 3]; wi++;
                } while (wi != need);
            }
        }
        uint32_t sh = (uint32_t)b & 31u;
        uint32_t rem = npos - p;
        uint32_t wv = __funnelshift_r(q0.w, q1.w, sh);
        uint32_t m = wv & ((rem >= (uint32_t)U) ? ((1u << U) - 1u) : ((1u << rem) - 1u));
        if (m == 0) {
            // no k-mer starts in this batch: both active strands jump to the next W bit among the 32 visible ones
            uint32_t vis = rem < 32u ? rem : 32u;
            uint32_t mv = (vis >= 32u) ? wv : (wv & ((1u << vis) - 1u));
            const uint32_t to = p + (mv ? (uint32_t)(__ffs(mv) - 1) : vis);
            if (use_f && nf < to) nf = to;
            if (use_r && nr < to) nr = to;
            continue;
        }
        // positions of this batch each strand still has to look at
        const uint32_t mf = !use_f || nf >= p + U ? 0u : (nf > p ? (m & (~0u << (nf - p))) : m);
        const uint32_t mr = !use_r || nr >= p + U ? 0u : (nr > p ? (m & (~0u << (nr - p))) : m);
        const KT hv = K32 ? (KT)__funnelshift_r(q0.x, q1.x, sh) : (KT)window64(q0.x, q1.x, q2.x, sh);
        const KT lv = K32 ? (KT)__funnelshift_r(q0.y, q1.y, sh) : (KT)window64(q0.y, q1.y, q2.y, sh);
        uint32_t af[U], ar[U];
        KT kf[U], kr[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            kf[u] = fwd_key((KT)(hv >> u), k);
            kr[u] = (KT)(~(hv >> u) & mask);
            af[u] = ar[u] = 0;
            if ((mf >> u) & 1u) af[u] = ld_probe_u32(filter + key_word(kf[u]));
            if ((mr >> u) & 1u) ar[u] = ld_probe_u32(filter + key_word(kr[u]));
        }
        bool hit_f = false, hit_r = false;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!hit_f && ((mf >> u) & 1u) && (af[u] & key_bit(kf[u], 0)) &&
                probe_bcd_of<KT>(filter, kf[u], (KT)(lv >> u), k, mask, false)) { hit_f = true; seen_f++; nf = p + (uint32_t)u + (uint32_t)k; }
            if (!hit_r && ((mr >> u) & 1u) && (ar[u] & key_bit(kr[u], 0)) &&
                probe_bcd_of<KT>(filter, kr[u], (KT)(lv >> u), k, mask, true)) { hit_r = true; seen_r++; nr = p + (uint32_t)u + (uint32_t)k; }
        }
        // `seen >= t` is only looked at after a hit (search_reads.h:55-57): t <= 1 behaves as t = 1
        if ((hit_f && seen_f >= t) || (hit_r && seen_r >= t)) return true;
        if (use_f && !hit_f && nf < p + U) nf = p + U;      // this batch is settled for a strand without a hit
        if (use_r && !hit_r && nr < p + U) nr = p + U;
        if (focus == 0) focus = hit_f ? 1 : (hit_r ? 2 : 0);
    }
}

// search_reads (search_reads.h:34-87): one lane per read, grid-stride.
// counters[0] += newly found, counters[1] += reads scanned; with COUNT also
// counters[2] += filter byte tests, counters[3] += k-mer lookups (reference semantics).
// BOTH: 0 = the reference's order (forward scan, then reverse); > 0 = one pass over both strands with BOTH
// positions per strand and batch (scan_both)
// (compiled for 4 resident blocks per SM = 64 registers: measured against 3, 5 and 6 -- 85, 48 and 40 registers -- at
// k=33 and k=27, profiles/r02_search_occupancy_ab.txt; both directions lose, up to 1.6x at k=27)
template <bool COUNT, int BOTH, int MINB = 4, bool K32 = false>
__global__ void __launch_bounds__(256, MINB)
k_search(const uint32_t *__restrict__ filter, const uint4 *__restrict__ planes,
         const uint64_t *__restrict__ offs, uint64_t n_reads, int k, int t,
         uint32_t *__restrict__ tags, unsigned long long *__restrict__ counters,
         const uint32_t *__restrict__ sel)
{
    const uint64_t mask = (1ull << k) - 1;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned int found = 0, searched = 0, tests = 0, lookups = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        if (sel && !((sel[r >> 5] >> (r & 31)) & 1u)) continue;   // not in the input vector: fasta_file.h:143-152
        if ((tags[r >> 5] >> (r & 31)) & 1u) continue;        // file_manager.h:99
        searched++;
        uint64_t o = offs[r];
        uint64_t len = offs[r + 1] - o;
        if (len < (uint64_t)k) continue;
        uint32_t npos = (uint32_t)(len - k + 1);
        bool f;
        if (COUNT || !BOTH) {      // the reference's order: forward scan, then reverse (what the probe counters describe)
            f = scan_strand<COUNT>(filter, planes, o, npos, k, t, mask, false, tests, lookups);
            if (!f) f = scan_strand<COUNT>(filter, planes, o, npos, k, t, mask, true, tests, lookups);
        } else {
            f = scan_both<(BOTH > 0 ? BOTH : 1), K32>(filter, planes, o, npos, k, t, mask);
        }
        if (f) {
            atomicOr(&tags[r >> 5], 1u << (r & 31));
            found++;
        }
    }
    for (int d = 16; d; d >>= 1) {
        found += __shfl_xor_sync(0xffffffffu, found, d);
        searched += __shfl_xor_sync(0xffffffffu, searched, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (found) atomicAdd(&counters[0], (unsigned long long)found);
        if (searched) atomicAdd(&counters[1], (unsigned long long)searched);
    }
    if (COUNT) {
        unsigned long long t64 = tests, l64 = lookups;
        for (int d = 16; d; d >>= 1) {
            t64 += __shfl_xor_sync(0xffffffffu, t64, d);
            l64 += __shfl_xor_sync(0xffffffffu, l64, d);
        }
        if ((threadIdx.x & 31) == 0) {
            if (t64) atomicAdd(&counters[2], t64);
            if (l64) atomicAdd(&counters[3], l64);
        }
    }
}

// ---- search with dynamic read hand-out (A/B: COMMET_B200_SEARCH_DYNAMIC=1) ----
// k_search gives every thread ONE read; the lanes of a warp finish at different times (a copy is found after a few
// probes, a read without a shared k-mer costs 2(L-k+1)), and reads already tagged by an earlier chunk leave their
// lanes idle from the start: 9.9 of 32 lanes are active on average at C2.  That does not matter while the DRAM
// row-activation rate is the limit (k >= 28), it does when the filter is L2-resident (k <= 27).  Here a lane that
// finishes its read takes the next one: warps claim runs of kDynChunk reads from a global cursor and hand them to
// their free lanes by ballot rank; the scan of a read is the state machine below, one batch per step, the same
// probes in the same order as scan_both.
struct ScanState {
    uint64_t o, wi;
    uint4 q0, q1, q2;
    uint32_t npos, nf, nr;
    int seen_f, seen_r, focus;
};

__device__ __forceinline__ void scan_init(ScanState &s, const uint4 *__restrict__ planes, uint64_t o, uint32_t npos)
{
    s.o = o;
    s.npos = npos;
    s.wi = o >> 5;
    s.q0 = planes[s.wi]; s.q1 = planes[s.wi + 1]; s.q2 = planes[s.wi + 2];
    s.seen_f = s.seen_r = 0;
    s.nf = s.nr = 0;
    s.focus = 0;
}

// one iteration of scan_both's loop: 0 = go on, 1 = read found, 2 = both strands scanned without t hits
template <int U>
__device__ __forceinline__ int scan_step(const uint32_t *__restrict__ filter, const uint4 *__restrict__ planes,
                                         ScanState &s, int k, int t, uint64_t mask)
{
    const bool use_f = s.focus != 2 && s.nf < s.npos, use_r = s.focus != 1 && s.nr < s.npos;
    if (!use_f && !use_r) {
        if (s.focus == 0) return 2;
        s.focus = 0;                                     // the focused strand ended below t: the other one resumes
        return 0;
    }
    const uint32_t p = use_f && use_r ? (s.nf < s.nr ? s.nf : s.nr) : (use_f ? s.nf : s.nr);
    const uint64_t b = s.o + p;
    const uint64_t need = b >> 5;
    if (need != s.wi) {
        if (need < s.wi || need - s.wi >= 3) {
            s.wi = need;
            s.q0 = planes[s.wi]; s.q1 = planes[s.wi + 1]; s.q2 = planes[s.wi + 2];
        } else {
            do {
                s.q0 = s.q1; s.q1 = s.q2; s.q2 = planes[s.wi + 3]; s.wi++;
            } while (s.wi != need);
        }
    }
    const uint32_t sh = (uint32_t)b & 31u;
    const uint32_t rem = s.npos - p;
    const uint32_t wv = __funnelshift_r(s.q0.w, s.q1.w, sh);
    const uint32_t m = wv & ((rem >= (uint32_t)U) ? ((1u << U) - 1u) : ((1u << rem) - 1u));
    if (m == 0) {
        const uint32_t vis = rem < 32u ? rem : 32u;
        const uint32_t mv = (vis >= 32u) ? wv : (wv & ((1u << vis) - 1u));
        const uint32_t to = p + (mv ? (uint32_t)(__ffs(mv) - 1) : vis);
        if (use_f && s.nf < to) s.nf = to;
        if (use_r && s.nr < to) s.nr = to;
        return 0;
    }
    const uint32_t mf = !use_f || s.nf >= p + U ? 0u : (s.nf > p ? (m & (~0u << (s.nf - p))) : m);
    const uint32_t mr = !use_r || s.nr >= p + U ? 0u : (s.nr > p ? (m & (~0u << (s.nr - p))) : m);
    const uint64_t hv = window64(s.q0.x, s.q1.x, s.q2.x, sh);
    const uint64_t lv = window64(s.q0.y, s.q1.y, s.q2.y, sh);
    uint32_t af[U], ar[U];
    uint64_t kf[U], kr[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        kf[u] = __brevll(hv >> u) >> (64 - k);
        kr[u] = ~(hv >> u) & mask;
        af[u] = ar[u] = 0;
        if ((mf >> u) & 1u) af[u] = ld_probe_u32(filter + key_word(kf[u]));
        if ((mr >> u) & 1u) ar[u] = ld_probe_u32(filter + key_word(kr[u]));
    }
    bool hit_f = false, hit_r = false;
    unsigned int dummy = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (!hit_f && ((mf >> u) & 1u) && (af[u] & key_bit(kf[u], 0))) {
            Keys q = make_keys(hv >> u, lv >> u, k, mask, false);
            if (probe_bcd(filter, q, dummy)) { hit_f = true; s.seen_f++; s.nf = p + (uint32_t)u + (uint32_t)k; }
        }
        if (!hit_r && ((mr >> u) & 1u) && (ar[u] & key_bit(kr[u], 0))) {
            Keys q = make_keys(hv >> u, lv >> u, k, mask, true);
            if (probe_bcd(filter, q, dummy)) { hit_r = true; s.seen_r++; s.nr = p + (uint32_t)u + (uint32_t)k; }
        }
    }
    if ((hit_f && s.seen_f >= t) || (hit_r && s.seen_r >= t)) return 1;
    if (use_f && !hit_f && s.nf < p + U) s.nf = p + U;
    if (use_r && !hit_r && s.nr < p + U) s.nr = p + U;
    if (s.focus == 0) s.focus = hit_f ? 1 : (hit_r ? 2 : 0);
    return 0;
}

constexpr unsigned kDynChunk = 256;          // reads a warp claims at a time

template <int U, int BPS>
__global__ void __launch_bounds__(256, BPS)
k_search_dyn(const uint32_t *__restrict__ filter, const uint4 *__restrict__ planes,
             const uint64_t *__restrict__ offs, uint64_t n_reads, int k, int t,
             uint32_t *__restrict__ tags, unsigned long long *__restrict__ counters,
             const uint32_t *__restrict__ sel, unsigned long long *__restrict__ cursor)
{
    const uint64_t mask = (1ull << k) - 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    uint64_t cur = 0, end = 0, r = 0;                  // cur/end: the warp's claimed run (warp-uniform)
    bool exhausted = false, active = false;
    ScanState s;
    unsigned int found = 0, searched = 0;
    while (true) {
        // hand the next reads of the run to the lanes without one, in lane order
        while (true) {
            const unsigned need = __ballot_sync(0xffffffffu, !active);
            if (!need) break;
            if (cur >= end) {
                if (exhausted) break;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kDynChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_reads) { exhausted = true; break; }
                cur = base;
                end = base + kDynChunk < n_reads ? base + kDynChunk : n_reads;
            }
            const uint64_t avail = end - cur;
            const unsigned rank = __popc(need & lt), cnt = __popc(need);
            if (!active && rank < avail) {
                const uint64_t rr = cur + rank;
                const bool selected = !sel || ((sel[rr >> 5] >> (rr & 31)) & 1u);        // fasta_file.h:143-152
                if (selected && !((tags[rr >> 5] >> (rr & 31)) & 1u)) {                  // file_manager.h:99
                    searched++;
                    const uint64_t o = offs[rr];
                    const uint64_t len = offs[rr + 1] - o;
                    if (len >= (uint64_t)k) {
                        scan_init(s, planes, o, (uint32_t)(len - k + 1));
                        r = rr;
                        active = true;
                    }
                }
            }
            cur += cnt < avail ? cnt : avail;
        }
        if (!__any_sync(0xffffffffu, active)) break;   // no read left to claim and none in flight
        if (active) {
            const int st = scan_step<U>(filter, planes, s, k, t, mask);
            if (st) {
                active = false;
                if (st == 1) {
                    atomicOr(&tags[r >> 5], 1u << (r & 31));
                    found++;
                }
            }
        }
    }
    for (int d = 16; d; d >>= 1) {
        found += __shfl_xor_sync(0xffffffffu, found, d);
        searched += __shfl_xor_sync(0xffffffffu, searched, d);
    }
    if (lane == 0) {
        if (found) atomicAdd(&counters[0], (unsigned long long)found);
        if (searched) atomicAdd(&counters[1], (unsigned long long)searched);
    }
}

}  // namespace commet
