// stage 1: the filter of a chunk -- direct RED.OR, the L2-blocked insert in its two forms, region passes (bloom_filter.h:112-121, index_reads.h:41-63)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"

namespace commet {

// ------------------------------------------------------- stage 1: index ----
// index_reads inner loop (index_reads.h:52-58) + BloomFilter::feed
// (bloom_filter.h:112-118), flat over stream positions [b0, b1): one warp per
// 32-position word, plane words are warp-uniform (broadcast) loads, each lane
// owns one k-mer start and issues four fire-and-forget 32-bit RED.OR.
__global__ void __launch_bounds__(256)
k_index(uint32_t *__restrict__ filter, const uint4 *__restrict__ planes, uint64_t b0, uint64_t b1,
        int k, unsigned long long *__restrict__ n_kmers)
{
    const uint64_t mask = (1ull << k) - 1;
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    uint64_t w_first = b0 >> 5, w_end = (b1 + 31) >> 5;
    unsigned long long local = 0;
    for (uint64_t wi = w_first + warp; wi < w_end; wi += n_warps) {
        uint4 q0 = planes[wi];
        uint32_t W = q0.w;
        uint64_t lo = wi << 5;
        if (lo < b0) W &= ~0u << (b0 - lo);
        if (lo + 32 > b1) W &= ~0u >> (lo + 32 - b1);
        if (W == 0) continue;                       // warp-uniform
        uint4 q1 = planes[wi + 1], q2 = planes[wi + 2];
        if ((W >> lane) & 1u) {
            uint64_t hv = window64(q0.x, q1.x, q2.x, lane);
            uint64_t lv = window64(q0.y, q1.y, q2.y, lane);
            Keys q = make_keys(hv, lv, k, mask, false);
            atomicOr(filter + key_word(q.a), key_bit(q.a, 0));
            atomicOr(filter + key_word(q.b), key_bit(q.b, 1));
            atomicOr(filter + key_word(q.c), key_bit(q.c, 2));
            atomicOr(filter + key_word(q.d), key_bit(q.d, 3));
        }
        local += __popc(W);
    }
    if (lane == 0 && local && n_kmers) atomicAdd(n_kmers, local);
}

// ------------------------------------------- stage 1, L2-blocked variant ----
// A DRAM-resident filter (k >= 28: 2^(k-1) bytes > L2) takes random RED.OR at
// the DRAM random-sector rate (~20 G/s measured).  Instead the key stream is
// first partitioned by filter REGION (2^kRegionLog2 bytes, L2-sized), then
// applied region after region so that every RED.OR hits L2:
//   k_bin_count   : per-region record counts            (streaming read)
//   k_bin_scan    : exclusive prefix -> region offsets   (1 block)
//   k_bin_scatter : 32-bit records, region-contiguous    (streaming write)
//   k_bin_apply   : tiles consumed in region order, RED.OR into L2-resident words
// record = low (kRegionLog2+1) key bits | j << (kRegionLog2+1); the region is
// the remaining high key bits, i.e. a function of the k-mer's FIRST bases.
constexpr int kRegionLog2 = 25;                 // 32 MiB regions
constexpr int kRecKeyBits = kRegionLog2 + 1;    // byte offset in region + odd/even bit
constexpr int kMaxBins = 512;
constexpr uint32_t kRecMask = (1u << kRecKeyBits) - 1u;
constexpr int kScatThreads = 512;
constexpr int kScatIters = 4;                                      // 32-position words per warp per tile
constexpr int kScatTileWords = (kScatThreads / 32) * kScatIters;   // 64 words = 2048 stream positions
constexpr int kScatTileRecs = kScatTileWords * 32 * 4;             // <= 8192 records per tile

__device__ __forceinline__ uint64_t ld_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}

// Region and record of the forward keys without building the 64-bit keys: with hw/lw = 64 plane bits from
// the k-mer's first base (bit 0), the TOP key bits are the first bases and the LOW key bits the last ones,
// both bit-reversed; c = a^b and d = a|b commute with taking bit fields.
//   bin  = key >> kRecKeyBits        = brev(first 32 bases) >> (32 - (k - kRecKeyBits))
//   rec  = key & kRecMask            = brev(32 bases from base k - kRecKeyBits) >> (32 - kRecKeyBits)
struct BinKeys { uint32_t bin[4], rec[4]; };

__device__ __forceinline__ void bin_only(uint64_t hw, uint64_t lw, int k, uint32_t bin[4])
{
    const int top = 32 - (k - kRecKeyBits);
    bin[0] = __brev((uint32_t)hw) >> top;
    bin[1] = __brev((uint32_t)lw) >> top;
    bin[2] = bin[0] ^ bin[1];
    bin[3] = bin[0] | bin[1];
}

__device__ __forceinline__ BinKeys bin_keys(uint64_t hw, uint64_t lw, int k)
{
    BinKeys q;
    bin_only(hw, lw, k, q.bin);
    const int skip = k - kRecKeyBits;                       // bases above the record bits (2..9)
    uint32_t ra = __brev((uint32_t)(hw >> skip)) >> (32 - kRecKeyBits);
    uint32_t rb = __brev((uint32_t)(lw >> skip)) >> (32 - kRecKeyBits);
    q.rec[0] = ra;
    q.rec[1] = rb | (1u << kRecKeyBits);
    q.rec[2] = (ra ^ rb) | (2u << kRecKeyBits);
    q.rec[3] = (ra | rb) | (3u << kRecKeyBits);
    return q;
}

// W word of stream word `wi`, restricted to positions [b0, b1)
__device__ __forceinline__ uint32_t w_in_range(uint32_t W, uint64_t wi, uint64_t b0, uint64_t b1)
{
    uint64_t lo = wi << 5;
    if (lo < b0) W &= ~0u << (b0 - lo);
    if (lo + 32 > b1) W &= ~0u >> (lo + 32 - b1);
    return W;
}

// Histogram of the four keys' regions.  c = a^b and d = a|b are functions of (a, b), so the block counts the
// PAIR (region of a, region of b) -- one shared-memory atomic per k-mer instead of four, spread over n_bins^2
// counters instead of n_bins -- and folds the pair table into the four marginals at the end.  JOINT needs
// n_bins^2 counters of dynamic shared memory (n_bins <= 128: 64 KB); larger bin counts use four atomics.
template <bool JOINT>
__global__ void __launch_bounds__(256)
k_bin_count(const uint4 *__restrict__ planes, uint64_t b0, uint64_t b1, int k, int n_bins,
            unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned int sh_cnt[];          // JOINT: n_bins^2 pair counters, then n_bins marginals
    const int n_cnt = JOINT ? n_bins * n_bins : n_bins;
    unsigned int *marg = JOINT ? sh_cnt + n_cnt : sh_cnt;
    for (int i = threadIdx.x; i < n_cnt + (JOINT ? n_bins : 0); i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t w_first = b0 >> 5, w_end = (b1 + 31) >> 5;
    // a warp takes 32 consecutive plane words per step: ONE coalesced 512-byte load (lane i holds word i, lane 0
    // also the first word of the next group), then the words are handed round by shuffles -- a warp-uniform
    // 16-byte load per word keeps a single request in flight per warp and is latency-bound (155 GB/s measured)
    for (uint64_t g0 = w_first + warp * 32; g0 < w_end; g0 += n_warps * 32) {
        const uint64_t wi = g0 + lane;
        uint4 mine = make_uint4(0u, 0u, 0u, 0u), extra = mine;
        if (wi <= w_end) mine = planes[wi];                          // planes carry 4 zero words past the end
        if (lane == 0 && g0 + 32 <= w_end) extra = planes[g0 + 32];
        const uint32_t ex = __shfl_sync(0xffffffffu, extra.x, 0), ey = __shfl_sync(0xffffffffu, extra.y, 0);
        uint32_t q0x = __shfl_sync(0xffffffffu, mine.x, 0), q0y = __shfl_sync(0xffffffffu, mine.y, 0);
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const uint32_t q1x = j == 31 ? ex : __shfl_sync(0xffffffffu, mine.x, (j + 1) & 31);
            const uint32_t q1y = j == 31 ? ey : __shfl_sync(0xffffffffu, mine.y, (j + 1) & 31);
            uint32_t W = __shfl_sync(0xffffffffu, mine.w, j);
            if (g0 + j < w_end) W = w_in_range(W, g0 + j, b0, b1); else W = 0;
            if ((W >> lane) & 1u) {
                uint32_t bin[4];
                bin_only(__funnelshift_r(q0x, q1x, lane), __funnelshift_r(q0y, q1y, lane), k, bin);
                if (JOINT) atomicAdd(&sh_cnt[bin[0] * n_bins + bin[1]], 1u);
                else {
                    atomicAdd(&sh_cnt[bin[0]], 1u);
                    atomicAdd(&sh_cnt[bin[1]], 1u);
                    atomicAdd(&sh_cnt[bin[2]], 1u);
                    atomicAdd(&sh_cnt[bin[3]], 1u);
                }
            }
            q0x = q1x;
            q0y = q1y;
        }
    }
    __syncthreads();
    if (JOINT) {
        for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) {
            unsigned int c = sh_cnt[i];
            if (c) {
                unsigned int x = i / n_bins, y = i - x * n_bins;
                atomicAdd(&marg[x], c);
                atomicAdd(&marg[y], c);
                atomicAdd(&marg[x ^ y], c);
                atomicAdd(&marg[x | y], c);
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (marg[i]) atomicAdd(&hist[i], (unsigned long long)marg[i]);
}

// base[b] = exclusive prefix of hist, base[n_bins] = total; cursor[b] = base[b]; tile counter reset
__global__ void k_bin_scan(const unsigned long long *__restrict__ hist, int n_bins,
                           unsigned long long *__restrict__ base, unsigned long long *__restrict__ cursor,
                           unsigned long long *__restrict__ tile_counter)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < n_bins; b++) {
            base[b] = acc;
            cursor[b] = acc;
            acc += hist[b];
        }
        base[n_bins] = acc;
        *tile_counter = 0;
    }
}

// One tile = 2048 stream positions.  Single pass over the keys: the shared-memory atomic that counts a
// region also hands the record its rank inside the tile's run for that region; (region, rank) and the
// record stay in registers across the block-wide scan, then records are placed region-sorted in shared
// memory and copied out with one coalesced store per record slot.
struct ScatterSmem {
    uint32_t stage[kScatTileRecs];            // 32 KB region-sorted records
    uint16_t sbin[kScatTileRecs];             // 16 KB region of every staged record
    unsigned long long delta[kMaxBins];       // global slot of the region's run minus its offset in `stage`
    uint32_t cnt[kMaxBins], start[kMaxBins];
    uint32_t wsum[kScatThreads / 32];
};

__global__ void __launch_bounds__(kScatThreads, 2)
k_bin_scatter(const uint4 *__restrict__ planes, uint64_t b0, uint64_t b1, int k, int n_bins,
              unsigned long long *__restrict__ cursor, uint32_t *__restrict__ recs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem &sm = *reinterpret_cast<ScatterSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t pol = ld_policy_evict_first();
    const uint64_t w_first = b0 >> 5, w_end = (b1 + 31) >> 5;
    const uint64_t n_tiles = (w_end - w_first + kScatTileWords - 1) / kScatTileWords;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = tid; i < n_bins; i += kScatThreads) sm.cnt[i] = 0;
        __syncthreads();
        uint32_t rec[4 * kScatIters], br[4 * kScatIters];      // record, region << 16 | rank (~0: none)
#pragma unroll
        for (int it = 0; it < kScatIters; it++) {
            uint64_t wi = w_first + tile * kScatTileWords + (uint64_t)it * (kScatThreads / 32) + warp;
            uint32_t W = 0;
            uint4 q0 = make_uint4(0u, 0u, 0u, 0u);
            if (wi < w_end) {                                   // warp-uniform
                q0 = planes[wi];
                W = w_in_range(q0.w, wi, b0, b1);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) br[4 * it + j] = ~0u;
            if (W != 0) {                                       // warp-uniform
                uint4 q1 = planes[wi + 1], q2 = planes[wi + 2];
                if ((W >> lane) & 1u) {
                    BinKeys q = bin_keys(window64(q0.x, q1.x, q2.x, lane), window64(q0.y, q1.y, q2.y, lane), k);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        rec[4 * it + j] = q.rec[j];
                        br[4 * it + j] = (q.bin[j] << 16) | atomicAdd(&sm.cnt[q.bin[j]], 1u);
                    }
                }
            }
        }
        __syncthreads();
        // exclusive scan of cnt over the regions (n_bins <= 512 = one per thread) + global reservation
        {
            uint32_t c = (int)tid < n_bins ? sm.cnt[tid] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += v;
            }
            if (lane == 31) sm.wsum[warp] = incl;
            __syncthreads();
            uint32_t off = 0;
            for (uint32_t w = 0; w < warp; w++) off += sm.wsum[w];
            uint32_t excl = off + incl - c;
            if ((int)tid < n_bins) {
                sm.start[tid] = excl;
                if (c) sm.delta[tid] = atomicAdd(&cursor[tid], (unsigned long long)c) - excl;
            }
        }
        __syncthreads();
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kScatThreads / 32; w++) total += sm.wsum[w];
#pragma unroll
        for (int i = 0; i < 4 * kScatIters; i++) {
            if (br[i] != ~0u) {
                uint32_t bin = br[i] >> 16;
                uint32_t pos = sm.start[bin] + (br[i] & 0xFFFFu);
                sm.stage[pos] = rec[i];
                sm.sbin[pos] = (uint16_t)bin;
            }
        }
        __syncthreads();
        for (uint32_t j = tid; j < total; j += kScatThreads)
            st_stream_u32(recs + (sm.delta[sm.sbin[j]] + j), sm.stage[j], pol);
        // no barrier here: the next tile's first barrier (after zeroing cnt) orders these reads of
        // stage/sbin/delta before any of its writes to them
    }
}

// Tiles of region-sorted records are taken in order from a global counter, so at any time the running
// blocks touch one or two regions: the RED.OR hit L2.  The next tile is claimed while the current one is
// processed (the claimer also looks up the region of the tile's first record, once per tile instead of a
// binary search per thread), and every thread has its 16-byte record loads in flight before the first RED.
template <int TILE, bool PREFETCH>
__global__ void __launch_bounds__(256)
k_bin_apply(uint32_t *__restrict__ filter, const uint32_t *__restrict__ recs,
            const unsigned long long *__restrict__ base, int n_bins,
            unsigned long long *__restrict__ tile_counter)
{
    constexpr int U = TILE / (256 * 4);                    // 16-byte loads per thread per tile
    __shared__ unsigned long long sbase[kMaxBins + 1];
    __shared__ unsigned long long s_next;
    __shared__ int s_bin;
    for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) sbase[i] = base[i];
    __syncthreads();
    const unsigned long long total = sbase[n_bins];
    const unsigned long long n_tiles = (total + TILE - 1) / TILE;
    const uint64_t pol = ld_policy_evict_first();
    auto bin_of = [&](unsigned long long first) {           // last region with base <= first
        int lo = 0, hi = n_bins - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (sbase[mid] <= first) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    if (threadIdx.x == 0) {
        unsigned long long tl = atomicAdd(tile_counter, 1ull);
        s_next = tl;
        s_bin = tl < n_tiles ? bin_of(tl * TILE) : 0;
    }
    __syncthreads();
    unsigned long long tl = s_next;
    int bin0 = s_bin;
    while (tl < n_tiles) {
        __syncthreads();                                   // everybody holds tl/bin0: the slots may be overwritten
        unsigned long long nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(tile_counter, 1ull);      // consumed after this tile
        const unsigned long long t0 = tl * TILE;
        uint4 v[U];
#pragma unroll
        for (int it = 0; it < U; it++) {
            unsigned long long idx = t0 + ((unsigned long long)it * 256 + threadIdx.x) * 4;
            v[it] = make_uint4(0u, 0u, 0u, 0u);
            if (idx < total) v[it] = ld_stream_u4(reinterpret_cast<const uint4 *>(recs + idx), pol);   // recs is padded to 16 B
        }
        int bin = bin0;
        unsigned long long lim = sbase[bin + 1];
        if (PREFETCH && bin0 + 1 < n_bins) {
            // A RED that misses L2 is served at the DRAM random-access rate, and every line of a region misses
            // once per sweep (measured: ~3-4.5 ms of every apply pass, whatever the number of records).  The
            // tiles of region b therefore pull region b+1 into L2 ahead of its first RED, each tile an equal
            // slice, as sequential line prefetches -- unless b+1 receives too few records to touch most lines.
            const unsigned long long rb = sbase[bin0], re = sbase[bin0 + 1], ne = sbase[bin0 + 2];
            if (ne - re >= (1ull << (kRegionLog2 - 8))) {                       // >= half a record per line
                const unsigned long long tf = rb / TILE, n_t = (re - 1) / TILE - tf + 1, rel = tl - tf;
                const unsigned long long lines = 1ull << (kRegionLog2 - 7);
                const unsigned long long l0 = lines * rel / n_t, l1 = lines * (rel + 1) / n_t;
                const char *nxt_region = reinterpret_cast<const char *>(filter) + ((uint64_t)(bin0 + 1) << kRegionLog2);
                for (unsigned long long l = l0 + threadIdx.x; l < l1; l += 256)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(nxt_region + (l << 7)));
            }
        }
#pragma unroll
        for (int it = 0; it < U; it++) {
            unsigned long long idx = t0 + ((unsigned long long)it * 256 + threadIdx.x) * 4;
            uint32_t r[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                unsigned long long i = idx + e;
                if (i >= total) break;
                while (i >= lim) lim = sbase[++bin + 1];   // a tile rarely spans more than two regions
                uint32_t key_low = r[e] & kRecMask;
                uint64_t word = ((uint64_t)bin << (kRegionLog2 - 2)) + (key_low >> 3);
                atomicOr(filter + word, key_bit((uint64_t)key_low, (int)(r[e] >> kRecKeyBits)));
            }
        }
        if (threadIdx.x == 0) {
            s_next = nxt;
            s_bin = nxt < n_tiles ? bin_of(nxt * TILE) : 0;
        }
        __syncthreads();
        tl = s_next;
        bin0 = s_bin;
    }
}

// ------------------------------- stage 1, L2-blocked variant, second form ----
// Same idea (records partitioned by filter region, then applied region after region), without the histogram
// pass and with half the shared-memory traffic and instructions per record:
//   * no k_bin_count: a region's records live in SLABS of 2^kSlabLog2 records handed out on demand.  A tile's
//     run for region b reserves `cnt` places with one atomicAdd on fill[b] (the region's virtual record stream);
//     the block whose reservation covers the first place of a slab allocates it (atomicAdd on the slab counter)
//     and publishes its id in table[b][slab]; blocks whose runs land in a slab they did not open wait for the
//     id.  The opener has already executed its atomicAdd -- it is resident and publishes before it waits for
//     anything itself -- so the wait always ends.
//   * a tile's plane words are loaded once into shared memory; the keys are generated twice from there (first
//     only their regions, for the per-region counts; then in full, taking their place in the region-sorted tile
//     from a shared-memory cursor) instead of being carried in registers across the scan: no per-record
//     register state, so tiles can be larger (longer runs per region) and more blocks fit an SM.
//   * copy-out by region run: a warp copies whole runs (coalesced), no per-record region lookup.
//   * k_bin_apply2 takes tiles that never span two regions (per-region tile ranges from fill[]), so the inner
//     loop has no region walk, and all index arithmetic inside a tile is 32-bit.
constexpr int kSlabLog2 = 18;                        // records per slab (1 MiB)
constexpr uint32_t kSlabRecs = 1u << kSlabLog2;
constexpr int kS2Threads = 512;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// dynamic shared memory of k_bin_scatter2<TW>: stage[TW*128] | planes[TW+2] (uint4) | 6 arrays of n_bins u32 | wsum[16]
__host__ __device__ inline size_t scatter2_smem_bytes(int tw, int n_bins)
{
    return (size_t)tw * 128 * 4 + (size_t)(tw + 2) * 16 + (size_t)6 * n_bins * 4 + 16 * 4;
}

// fill[b]: records reserved for region b so far; table[b * max_q + q]: 1 + id of the q-th slab of region b
// (0: not opened yet); n_slabs: slabs handed out.  All zeroed by the host before the launch.
template <int TW>
__global__ void __launch_bounds__(kS2Threads, (TW <= 96 ? 3 : 2))
k_bin_scatter2(const uint4 *__restrict__ planes, uint64_t b0, uint64_t b1, int k, int n_bins,
               uint32_t *__restrict__ fill, uint32_t *__restrict__ table, uint32_t max_q,
               uint32_t *__restrict__ n_slabs, uint32_t max_slabs, uint32_t *__restrict__ recs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw);
    uint4 *pl = reinterpret_cast<uint4 *>(smem_raw + (size_t)TW * 128 * 4);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(pl + TW + 2);
    uint32_t *cur = cnt + n_bins, *start = cur + n_bins, *gpos = start + n_bins, *sid0 = gpos + n_bins,
             *sid1 = sid0 + n_bins, *wsum = sid1 + n_bins;
    constexpr int kWarps = kS2Threads / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t pol = ld_policy_evict_first();
    const uint64_t w_first = b0 >> 5, w_end = (b1 + 31) >> 5;
    const uint64_t n_tiles = (w_end - w_first + TW - 1) / TW;
    const int top = 32 - (k - kRecKeyBits);
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t w0 = w_first + tile * TW;
        // ---- the tile's plane words (+2 of halo), W restricted to [b0, b1) -----------------------------
        for (uint32_t i = tid; i < (uint32_t)TW + 2; i += kS2Threads) {
            const uint64_t wi = w0 + i;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (wi < w_end + 2) {                       // planes carry 4 zero words past the end
                q = ld_nc_u4(planes + wi);
                q.w = (wi < w_end && i < (uint32_t)TW) ? w_in_range(q.w, wi, b0, b1) : 0u;
            }
            pl[i] = q;
        }
        for (uint32_t i = tid; i < (uint32_t)n_bins; i += kS2Threads) cnt[i] = 0;
        __syncthreads();
        // ---- pass A: regions only -> per-region counts -----------------------------------------------------
#pragma unroll 2
        for (int wl = (int)warp; wl < TW; wl += kWarps) {
            const uint4 q0 = pl[wl];
            if (!((q0.w >> lane) & 1u)) continue;
            const uint4 q1 = pl[wl + 1];
            const uint32_t ba = __brev(__funnelshift_r(q0.x, q1.x, lane)) >> top;
            const uint32_t bb = __brev(__funnelshift_r(q0.y, q1.y, lane)) >> top;
            atomicAdd(&cnt[ba], 1u);
            atomicAdd(&cnt[bb], 1u);
            atomicAdd(&cnt[ba ^ bb], 1u);
            atomicAdd(&cnt[ba | bb], 1u);
        }
        __syncthreads();
        // ---- scan over the regions (one per thread), reservation in the regions' record streams, slabs ---------
        {
            const uint32_t c = (int)tid < n_bins ? cnt[tid] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += v;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t off = 0;
            for (uint32_t w = 0; w < warp; w++) off += wsum[w];
            if ((int)tid < n_bins) {
                const uint32_t excl = off + incl - c;
                start[tid] = excl;
                cur[tid] = excl;
                if (c) {
                    const uint32_t g = atomicAdd(&fill[tid], c);
                    gpos[tid] = g;
                    const uint32_t qa = g >> kSlabLog2, qb = (g + c - 1) >> kSlabLog2;
                    uint32_t *row = table + (size_t)tid * max_q;
                    uint32_t ida = 0, idb = 0;
                    // open the slabs whose first place this run covers -- before waiting for anything
                    // (a pool that is too small for the records -- the host sizes it from an upper bound -- is reported through
                    // n_slabs[1]; the records then land in slab 0: in bounds, and the call fails)
                    if ((g & (kSlabRecs - 1)) == 0) {
                        ida = atomicAdd(n_slabs, 1u) + 1;
                        if (ida > max_slabs) { n_slabs[1] = 1; ida = 1; }
                        st_release_u32(row + qa, ida);
                    }
                    if (qb != qa) {
                        idb = atomicAdd(n_slabs, 1u) + 1;
                        if (idb > max_slabs) { n_slabs[1] = 1; idb = 1; }
                        st_release_u32(row + qb, idb);
                    }
                    while (ida == 0) ida = ld_acquire_u32(row + qa);
                    sid0[tid] = ida - 1;
                    sid1[tid] = qb != qa ? idb - 1 : ida - 1;
                }
            }
        }
        __syncthreads();
        // ---- pass B: full records, placed region-sorted through the cursors ----------------------------------------
#pragma unroll 2
        for (int wl = (int)warp; wl < TW; wl += kWarps) {
            const uint4 q0 = pl[wl];
            if (!((q0.w >> lane) & 1u)) continue;
            const uint4 q1 = pl[wl + 1], q2 = pl[wl + 2];
            const BinKeys q = bin_keys(window64(q0.x, q1.x, q2.x, lane), window64(q0.y, q1.y, q2.y, lane), k);
#pragma unroll
            for (int j = 0; j < 4; j++) stage[atomicAdd(&cur[q.bin[j]], 1u)] = q.rec[j];
        }
        __syncthreads();
        // ---- copy-out: a warp per region run ----------------------------------------------------------------
        // (runs are ~100 records: the per-run overhead matters as much as the loop body -- no unrolling, one pointer
        // per slab; a run continues in a second slab once in 2^kSlabLog2 records)
        for (int b = (int)warp; b < n_bins; b += kWarps) {
            const uint32_t c = cnt[b];
            if (c == 0) continue;
            const uint32_t off = gpos[b] & (kSlabRecs - 1);
            const uint32_t n0 = min(c, kSlabRecs - off);                 // records that fit in the first slab
            const uint32_t *src = stage + start[b];
            uint32_t *dst = recs + (((size_t)sid0[b] << kSlabLog2) + off);
#pragma unroll 1
            for (uint32_t i = lane; i < n0; i += 32) st_stream_u32(dst + i, src[i], pol);
            if (n0 < c) {
                dst = recs + ((size_t)sid1[b] << kSlabLog2) - n0;
#pragma unroll 1
                for (uint32_t i = n0 + lane; i < c; i += 32) st_stream_u32(dst + i, src[i], pol);
            }
        }
        __syncthreads();                   // stage / cnt / planes are rewritten by the next tile
    }
}

// per-region tile ranges of k_bin_apply2: tbase[b] = first tile of region b, tbase[n_bins] = number of tiles
template <int TILE>
__global__ void k_bin_plan2(const uint32_t *__restrict__ fill, int n_bins, uint32_t *__restrict__ tbase,
                            unsigned long long *__restrict__ tile_counter)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < n_bins; b++) {
            tbase[b] = acc;
            acc += (fill[b] + TILE - 1) / TILE;
        }
        tbase[n_bins] = acc;
        *tile_counter = 0;
    }
}

// Tiles are taken in order from a global counter (region-major), every tile lies inside one region and one slab
// (TILE divides the slab size): the inner loop is load, mask, RED.
template <int TILE, bool PREFETCH>
__global__ void __launch_bounds__(256)
k_bin_apply2(uint32_t *__restrict__ filter, const uint32_t *__restrict__ recs, const uint32_t *__restrict__ fill,
             const uint32_t *__restrict__ tbase, const uint32_t *__restrict__ table, uint32_t max_q, int n_bins,
             unsigned long long *__restrict__ tile_counter)
{
    constexpr int U = TILE / (256 * 4);                    // 16-byte loads per thread per tile
    static_assert((kSlabRecs % TILE) == 0, "a tile must not span two slabs");
    __shared__ uint32_t sbase[kMaxBins + 1];
    __shared__ uint32_t sfill[kMaxBins];
    __shared__ uint32_t s_next;
    __shared__ int s_bin;
    for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) sbase[i] = tbase[i];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) sfill[i] = fill[i];
    __syncthreads();
    const uint32_t n_tiles = sbase[n_bins];
    const uint64_t pol = ld_policy_evict_first();
    auto bin_of = [&](uint32_t t) {                        // last region with tbase <= t (regions without tiles are skipped)
        int lo = 0, hi = n_bins - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (sbase[mid] <= t) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    if (threadIdx.x == 0) {
        unsigned long long tl = atomicAdd(tile_counter, 1ull);
        s_next = tl < n_tiles ? (uint32_t)tl : n_tiles;
        s_bin = tl < n_tiles ? bin_of((uint32_t)tl) : 0;
    }
    __syncthreads();
    uint32_t tl = s_next;
    int bin = s_bin;
    while (tl < n_tiles) {
        __syncthreads();                                   // everybody holds tl/bin: the slots may be overwritten
        unsigned long long nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(tile_counter, 1ull);      // consumed after this tile
        const uint32_t lt = tl - sbase[bin];                          // tile inside the region
        const uint32_t v0 = lt * TILE;                                // first record of the tile in the region's stream
        const uint32_t n_here = min((uint32_t)TILE, sfill[bin] - v0);
        const uint32_t slab = table[(size_t)bin * max_q + (v0 >> kSlabLog2)] - 1u;
        const uint4 *src = reinterpret_cast<const uint4 *>(recs + (((size_t)slab << kSlabLog2) + (v0 & (kSlabRecs - 1))));
        uint4 v[U];
#pragma unroll
        for (int it = 0; it < U; it++) {
            const uint32_t e = (it * 256 + threadIdx.x) * 4;
            v[it] = make_uint4(0u, 0u, 0u, 0u);
            if (e < n_here) v[it] = ld_stream_u4(src + (it * 256 + threadIdx.x), pol);      // slabs are whole: the tail of a vector is readable
        }
        if (PREFETCH && bin + 1 < n_bins) {
            // the tiles of region b pull region b+1 into L2 ahead of its first RED, each tile an equal slice, as
            // sequential line prefetches -- unless b+1 receives too few records to touch most of its lines
            const uint32_t nf = sfill[bin + 1];
            if (nf >= (1u << (kRegionLog2 - 8))) {                              // >= half a record per line
                const uint32_t n_t = sbase[bin + 1] - sbase[bin];
                const uint32_t lines = 1u << (kRegionLog2 - 7);
                const uint32_t l0 = (uint32_t)((uint64_t)lines * lt / n_t), l1 = (uint32_t)((uint64_t)lines * (lt + 1) / n_t);
                const char *nxt_region = reinterpret_cast<const char *>(filter) + ((uint64_t)(bin + 1) << kRegionLog2);
                for (uint32_t l = l0 + threadIdx.x; l < l1; l += 256)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(nxt_region + ((uint64_t)l << 7)));
            }
        }
        uint32_t *region = filter + ((uint64_t)bin << (kRegionLog2 - 2));
#pragma unroll
        for (int it = 0; it < U; it++) {
            const uint32_t e = (it * 256 + threadIdx.x) * 4;
            const uint32_t r[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
            for (int x = 0; x < 4; x++) {
                if (e + x < n_here) {
                    const uint32_t key_low = r[x] & kRecMask;
                    atomicOr(region + (key_low >> 3), key_bit((uint64_t)key_low, (int)(r[x] >> kRecKeyBits)));
                }
            }
        }
        if (threadIdx.x == 0) {
            s_next = nxt < n_tiles ? (uint32_t)nxt : n_tiles;
            s_bin = nxt < n_tiles ? bin_of((uint32_t)nxt) : 0;
        }
        __syncthreads();
        tl = s_next;
        bin = s_bin;
    }
}

// k_bin_apply2 with the record tiles brought in by the bulk-copy engine (cp.async.bulk, completion on an mbarrier)
// into a ring of shared-memory stages instead of by LDG.  In k_bin_apply2 a tile's loads are issued by the same LSU
// pipe that is draining thousands of queued RED lane-operations: the loads wait behind them, the warps wait for
// the loads (ncu: 43 long-scoreboard + 35 barrier stall cycles per issue, 150 G RED/s against a 218 G/s ceiling).
// Here one thread claims tiles and issues one bulk copy per tile (and one bulk L2 prefetch for the slice of the
// next region the tile is responsible for); all threads only read records from shared memory and issue REDs.
// A stage is released through an "empty" mbarrier (256 arrivals), so warps drift apart by up to STAGES tiles
// instead of meeting at a block barrier per tile.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}

template <int TILE, int STAGES, bool PREFETCH>
__global__ void __launch_bounds__(256)
k_bin_apply3(uint32_t *__restrict__ filter, const uint32_t *__restrict__ recs, const uint32_t *__restrict__ fill,
             const uint32_t *__restrict__ tbase, const uint32_t *__restrict__ table, uint32_t max_q, int n_bins,
             unsigned long long *__restrict__ tile_counter)
{
    constexpr int U = TILE / (256 * 4);                    // 16-byte vectors per thread per tile
    static_assert((kSlabRecs % TILE) == 0, "a tile must not span two slabs");
    extern __shared__ __align__(16) unsigned char smem_raw[];        // STAGES tiles of TILE records (bulk copies need 16-byte alignment)
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    __shared__ uint32_t s_bin[STAGES], s_n[STAGES];
    __shared__ uint32_t sbase[kMaxBins + 1];
    __shared__ uint32_t sfill[kMaxBins];
    for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) sbase[i] = tbase[i];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) sfill[i] = fill[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t n_tiles = sbase[n_bins];
    const uint64_t pol = ld_policy_evict_first();
    // thread 0: claim the next tile and start its copy into stage s (or mark the stage as the end of the work)
    auto produce = [&](int s) {
        const unsigned long long tl64 = atomicAdd(tile_counter, 1ull);
        if (tl64 >= n_tiles) {
            s_n[s] = 0;
            mbar_arrive(&full[s]);
            return;
        }
        const uint32_t tl = (uint32_t)tl64;
        int lo = 0, hi = n_bins - 1;                       // last region with tbase <= tl (regions without tiles are skipped)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sbase[mid] <= tl) lo = mid; else hi = mid - 1;
        }
        const int bin = lo;
        const uint32_t lt = tl - sbase[bin], v0 = lt * TILE;
        const uint32_t slab = table[(size_t)bin * max_q + (v0 >> kSlabLog2)] - 1u;
        s_bin[s] = (uint32_t)bin;
        s_n[s] = min((uint32_t)TILE, sfill[bin] - v0);
        mbar_arrive_expect_tx(&full[s], TILE * 4);
        bulk_g2s(smem_raw + (size_t)s * TILE * 4, recs + (((size_t)slab << kSlabLog2) + (v0 & (kSlabRecs - 1))), TILE * 4, &full[s], pol);
        if (PREFETCH && bin + 1 < n_bins && sfill[bin + 1] >= (1u << (kRegionLog2 - 8))) {
            // the tiles of region b pull region b+1 into L2 ahead of its first RED, each tile an equal slice
            const uint32_t n_t = sbase[bin + 1] - sbase[bin];
            const uint32_t lines = 1u << (kRegionLog2 - 7);
            const uint32_t l0 = (uint32_t)((uint64_t)lines * lt / n_t), l1 = (uint32_t)((uint64_t)lines * (lt + 1) / n_t);
            if (l1 > l0)
                bulk_prefetch_l2(reinterpret_cast<const char *>(filter) + ((uint64_t)(bin + 1) << kRegionLog2) + ((uint64_t)l0 << 7), (l1 - l0) << 7);
        }
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES; s++) produce(s);
    for (uint32_t it = 0;; it++) {
        const int s = (int)(it % STAGES);
        const uint32_t round = it / STAGES;
        mbar_wait(&full[s], round & 1u);
        const uint32_t n_here = s_n[s];
        if (n_here == 0) break;                            // tiles are claimed in order: every later stage is empty too
        uint32_t *region = filter + ((uint64_t)s_bin[s] << (kRegionLog2 - 2));
        const uint4 *src = reinterpret_cast<const uint4 *>(smem_raw + (size_t)s * TILE * 4);
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = src[u * 256 + threadIdx.x];
        mbar_arrive(&empty[s]);                            // this thread's records are in registers
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t e = (u * 256 + threadIdx.x) * 4;
            const uint32_t r[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int x = 0; x < 4; x++) {
                if (e + x < n_here) {
                    const uint32_t key_low = r[x] & kRecMask;
                    atomicOr(region + (key_low >> 3), key_bit((uint64_t)key_low, (int)(r[x] >> kRecKeyBits)));
                }
            }
        }
        if (threadIdx.x == 0) {
            mbar_wait(&empty[s], round & 1u);              // every thread has taken its records out of the stage
            produce(s);
        }
    }
}

// --------------------------------- stage 1, region-pass variant (no sort) ----
// The region of a key is its TOP bits = the k-mer's FIRST R bases, so "which k-mers of this 32-position
// word fall into region r" is a bit-parallel pattern match on the plane words: R funnel-shift + LOP3 pairs
// per key type, no key is built for the (2^R - 1)/2^R positions that miss.  The insert is then a sequence
// of 2^R passes over the stream, pass r inserting only the keys of region r (2^(k-1-R) bytes, L2-sized):
// every RED.OR hits L2, each key is still inserted exactly once, and there is no record buffer, no
// histogram and no scatter.  Work items are (region, tile) pairs taken region-major, so the blocks running
// at any time touch one or two regions.  Plane words are streamed with an evict-first policy so they do
// not push the region out of L2.
// MEASURED (profiles/r01_region_pass_diag.txt, C2): not the default.  The scan alone costs 0.53 ms per pass
// (35 ms for the 64 passes of k=33), and blocks drift by more than one pass, so two or three regions are
// live at once and the RED.OR fall back to DRAM rate: 104 ms against 32 ms for the sorted path.
constexpr int kPassTileWords = 256;                 // one 32-position word per thread per item

__device__ __forceinline__ uint4 ld_planes_stream(const uint4 *p, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

// positions s (bit s of the result) of this word whose first R bases spell region r in plane x:
// base s+i must equal bit R-1-i of r (first base most significant, hash_key.h:65-91)
__device__ __forceinline__ uint32_t match_region(uint32_t x0, uint32_t x1, uint32_t r, int R, uint32_t m)
{
    for (int i = 0; i < R; i++) {
        uint32_t inv = ((r >> (R - 1 - i)) & 1u) ? 0u : ~0u;
        m &= __funnelshift_r(x0, x1, i) ^ inv;
    }
    return m;
}

__global__ void __launch_bounds__(256)
k_index_regions(uint32_t *__restrict__ filter, const uint4 *__restrict__ planes, uint64_t b0, uint64_t b1,
                int k, int R)
{
    const uint64_t pol = ld_policy_evict_first();
    const uint64_t w_first = b0 >> 5, w_end = (b1 + 31) >> 5;
    const uint64_t n_tiles = (w_end - w_first + kPassTileWords - 1) / kPassTileWords;
    const int low_bits = k - R;                              // key bits below the region bits (<= 32)
    const uint32_t low_mask = low_bits >= 32 ? ~0u : ((1u << low_bits) - 1u);
    // items (region r, tile) are taken region-major by block index: item = r * n_tiles + tile
    uint32_t r = (uint32_t)(blockIdx.x / n_tiles);
    uint64_t tile = blockIdx.x % n_tiles;
    for (; r < (1u << R); ) {
        const uint64_t wi = w_first + tile * kPassTileWords + threadIdx.x;
        tile += gridDim.x;
        const uint32_t r_now = r;
        while (tile >= n_tiles) { tile -= n_tiles; r++; }
        if (wi >= w_end) continue;
        uint4 q0 = ld_planes_stream(planes + wi, pol);
        uint32_t W = w_in_range(q0.w, wi, b0, b1);
        if (W == 0) continue;
        uint4 q1 = ld_planes_stream(planes + wi + 1, pol);
        // key types a, b, c, d = planes H, L, H^L, H|L
        uint32_t x0[4] = {q0.x, q0.y, q0.x ^ q0.y, q0.x | q0.y};
        uint32_t x1[4] = {q1.x, q1.y, q1.x ^ q1.y, q1.x | q1.y};
        uint32_t M[4];
#pragma unroll
        for (int j = 0; j < 4; j++) M[j] = match_region(x0[j], x1[j], r_now, R, W);
        if ((M[0] | M[1] | M[2] | M[3]) == 0) continue;
        uint4 q2 = ld_planes_stream(planes + wi + 2, pol);
        uint32_t x2[4] = {q2.x, q2.y, q2.x ^ q2.y, q2.x | q2.y};
        uint32_t *region = filter + ((uint64_t)r_now << (low_bits - 3));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t m = M[j];
            while (m) {
                uint32_t sbit = __ffs(m) - 1;
                m &= m - 1;
                // low key bits = bases s+R .. s+k-1, last base least significant
                uint64_t w64 = window64(x0[j], x1[j], x2[j], sbit);
                uint32_t low = (__brev((uint32_t)(w64 >> R)) >> (32 - low_bits)) & low_mask;
                atomicOr(region + (low >> 3), key_bit((uint64_t)low, j));
            }
        }
    }
}

}  // namespace commet
