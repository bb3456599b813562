// Device code of the B200-native Commet hot path (sm_100a).
//
// Data layout in HBM:
//   planes : uint4 per 32 consecutive bases of a read stream
//            .x = H  bit-plane (1 for G,T)   -> key a   (hash_key.h:65-91)
//            .y = L  bit-plane (1 for C,T)   -> key b ; c = H^L ; d = H|L
//            .z = V  validity  (1 for ACGTacgt, alphabet.h:44-58)
//            .w = W  "a k-mer starts here" for the k the stream was prepared
//                    for: k valid bases that do not cross a read boundary
//            bit j of a word = base 32*word + j (LSB first).
//   filter : the bloom_filter.h byte array viewed as little-endian u32 words:
//            key -> word key>>3, bit 8*((key>>1)&3) + (key&1 ? 3-j : 7-j).
//   tags   : u32 words, bit r%32 of word r/32 = read r (= .bv payload bytes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace commet {

constexpr int kMaxK = 61;          // 64-bit window + batch of 4 positions
constexpr int kSearchBatch = 4;    // a-probes issued together per lane

// ---------------------------------------------------------------- loads ----
__device__ __forceinline__ uint32_t ld_nc_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// Filter probes: one random word per DRAM row activation.  The L2 fill of a probe miss is limited to 64 bytes
// (the smallest prefetch-size qualifier): same probe rate -- random probes are bound by DRAM row activations,
// 37.9 G lines/s measured, not by bytes -- but half the DRAM traffic of the default 128-byte fill
// (profiles/r01_ubench_sectors.txt: 63 B instead of 125 B per probe).
__device__ __forceinline__ uint32_t ld_probe_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_nc_u4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// 64 stream bits starting `sh` (0..31) bits into the 96-bit register window x0:x1:x2
__device__ __forceinline__ uint64_t window64(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t sh)
{
    uint32_t lo = __funnelshift_r(x0, x1, sh);
    uint32_t hi = __funnelshift_r(x1, x2, sh);
    return ((uint64_t)hi << 32) | lo;
}

// ----------------------------------------------------------------- keys ----
// hv/lv: k-mer window of the H and L planes, bit 0 = first base of the window.
// Forward keys (HashKey::add, hash_key.h:65-91): first base most significant.
// Reverse keys (HashKey::rv_add, hash_key.h:99-125): complement coding, first
// base least significant -> no bit reversal needed.
struct Keys { uint64_t a, b, c, d; };

__device__ __forceinline__ Keys make_keys(uint64_t hv, uint64_t lv, int k, uint64_t mask, bool rev)
{
    Keys q;
    if (rev) {
        q.a = ~hv & mask;
        q.b = ~lv & mask;
    } else {
        q.a = __brevll(hv) >> (64 - k);
        q.b = __brevll(lv) >> (64 - k);
    }
    q.c = q.a ^ q.b;
    q.d = q.a | q.b;
    return q;
}

// BloomFilter byte/mask (bloom_filter.h:112-131) in the u32-word view
__device__ __forceinline__ uint64_t key_word(uint64_t key) { return key >> 3; }
__device__ __forceinline__ uint32_t key_word(uint32_t key) { return key >> 3; }
__device__ __forceinline__ uint32_t key_bit(uint32_t key, int j)
{
    const uint32_t byte = (key >> 1) & 3u;
    const uint32_t in_byte = (key & 1u) ? (3 - j) : (7 - j);
    return 1u << (byte * 8 + in_byte);
}
__device__ __forceinline__ uint32_t key_bit(uint64_t key, int j)
{
    uint32_t byte = (uint32_t)(key >> 1) & 3u;
    uint32_t in_byte = (key & 1) ? (3 - j) : (7 - j);
    return 1u << (byte * 8 + in_byte);
}

// ------------------------------------------------------------- staging ----
// ASCII -> H/L/V planes, 32 bases per thread via two 16-byte vector loads.
// `bases` is zero-padded to a multiple of 32 bytes.
//
// Four bases (one 32-bit word c of the ASCII stream) at a time, every bit of interest brought to bit 7 of its byte by a
// LEFT shift (c7 needs none; a left shift can issue as a multiply on the FMA pipe, leaving the ALU pipe to the logic):
//   H = c2 (A,C -> 0; G,T -> 1), L = c1 ^ c2 (A,G -> 0; C,T -> 1)                             hash_key.h:65-91
//   V (alphabet.h:44-58, `ACGTacgt`): with q = c2 & ~c1 ("is T" among the four), a byte is valid iff
//       c7 = 0, c6 = 1, (c5 = case, ignored), c4 = q, c3 = 0, c0 = ~q
//     A 0100 0001   C 0100 0011   G 0100 0111   T 0101 0100
// The four flags of a word (bits 7, 15, 23, 31) are gathered into the top nibble of flags * 0x00204081 (partial
// products 7+21, 15+14, 23+7, 31+0 = bits 28..31; the other twelve land on distinct lower bits or above bit 31: no
// carries), and a funnel shift pushes that nibble into the plane word -- words taken last to first, so word j's
// nibble ends at bits 4j..4j+3.  Per word: 6 shifts, 6 three-input logic ops, 3 multiplies, 3 funnel shifts.
__device__ __forceinline__ void encode_word(uint32_t c, uint32_t &H, uint32_t &L, uint32_t &V)
{
    constexpr uint32_t M7 = 0x80808080u, K = 0x00204081u;
    const uint32_t x6 = c << 1, x4 = c << 3, x3 = c << 4, x2 = c << 5, x1 = c << 6, x0 = c << 7;
    const uint32_t q = x2 & ~x1;
    const uint32_t a = (x0 ^ q) & ~(x4 ^ q);
    const uint32_t b = ~x3 & x6 & ~c;
    const uint32_t v = a & b & M7;
    const uint32_t h = x2 & M7;
    const uint32_t l = (x1 ^ x2) & M7;
    H = __funnelshift_l(h * K, H, 4);
    L = __funnelshift_l(l * K, L, 4);
    V = __funnelshift_l(v * K, V, 4);
}

__device__ __forceinline__ void encode32(uint4 q0, uint4 q1, uint32_t &H, uint32_t &L, uint32_t &V)
{
    const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    H = L = V = 0;
#pragma unroll
    for (int j = 7; j >= 0; j--) encode_word(w[j], H, L, V);
}

__global__ void __launch_bounds__(256)
k_encode(const uint4 *__restrict__ bases16, uint4 *__restrict__ planes, uint64_t n_words)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        uint4 q0 = ld_nc_u4(bases16 + 2 * i);
        uint4 q1 = ld_nc_u4(bases16 + 2 * i + 1);
        uint32_t H, L, V;
        encode32(q0, q1, H, L, V);
        planes[i] = make_uint4(H, L, V, 0u);
    }
}

// offsets of a part of a larger stream -> offsets inside the part
__global__ void __launch_bounds__(256)
k_rebase(uint64_t *__restrict__ offs, uint64_t n, uint64_t base)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) offs[i] -= base;
}

// start-of-read marks: bit offs[r] of S for every read r
__global__ void __launch_bounds__(256)
k_mark_starts(const uint64_t *__restrict__ offs, uint64_t n_reads, uint32_t *__restrict__ S)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        uint64_t o = offs[r];
        atomicOr(&S[o >> 5], 1u << (o & 31));
    }
}

// W plane: position b starts a k-mer iff V[b..b+k) are all set and no read
// starts at b+1..b+k-1 (index_reads.h:52-58: hash.clear() per read and per
// non-ACGT char; a k-mer is fed once hash_size >= k).  One thread per 32
// positions: with A = V & ~S on a 96-bit window, W = V & AND_{d=1..k-1} A[b+d]
// is built from log2(k) shift-and-AND doublings instead of k tests.
__global__ void __launch_bounds__(256)
k_windows(uint4 *__restrict__ planes, const uint32_t *__restrict__ S, uint64_t n_words, int k)
{
    typedef unsigned __int128 u128;
    const int m = k - 1;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t *P32 = reinterpret_cast<uint32_t *>(planes);
    for (uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += stride) {
        uint32_t v0 = P32[4 * wi + 2], v1 = P32[4 * (wi + 1) + 2], v2 = P32[4 * (wi + 2) + 2];
        uint32_t a0 = v0 & ~S[wi], a1 = v1 & ~S[wi + 1], a2 = v2 & ~S[wi + 2];
        u128 A = (u128)a0 | ((u128)a1 << 32) | ((u128)a2 << 64);
        u128 P = A >> 1;                      // P[b] = A[b+1]
        u128 acc = ~(u128)0;
        int off = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            if (m & (1 << i)) {
                acc &= P >> off;
                off += 1 << i;
            }
            P &= P >> (1 << i);               // runs of 2^(i+1)
        }
        P32[4 * wi + 3] = v0 & (uint32_t)acc;  // V is 0 past the end of the stream: no k-mer crosses it
    }
}

// Read selection (the input boolean vector of a read file, fasta_file.h:143-152: reads whose bit is 0 are never
// handed out by get_next_read): the W bits of every position of an unselected read are cleared, so the flat
// per-position kernels (k-mer counts, insert) skip those reads without knowing about reads at all.
__global__ void __launch_bounds__(256)
k_mask_unselected(uint4 *__restrict__ planes, const uint64_t *__restrict__ offs, uint64_t n_reads,
                  const uint32_t *__restrict__ sel)
{
    uint32_t *P32 = reinterpret_cast<uint32_t *>(planes);
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        if ((sel[r >> 5] >> (r & 31)) & 1u) continue;
        uint64_t o = offs[r], e = offs[r + 1];
        for (uint64_t wi = o >> 5; (wi << 5) < e; wi++) {
            uint32_t m = ~0u;
            uint64_t lo = wi << 5;
            if (lo < o) m &= ~0u << (o - lo);
            if (lo + 32 > e) m &= ~0u >> (lo + 32 - e);
            if (m == ~0u) P32[4 * wi + 3] = 0u;               // the whole word belongs to this read
            else atomicAnd(&P32[4 * wi + 3], ~m);             // shared with a neighbouring read
        }
    }
}

// per-read k-mer count = popcount of W over the read's positions
__global__ void __launch_bounds__(256)
k_kmer_counts(const uint4 *__restrict__ planes, const uint64_t *__restrict__ offs,
              uint64_t n_reads, uint32_t *__restrict__ counts, unsigned long long *__restrict__ total)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    const uint32_t *P = reinterpret_cast<const uint32_t *>(planes);
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        uint64_t o = offs[r], e = offs[r + 1];
        uint32_t c = 0;
        for (uint64_t wi = o >> 5; (wi << 5) < e; wi++) {
            uint32_t W = P[4 * wi + 3];
            uint64_t lo = wi << 5;
            if (lo < o) W &= ~0u << (o - lo);
            if (lo + 32 > e) W &= ~0u >> (lo + 32 - e);
            c += __popc(W);
        }
        counts[r] = c;
        local += c;
    }
    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(total, local);
}

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

// ------------------------------------------------ stage 3: filter_reads ----
// classes: 0 selected, 1 too short, 2 too many N, 3 low Shannon, 4 undecided
// (|H - e| within the device/glibc log margin: resolved by the host from the
// exact counts written to `border`).
struct FilterParams {
    long long min_len;
    long long max_N;
    float min_shannon;
    float margin;
};
struct BorderRec { unsigned long long read; unsigned int cnt[5]; unsigned int len; };

__device__ __forceinline__ void base_counts(const uint4 *__restrict__ planes, uint64_t o, uint64_t e,
                                            unsigned int cnt[5])
{
    unsigned int a = 0, c = 0, g = 0, tt = 0;
    for (uint64_t wi = o >> 5; (wi << 5) < e; wi++) {
        uint4 q = planes[wi];
        uint32_t m = q.z;
        uint64_t lo = wi << 5;
        if (lo < o) m &= ~0u << (o - lo);
        if (lo + 32 > e) m &= ~0u >> (lo + 32 - e);
        a += __popc(~q.x & ~q.y & m);
        c += __popc(~q.x & q.y & m);
        g += __popc(q.x & ~q.y & m);
        tt += __popc(q.x & q.y & m);
    }
    cnt[0] = a; cnt[1] = c; cnt[2] = g; cnt[3] = tt;
    cnt[4] = (unsigned int)(e - o) - (a + c + g + tt);
}

// shannon_index (filter_reads.cpp:265-306): float freq, double term, float sum.
__device__ __forceinline__ float shannon_dev(const unsigned int cnt[5], unsigned int len)
{
    float idx = 0.f;
    const float flen = (float)len;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float f = __fdiv_rn((float)cnt[j], flen);
        if (f != 0.f) {
            double term = __ddiv_rn(__dmul_rn((double)f, log((double)f)), 0.6931471805599453);
            idx = __double2float_rn(__dadd_rn((double)idx, term));
        }
    }
    return fabsf(idx);
}

// single-precision estimate of the same index (MUFU.LG2): within 1e-5 of the reference's value, used to
// settle the reads that are nowhere near the threshold without the five double-precision logarithms
__device__ __forceinline__ float shannon_fast(const unsigned int cnt[5], unsigned int len)
{
    float idx = 0.f;
    const float inv = __frcp_rn((float)len);
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float f = (float)cnt[j] * inv;
        if (cnt[j]) idx = fmaf(f, __log2f(f), idx);
    }
    return fabsf(idx);
}

// the N and Shannon tests on a read's base counts (the length test comes first, filter_reads.cpp:189)
__device__ __forceinline__ int classify_counts(long long len, const unsigned int cnt[5], const FilterParams &fp)
{
    if ((long long)cnt[4] > fp.max_N) return 2;           // :192
    if (fp.min_shannon > 0.f) {                           // fabs() >= 0: e <= 0 never drops
        float hf = shannon_fast(cnt, (unsigned int)len);
        if (fabsf(hf - fp.min_shannon) > 1e-3f) return hf < fp.min_shannon ? 3 : 0;      // :195, decided 100x outside the error
        float h = shannon_dev(cnt, (unsigned int)len);
        if (fabsf(h - fp.min_shannon) <= fp.margin) return 4;
        if (h < fp.min_shannon) return 3;                 // :195
    }
    return 0;
}

__device__ __forceinline__ int classify_read(const uint4 *__restrict__ planes, uint64_t o, uint64_t e,
                                             const FilterParams &fp, unsigned int cnt[5])
{
    long long len = (long long)(e - o);
    if (len < fp.min_len) return 1;                       // filter_reads.cpp:189
    base_counts(planes, o, e, cnt);
    return classify_counts(len, cnt, fp);
}

// One block = 1024 consecutive reads.  Writes the selection bits (ballot,
// one store per 32 reads), the class of every read (1 byte, only when
// `classes` != null, i.e. when a -m cutoff must be located) and per-block
// class totals [4].
constexpr int kFilterBlock = 1024;

__global__ void __launch_bounds__(kFilterBlock)
k_filter(const uint4 *__restrict__ planes, const uint64_t *__restrict__ offs, uint64_t n_reads,
         FilterParams fp, uint32_t *__restrict__ bv, uint64_t n_bv_words,
         uint8_t *__restrict__ classes,
         unsigned int *__restrict__ block_totals, BorderRec *__restrict__ border,
         unsigned int border_cap, unsigned int *__restrict__ n_border)
{
    __shared__ unsigned int tot[4];
    if (threadIdx.x < 4) tot[threadIdx.x] = 0;
    __syncthreads();
    uint64_t r = (uint64_t)blockIdx.x * kFilterBlock + threadIdx.x;
    int cls = -1;
    if (r < n_reads) {
        unsigned int cnt[5];
        cls = classify_read(planes, offs[r], offs[r + 1], fp, cnt);
        if (cls == 4) {
            unsigned int slot = atomicAdd(n_border, 1u);
            if (slot < border_cap) {
                BorderRec br;
                br.read = r;
                for (int j = 0; j < 5; j++) br.cnt[j] = cnt[j];
                br.len = (unsigned int)(offs[r + 1] - offs[r]);
                border[slot] = br;
            }
            cls = 0;    // provisional; the host patches classes/bits/totals
        }
        if (classes) classes[r] = (uint8_t)cls;
    }
    uint32_t sel = __ballot_sync(0xffffffffu, cls == 0);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < n_bv_words) bv[r >> 5] = sel;   // padding bits stay 0
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t mc = __ballot_sync(0xffffffffu, cls == (c == 3 ? 0 : c + 1));
        if ((threadIdx.x & 31) == 0 && mc) atomicAdd(&tot[c], __popc(mc));
    }
    __syncthreads();
    if (threadIdx.x < 4) block_totals[4 * (uint64_t)blockIdx.x + threadIdx.x] = tot[threadIdx.x];
}

// The staging pass and the selection in ONE kernel (north_star stage 3): every ASCII base is read once (two 16-byte
// vector loads per 32-base word); its H/L/V bits are computed as k_encode does; the per-read A/C/G/T/other counts
// are popcounts of those bits restricted to the read's range -- no per-byte counting at all; with PLANES the bits
// are also stored as the stream's bit-planes, so a set that is filtered AND indexed is read from HBM once.
// `bases`: 16-byte aligned, readable up to `readable` bytes (a multiple of 16 >= the last offset = n_bases).
// The same fusion with k_encode's regularity.  A block takes 1024 consecutive reads (the unit of k_filter's outputs)
// and sweeps the WORDS of their span of the stream, a tile of sf2_tile_words() at a time: thread t encodes words t, t+1024,
// ... -- every word once, perfectly balanced, coalesced 32-byte loads -- into shared memory (and, with PLANES, into
// the stream's bit-planes: a word is stored by the block whose span holds its first byte); then thread t counts ITS read
// from the shared-memory bits of the part of the read that lies in the tile.  No per-read loop over global memory, no
// word encoded twice inside a block, no lane waiting for the longest read of its warp.
// THREADS reads per block, 1024 / THREADS blocks per SM (one sweeps while another counts); a tile of 4 * THREADS - 16
// words (12 bytes of shared memory each): 2032 words = 65 024 bases = 24 KB for 512 threads
template <int THREADS> __host__ __device__ constexpr int sf2_tile_words() { return 4 * THREADS - 16; }

template <bool PLANES, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_stage_filter(const uint8_t *__restrict__ bases, uint64_t readable, uint64_t n_bases, const uint64_t *__restrict__ offs,
                uint64_t n_reads, uint4 *__restrict__ planes, FilterParams fp, uint32_t *__restrict__ bv, uint64_t n_bv_words,
                uint8_t *__restrict__ classes, unsigned int *__restrict__ block_totals, BorderRec *__restrict__ border,
                unsigned int border_cap, unsigned int *__restrict__ n_border)
{
    constexpr int T_WORDS = sf2_tile_words<THREADS>();
    extern __shared__ uint32_t sf2_smem[];              // H[T] | L[T] | V[T]
    uint32_t *sH = sf2_smem, *sL = sH + T_WORDS, *sV = sL + T_WORDS;
    __shared__ unsigned int tot[4];
    __shared__ uint64_t s_span[2];
    const uint32_t tid = threadIdx.x;
    const uint64_t r_first = (uint64_t)blockIdx.x * THREADS, r = r_first + tid;
    if (tid < 4) tot[tid] = 0;
    if (tid == 0) {
        const uint64_t r_last = min(r_first + (uint64_t)THREADS, n_reads);
        s_span[0] = r_first < n_reads ? offs[r_first] : 0;
        s_span[1] = r_first < n_reads ? offs[r_last] : 0;
    }
    uint64_t my_o = 0, my_e = 0;
    if (r < n_reads) { my_o = offs[r]; my_e = offs[r + 1]; }
    __syncthreads();
    const uint64_t o_first = s_span[0], e_last = s_span[1];
    const uint64_t ws = o_first >> 5, we = (e_last + 31) >> 5;
    unsigned int cnt[5] = {0, 0, 0, 0, 0};
    for (uint64_t tw = ws; tw < we; tw += T_WORDS) {
        const uint32_t n_w = (uint32_t)min((uint64_t)T_WORDS, we - tw);
        for (uint32_t i = tid; i < n_w; i += THREADS) {
            const uint64_t w = tw + i, c = w << 5;
            const uint4 q0 = ld_nc_u4(reinterpret_cast<const uint4 *>(bases + c));
            uint4 q1 = make_uint4(0u, 0u, 0u, 0u);
            if (c + 16 < readable) q1 = ld_nc_u4(reinterpret_cast<const uint4 *>(bases + c + 16));
            uint32_t H, L, V;
            encode32(q0, q1, H, L, V);
            if (c + 32 > n_bases) V &= ~0u >> (c + 32 - n_bases);      // nothing is valid past the end of the stream
            sH[i] = H; sL[i] = L; sV[i] = V;
            if (PLANES && c >= o_first) planes[w] = make_uint4(H, L, V, 0u);
        }
        __syncthreads();
        const uint64_t lo = max(my_o, tw << 5), hi = min(my_e, (tw + T_WORDS) << 5);
        if (lo < hi) {
            for (uint64_t w = lo >> 5; (w << 5) < hi; w++) {
                const uint32_t i = (uint32_t)(w - tw);
                const uint64_t c = w << 5;
                uint32_t m = sV[i];
                if (c < lo) m &= ~0u << (lo - c);
                if (c + 32 > hi) m &= ~0u >> (c + 32 - hi);
                const uint32_t H = sH[i], L = sL[i];
                cnt[0] += __popc(~H & ~L & m);
                cnt[1] += __popc(~H & L & m);
                cnt[2] += __popc(H & ~L & m);
                cnt[3] += __popc(H & L & m);
            }
        }
        __syncthreads();
    }
    int cls = -1;
    if (r < n_reads) {
        const long long len = (long long)(my_e - my_o);
        if (len < fp.min_len) cls = 1;                              // filter_reads.cpp:189
        else {
            cnt[4] = (unsigned int)len - (cnt[0] + cnt[1] + cnt[2] + cnt[3]);
            cls = classify_counts(len, cnt, fp);
            if (cls == 4) {
                const unsigned int slot = atomicAdd(n_border, 1u);
                if (slot < border_cap) {
                    BorderRec br;
                    br.read = r;
                    for (int q = 0; q < 5; q++) br.cnt[q] = cnt[q];
                    br.len = (unsigned int)len;
                    border[slot] = br;
                }
                cls = 0;    // provisional; the host patches classes/bits/totals
            }
        }
        if (classes) classes[r] = (uint8_t)cls;
    }
    const uint32_t sel = __ballot_sync(0xffffffffu, cls == 0);
    if ((tid & 31) == 0 && (r >> 5) < n_bv_words) bv[r >> 5] = sel;   // padding bits stay 0
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t mc = __ballot_sync(0xffffffffu, cls == (c == 3 ? 0 : c + 1));
        if ((tid & 31) == 0 && mc) atomicAdd(&tot[c], __popc(mc));
    }
    __syncthreads();
    // (two blocks share the totals of one k_filter block of 1024 reads: the host zeroes them before the launch)
    if (tid < 4 && tot[tid]) atomicAdd(&block_totals[4 * (r_first / kFilterBlock) + tid], tot[tid]);
}

// apply host decisions for the undecided reads: newcls[i] for border[i].read
__global__ void k_filter_patch(const BorderRec *__restrict__ border, const uint8_t *__restrict__ newcls,
                               unsigned int n, uint32_t *__restrict__ bv, uint8_t *__restrict__ classes,
                               unsigned int *__restrict__ block_totals)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (newcls[i] == 0) return;                      // stays selected
    uint64_t r = border[i].read;
    atomicAnd(&bv[r >> 5], ~(1u << (r & 31)));
    if (classes) classes[r] = newcls[i];
    uint64_t blk = r / kFilterBlock;
    atomicSub(&block_totals[4 * blk + 3], 1u);       // selected--
    atomicAdd(&block_totals[4 * blk + (newcls[i] - 1)], 1u);
}

// The -m cutoff (filter_reads.cpp:188,203-205): the loop stops once
// max_reads reads are selected; counters only cover reads before the stop and
// every later bit is cleared.  Single block: locate the stop from the block
// totals, then the exact read inside that block from the class bytes.
// out[0..3] = rm_len, rm_N, rm_shannon, selected ; out[4] = cutoff position.
__global__ void __launch_bounds__(1024)
k_filter_cutoff(const unsigned int *__restrict__ block_totals, uint64_t n_blocks,
                const uint8_t *__restrict__ classes, uint64_t n_reads, long long max_reads,
                unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long part[1024][4];
    __shared__ unsigned long long base[4];
    __shared__ unsigned long long stop_block;
    const unsigned int tid = threadIdx.x;
    uint64_t per = (n_blocks + 1023) / 1024;
    uint64_t lo = tid * per, hi = lo + per < n_blocks ? lo + per : n_blocks;
    unsigned long long s[4] = {0, 0, 0, 0};
    for (uint64_t b = lo; b < hi; b++)
        for (int c = 0; c < 4; c++) s[c] += block_totals[4 * b + c];
    for (int c = 0; c < 4; c++) part[tid][c] = s[c];
    if (tid == 0) stop_block = n_blocks;
    __syncthreads();
    if (tid == 0) {
        // serial scan over 1024 partials, then over the owning thread's range
        unsigned long long acc[4] = {0, 0, 0, 0};
        unsigned int owner = 1024;
        for (unsigned int i = 0; i < 1024; i++) {
            if (max_reads >= 0 && acc[3] + part[i][3] >= (unsigned long long)max_reads) { owner = i; break; }
            for (int c = 0; c < 4; c++) acc[c] += part[i][c];
        }
        if (owner < 1024) {
            uint64_t b = owner * per, e = b + per < n_blocks ? b + per : n_blocks;
            for (; b < e; b++) {
                if (acc[3] + block_totals[4 * b + 3] >= (unsigned long long)max_reads) break;
                for (int c = 0; c < 4; c++) acc[c] += block_totals[4 * b + c];
            }
            stop_block = b;
        }
        for (int c = 0; c < 4; c++) base[c] = acc[c];
    }
    __syncthreads();
    if (stop_block >= n_blocks) {            // never reached: counters are the grand totals
        if (tid == 0) {
            for (int c = 0; c < 4; c++) out[c] = base[c];
            out[4] = n_reads;
        }
        return;
    }
    // inside the stop block: inclusive scan of selected flags over its 1024 reads
    __shared__ unsigned int scan[1024];
    uint64_t r = stop_block * kFilterBlock + tid;
    int cls = (r < n_reads) ? classes[r] : -1;
    scan[tid] = (cls == 0);
    __syncthreads();
    for (unsigned int d = 1; d < 1024; d <<= 1) {
        unsigned int v = (tid >= d) ? scan[tid - d] : 0;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
    }
    unsigned long long need = (unsigned long long)max_reads - base[3];   // >= 1 selected reads from this block
    __shared__ unsigned int cut;     // index within block of the read that reaches max_reads
    if (tid == 0) cut = 1024;
    __syncthreads();
    if (max_reads == 0) { if (tid == 0) cut = 0; }
    else if (cls == 0 && scan[tid] == need) cut = tid;
    __syncthreads();
    // reads [0, cut] of the block are processed (cut itself is the last selected one);
    // with max_reads == 0 nothing is processed at all.
    unsigned int last = (max_reads == 0) ? 0 : cut + 1;     // number of processed reads in block
    __shared__ unsigned int cnt[4];
    if (tid < 4) cnt[tid] = 0;
    __syncthreads();
    if (tid < last && cls >= 0) atomicAdd(&cnt[cls == 0 ? 3 : cls - 1], 1u);
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < 4; c++) out[c] = base[c] + cnt[c];
        out[4] = stop_block * kFilterBlock + last;
    }
}

// clear bits [cutoff, n) (untag_last_reads, read_file.h:76-81)
__global__ void __launch_bounds__(256)
k_clear_from(uint32_t *__restrict__ bv, const unsigned long long *__restrict__ cutoff_p, uint64_t n_words)
{
    uint64_t cutoff = *cutoff_p;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (cutoff >> 5) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        if ((w << 5) >= cutoff) bv[w] = 0;
        else bv[w] &= ~(~0u << (cutoff - (w << 5)));
    }
}

// --------------------------------------------------------- stage 4: bvop ----
// BooleanVector::full_and/or/and_not/not (boolean_vector.h:418-462): 16-byte
// vectors grid-stride, byte tail by the last threads.
template <int OP>
__device__ __forceinline__ uint32_t bv_apply(uint32_t a, uint32_t b)
{
    if (OP == 0) return a & b;
    if (OP == 1) return a | b;
    if (OP == 2) return a & ~b;
    return ~a;
}

template <int OP>
__global__ void __launch_bounds__(256)
k_bvop(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint4 *__restrict__ out,
       uint64_t n_vec, uint64_t n_bytes)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = i0; i < n_vec; i += stride) {
        uint4 x = ld_nc_u4(a + i);
        uint4 y = (OP == 3) ? x : ld_nc_u4(b + i);
        uint4 r;
        r.x = bv_apply<OP>(x.x, y.x); r.y = bv_apply<OP>(x.y, y.y);
        r.z = bv_apply<OP>(x.z, y.z); r.w = bv_apply<OP>(x.w, y.w);
        out[i] = r;
    }
    const uint8_t *a8 = reinterpret_cast<const uint8_t *>(a);
    const uint8_t *b8 = reinterpret_cast<const uint8_t *>(b);
    uint8_t *o8 = reinterpret_cast<uint8_t *>(out);
    for (uint64_t j = n_vec * 16 + i0; j < n_bytes; j += stride)
        o8[j] = (uint8_t)bv_apply<OP>(a8[j], (OP == 3) ? 0u : b8[j]);
}

// nb_one (boolean_vector.h:244-270): popcount of all n_bytes (clamp on host)
__global__ void __launch_bounds__(256)
k_popcount(const uint4 *__restrict__ a, uint64_t n_vec, uint64_t n_bytes,
           unsigned long long *__restrict__ total)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    for (uint64_t i = i0; i < n_vec; i += stride) {
        uint4 x = ld_nc_u4(a + i);
        local += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
    }
    const uint8_t *a8 = reinterpret_cast<const uint8_t *>(a);
    for (uint64_t j = n_vec * 16 + i0; j < n_bytes; j += stride) local += __popc((uint32_t)a8[j]);
    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < 8; i++) s += ws[i];
        if (s) atomicAdd(total, s);
    }
}

// filter |= other (multi-GPU merge of partial filters; `other` may be peer memory)
__global__ void __launch_bounds__(256)
k_or_into(uint4 *__restrict__ dst, const uint4 *__restrict__ src, uint64_t n_vec)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        uint4 s = ld_nc_u4(src + i);
        if ((s.x | s.y | s.z | s.w) == 0) continue;      // sparse partials: skip the write
        uint4 d = dst[i];
        d.x |= s.x; d.y |= s.y; d.z |= s.z; d.w |= s.w;
        dst[i] = d;
    }
}

// ----------------------------------------- multi-GPU: one-kernel OR all-reduce ----
// Every rank holds a partial filter (its shard of the chunk's reads).  Rank `me`
// owns vectors [v0, v1) of the filter: it pulls that slice from every peer
// through NVLink (P2P loads on IPC-mapped peer memory), ORs the partials with its
// own, and pushes the merged slice back into EVERY rank's filter (P2P stores).
// When all ranks have run this kernel every filter is the OR of all partials:
// reduce-scatter + all-gather in one pass, (G-1)/G of the filter in each
// direction per GPU, instead of NCCL all-gather (G-1 filters in) + local OR.
constexpr int kMaxPeers = 8;
struct PeerFilters { uint4 *f[kMaxPeers]; };

// Peer accesses are plain (weak) 16-byte LDG/STG: the partials were completed before the kernel started and
// the pushes are consumed after it ends (the caller brackets the launch with rank barriers), so no
// system-scope ordering is needed inside the kernel -- .sys-scoped accesses cost NVLink round trips.
__device__ __forceinline__ uint4 ld_peer_u4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer_u4(uint4 *p, uint4 v)
{
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int G>
__global__ void __launch_bounds__(256)
k_merge_peers(PeerFilters pf, int me, uint64_t v0, uint64_t v1)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += stride) {
        uint4 part[G];
#pragma unroll
        for (int p = 0; p < G; p++) part[p] = (p == me) ? pf.f[p][i] : ld_peer_u4(pf.f[p] + i);   // G loads in flight
        uint4 acc = part[0];
#pragma unroll
        for (int p = 1; p < G; p++) { acc.x |= part[p].x; acc.y |= part[p].y; acc.z |= part[p].z; acc.w |= part[p].w; }
#pragma unroll
        for (int p = 0; p < G; p++) {
            if (p == me) pf.f[p][i] = acc;
            else st_peer_u4(pf.f[p] + i, acc);
        }
    }
}

// ------------------------------- multi-GPU: owner-applied insert + slice all-gather ----
// Merging whole partial filters moves 2 (G-1)/G F bytes per GPU and direction (reduce-scatter + all-gather of the OR).
// The reduce-scatter half is avoidable: what a rank contributes to a part of the filter is not a dense slice but
// the RECORDS of its reads that fall into it -- (G-1)/G of 16 bytes per k-mer instead of (G-1)/G F.  So the filter's
// 32 MiB regions are dealt to the ranks; every rank scatters the records of ITS reads into its own slabs, as on one
// GPU; then every owner applies the records of ITS regions from ALL ranks' slabs -- the record tiles of the peers are
// read straight out of their memory over NVLink by the apply kernel itself (no copy pass, the link transfer overlaps
// the RED.OR) -- and sweeps only its own regions through L2; finally every rank pulls the finished regions of the
// others (k_gather_regions).  The regions are dealt by their record counts (largest first to the least loaded
// rank, computed identically by every rank from the exchanged counters): key d = a|b piles 13 % of its records
// into the all-ones region, and real reads are worse, so equal SHARES of the regions are not equal shares of the work.
struct PeerInsert {
    const uint32_t *recs[kMaxPeers];      // slab pools
    const uint32_t *fill[kMaxPeers];      // records per region
    const uint32_t *table[kMaxPeers];     // slab tables
    uint32_t max_q[kMaxPeers];            // row length of each table
};
struct OwnerTile {
    const uint32_t *src;                  // the tile's records (peer or local memory)
    uint32_t n;                           // records in the tile; bit 31: the records are in peer memory
    uint32_t bin;                         // region (global index)
    uint32_t rt, rn;                      // tile index inside the region, tiles of the region (all sources)
    uint32_t next_fill;                   // records of the next owned region (decides whether it is prefetched)
    uint32_t next_bin;                    // the next owned region (~0: none)
};

// one thread per tile: where its records are (the slab id is looked up in the source's table, over NVLink for a peer).
// pairs e = i * world + s (i-th owned region own_bins[i], source rank s), region-major; fills[e] = records of the pair,
// tbase[e] = its first tile, tbase[n_pairs] = number of tiles -- computed on the host from the ranks' counters.
template <int TILE>
__global__ void __launch_bounds__(256)
k_owner_tiles(PeerInsert pi, int world, int me, const uint32_t *__restrict__ own_bins, int n_pairs, const uint32_t *__restrict__ fills,
              const uint32_t *__restrict__ tbase, OwnerTile *__restrict__ tiles)
{
    const uint32_t n_tiles = tbase[n_pairs];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        int lo = 0, hi = n_pairs - 1;                      // last pair with tbase <= t (pairs without tiles are skipped)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tbase[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int e = lo, i = e / world, s = e % world;
        const uint32_t bin = own_bins[i];
        const uint32_t lt = t - tbase[e], v0 = lt * TILE;
        const uint32_t slab = pi.table[s][(size_t)bin * pi.max_q[s] + (v0 >> kSlabLog2)] - 1u;
        OwnerTile o;
        o.src = pi.recs[s] + (((size_t)slab << kSlabLog2) + (v0 & (kSlabRecs - 1)));
        o.n = min((uint32_t)TILE, fills[e] - v0) | (s != me ? 0x80000000u : 0u);
        o.bin = bin;
        o.rt = t - tbase[i * world];
        o.rn = tbase[(i + 1) * world] - tbase[i * world];
        uint32_t nf = 0;
        o.next_bin = 0xFFFFFFFFu;
        if ((i + 1) * world < n_pairs) {
            for (int q = 0; q < world; q++) nf += fills[(i + 1) * world + q];
            o.next_bin = own_bins[i + 1];
        }
        o.next_fill = nf;
        tiles[t] = o;
    }
}

template <int TILE, bool PREFETCH>
__global__ void __launch_bounds__(256)
k_owner_apply(uint32_t *__restrict__ filter, const OwnerTile *__restrict__ tiles, const uint32_t *__restrict__ tbase, int n_pairs,
              unsigned long long *__restrict__ tile_counter)
{
    constexpr int U = TILE / (256 * 4);
    __shared__ unsigned long long s_next;
    const uint32_t n_tiles = tbase[n_pairs];
    const uint64_t pol = ld_policy_evict_first();
    if (threadIdx.x == 0) s_next = atomicAdd(tile_counter, 1ull);
    __syncthreads();
    unsigned long long tl = s_next;
    while (tl < n_tiles) {
        __syncthreads();                                   // everybody holds tl: the slot may be overwritten
        unsigned long long nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(tile_counter, 1ull);      // consumed after this tile
        const OwnerTile o = tiles[tl];
        const uint32_t n_here = o.n & 0x7FFFFFFFu;
        const bool remote = (o.n >> 31) != 0;
        const uint4 *src = reinterpret_cast<const uint4 *>(o.src);
        uint4 v[U];
#pragma unroll
        for (int it = 0; it < U; it++) {
            const uint32_t e = (it * 256 + threadIdx.x) * 4;
            v[it] = make_uint4(0u, 0u, 0u, 0u);
            if (e < n_here) v[it] = remote ? ld_peer_u4(src + (it * 256 + threadIdx.x)) : ld_stream_u4(src + (it * 256 + threadIdx.x), pol);
        }
        if (PREFETCH && o.next_bin != 0xFFFFFFFFu && o.next_fill >= (1u << (kRegionLog2 - 8))) {
            const uint32_t lines = 1u << (kRegionLog2 - 7);
            const uint32_t l0 = (uint32_t)((uint64_t)lines * o.rt / o.rn), l1 = (uint32_t)((uint64_t)lines * (o.rt + 1) / o.rn);
            const char *nxt_region = reinterpret_cast<const char *>(filter) + ((uint64_t)o.next_bin << kRegionLog2);
            for (uint32_t l = l0 + threadIdx.x; l < l1; l += 256)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(nxt_region + ((uint64_t)l << 7)));
        }
        uint32_t *region = filter + ((uint64_t)o.bin << (kRegionLog2 - 2));
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
        if (threadIdx.x == 0) s_next = nxt;
        __syncthreads();
        tl = s_next;
    }
}

// every rank pulls the finished regions of the other owners into its own filter.  list[j] = region | owner << 16; the
// work is cut into pieces of kGatherPiece vectors dealt round-robin over the list, so that the blocks running at any
// time pull from different peers
constexpr uint32_t kGatherPiece = 2048;                    // 16-byte vectors per piece (32 KiB)
__global__ void __launch_bounds__(256)
k_gather_regions(PeerFilters pf, int me, const uint32_t *__restrict__ list, uint32_t n_list, uint32_t region_vecs)
{
    const uint32_t pieces_per_region = region_vecs / kGatherPiece;
    const uint64_t n_pieces = (uint64_t)pieces_per_region * n_list;
    uint4 *mine = pf.f[me];
    for (uint64_t pc = blockIdx.x; pc < n_pieces; pc += gridDim.x) {
        const uint32_t j = (uint32_t)(pc % n_list), part = (uint32_t)(pc / n_list);
        const uint32_t ent = list[j];
        const uint64_t v0 = (uint64_t)(ent & 0xFFFFu) * region_vecs + (uint64_t)part * kGatherPiece;
        const uint4 *src = pf.f[ent >> 16] + v0;
        uint4 *dst = mine + v0;
        uint4 part_v[kGatherPiece / 256];
#pragma unroll
        for (int u = 0; u < (int)(kGatherPiece / 256); u++) part_v[u] = ld_peer_u4(src + u * 256 + threadIdx.x);      // 8 peer loads in flight
#pragma unroll
        for (int u = 0; u < (int)(kGatherPiece / 256); u++) dst[u * 256 + threadIdx.x] = part_v[u];
    }
}

// ------------------------------------------------- measurement kernels ----
// random 32-byte-sector ceilings: every lane touches an independent random
// sector (one u32 load, or one RED.OR) of a `n_words`-word buffer.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <bool ATOMIC>
__global__ void __launch_bounds__(256)
k_random_sectors(uint32_t *__restrict__ buf, uint64_t n_words_mask, uint64_t n_ops,
                 unsigned long long *__restrict__ sink)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_ops; i += 4 * stride) {
        uint64_t x0 = splitmix64(i) & n_words_mask, x1 = splitmix64(i + stride) & n_words_mask;
        uint64_t x2 = splitmix64(i + 2 * stride) & n_words_mask, x3 = splitmix64(i + 3 * stride) & n_words_mask;
        if (ATOMIC) {
            atomicOr(buf + x0, 1u << (x0 & 31)); atomicOr(buf + x1, 1u << (x1 & 31));
            atomicOr(buf + x2, 1u << (x2 & 31)); atomicOr(buf + x3, 1u << (x3 & 31));
        } else {
            uint32_t v0 = ld_nc_u32(buf + x0), v1 = ld_nc_u32(buf + x1);
            uint32_t v2 = ld_nc_u32(buf + x2), v3 = ld_nc_u32(buf + x3);
            acc += v0 + v1 + v2 + v3;
        }
    }
    for (; i < n_ops; i += stride) {
        uint64_t x = splitmix64(i) & n_words_mask;
        if (ATOMIC) atomicOr(buf + x, 1u << (x & 31));
        else acc += ld_nc_u32(buf + x);
    }
    if (!ATOMIC && acc == 0x12345678u) atomicAdd(sink, 1ull);
}

}  // namespace commet
