// ASCII -> H/L/V bit-planes, read starts, the W plane, per-read k-mer counts (alphabet.h:44-58, index_reads.h:52-58)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"

namespace commet {

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

}  // namespace commet
