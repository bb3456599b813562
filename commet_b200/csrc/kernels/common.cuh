// loads, the four keys of a k-mer as plane windows, their place in the filter (hash_key.h:65-125, bloom_filter.h:112-131)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
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

}  // namespace commet
