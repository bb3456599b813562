// measurement kernels: random 32-byte-sector loads / RED.OR (the ceilings bench.py reports against)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"

namespace commet {

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
