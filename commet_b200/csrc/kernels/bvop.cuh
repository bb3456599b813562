// stage 4: boolean-vector operators and nb_one (boolean_vector.h:244-270,418-462)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"

namespace commet {

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

}  // namespace commet
