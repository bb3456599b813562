// Random-sector microbenchmark: which load flavour / access shape gets the most random filter probes per second
// out of a DRAM-resident buffer, and how many DRAM bytes does each probe cost (run under ncu for the bytes).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_sectors scripts/ubench_sectors.cu
//   ./ubench_sectors [gran]      gran = cudaLimitMaxL2FetchGranularity set before the first allocation (0: leave)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int V>
__device__ __forceinline__ uint32_t ld_variant(const uint32_t *p, uint64_t pol)
{
    uint32_t v;
    if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 1) asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 2) asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 5) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 6) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 7) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    else if (V == 8) asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (V == 9) asm volatile("ld.global.L2::256B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// MODE 0: every lane an independent random word.  MODE 1: lanes pair up on the two halves of one 64-byte block.
// MODE 2: 4 lanes share a 128-byte line (one sector each).  MODE 3: 32 lanes read one 128-byte line (4 B each).
template <int V, int MODE>
__global__ void __launch_bounds__(256) k_gather(const uint32_t *__restrict__ buf, uint64_t mask, uint64_t n_ops,
                                                unsigned long long *sink)
{
    uint64_t pol = 0;
    if (V == 7) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n_ops; i += 4 * stride) {
        uint64_t x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint64_t j = i + u * stride;
            if (MODE == 0) x[u] = splitmix64(j) & mask;
            else if (MODE == 1) x[u] = ((splitmix64(j >> 1) & mask) & ~15ull) | ((j & 1) << 3);
            else if (MODE == 2) x[u] = ((splitmix64(j >> 2) & mask) & ~31ull) | ((j & 3) << 3);
            else if (MODE == 3) x[u] = ((splitmix64(j >> 5) & mask) & ~31ull) | (j & 31);
            else {
                // MODE 4/5/6/7: groups of 4 lanes read 4 random lines of one 1 / 2 / 4 / 16 KiB block (DRAM page locality)
                const uint64_t blk = MODE == 4 ? 256 : MODE == 5 ? 512 : MODE == 6 ? 1024 : 4096;      // words
                uint64_t h = splitmix64(j >> 2);
                uint64_t in = splitmix64(j * 0x51ED27ull + 77) & (blk - 1);
                x[u] = ((h & mask) & ~(blk - 1)) | in;
            }
        }
        uint32_t v0 = ld_variant<V>(buf + x[0], pol), v1 = ld_variant<V>(buf + x[1], pol);
        uint32_t v2 = ld_variant<V>(buf + x[2], pol), v3 = ld_variant<V>(buf + x[3], pol);
        acc += v0 + v1 + v2 + v3;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// how many independent random loads per thread and how many blocks does it take to reach the DRAM random-access rate?
template <int ILP>
__global__ void __launch_bounds__(256) k_gather_ilp(const uint32_t *__restrict__ buf, uint64_t mask, uint64_t n_ops,
                                                    unsigned long long *sink)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i + (ILP - 1) * stride < n_ops; i += ILP * stride) {
        uint32_t v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) v[u] = ld_variant<2>(buf + (splitmix64(i + u * stride) & mask), 0);
#pragma unroll
        for (int u = 0; u < ILP; u++) acc += v[u];
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int ILP>
static void run_ilp(const uint32_t *buf, uint64_t bytes, uint64_t n_ops, unsigned blocks, unsigned long long *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    uint64_t mask = bytes / 4 - 1;
    k_gather_ilp<ILP><<<blocks, 256>>>(buf, mask, n_ops / 8, sink);
    cudaEventRecord(e0);
    k_gather_ilp<ILP><<<blocks, 256>>>(buf, mask, n_ops, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"variant\": \"L2::64B, %d loads in flight per thread, %u blocks\", \"MiB\": %llu, \"ms\": %.3f, \"Gops_s\": %.2f}\n", ILP,
           blocks, (unsigned long long)(bytes >> 20), ms, n_ops / (ms * 1e6));
    fflush(stdout);
}

template <int V, int MODE>
static void run(const char *name, const uint32_t *buf, uint64_t bytes, uint64_t n_ops, unsigned long long *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    uint64_t mask = bytes / 4 - 1;
    k_gather<V, MODE><<<148 * 8, 256>>>(buf, mask, n_ops / 8, sink);      // warm-up
    cudaEventRecord(e0);
    k_gather<V, MODE><<<148 * 8, 256>>>(buf, mask, n_ops, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    printf("{\"variant\": \"%s\", \"mode\": %d, \"MiB\": %llu, \"ms\": %.3f, \"Gops_s\": %.2f%s}\n", name, MODE,
           (unsigned long long)(bytes >> 20), ms, n_ops / (ms * 1e6), e == cudaSuccess ? "" : ", \"error\": true");
    fflush(stdout);
}

int main(int argc, char **argv)
{
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"set_l2_fetch\": %d, \"rc\": %d, \"now\": %zu}\n", gran, (int)e, got);
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"default_l2_fetch\": %zu}\n", got);
    }
    uint32_t *buf = nullptr;
    unsigned long long *sink = nullptr;
    const uint64_t cap = 16ull << 30;
    if (cudaMalloc(&buf, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, cap);
    cudaMemset(sink, 0, 8);
    const uint64_t n = 1ull << 30;
    const uint64_t G4 = 4ull << 30;
    run<0, 0>("nc.no_allocate", buf, G4, n, sink);
    run<1, 0>("plain", buf, G4, n, sink);
    run<2, 0>("L2::64B", buf, G4, n, sink);
    run<3, 0>("nc.no_allocate.L2::64B", buf, G4, n, sink);
    run<4, 0>("cv", buf, G4, n, sink);
    run<5, 0>("cg", buf, G4, n, sink);
    run<6, 0>("relaxed.gpu", buf, G4, n, sink);
    run<7, 0>("nc.evict_first", buf, G4, n, sink);
    run<8, 0>("L2::128B", buf, G4, n, sink);
    run<9, 0>("L2::256B", buf, G4, n, sink);
    run<10, 0>("cs", buf, G4, n, sink);
    run<0, 1>("nc.no_allocate pairs/64B", buf, G4, n, sink);
    run<0, 2>("nc.no_allocate 4 sectors/line", buf, G4, n, sink);
    run<0, 3>("nc.no_allocate 32 lanes/line", buf, G4, n, sink);
    run<0, 4>("4 lines of one 1 KiB block", buf, G4, n, sink);
    run<0, 5>("4 lines of one 2 KiB block", buf, G4, n, sink);
    run<0, 6>("4 lines of one 4 KiB block", buf, G4, n, sink);
    run<0, 7>("4 lines of one 16 KiB block", buf, G4, n, sink);
    for (unsigned blocks : {148u * 4, 148u * 8, 148u * 16, 148u * 64}) {
        run_ilp<1>(buf, G4, n, blocks, sink);
        run_ilp<2>(buf, G4, n, blocks, sink);
        run_ilp<4>(buf, G4, n, blocks, sink);
        run_ilp<8>(buf, G4, n, blocks, sink);
        run_ilp<16>(buf, G4, n, blocks, sink);
    }
    run<0, 0>("nc.no_allocate", buf, 256ull << 20, n, sink);
    run<0, 0>("nc.no_allocate", buf, 1ull << 30, n, sink);
    run<0, 0>("nc.no_allocate", buf, 16ull << 30, n, sink);
    return 0;
}
