// stage 3: filter_reads -- length, N, Shannon, -m cutoff (filter_reads.cpp:181-205,265-306)
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"
#include "staging.cuh"

namespace commet {

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
// The same fusion with k_encode's regularity.  A block takes THREADS consecutive reads (kFilterBlock / THREADS blocks
// share one unit of k_filter's outputs) and sweeps the WORDS of their span of the stream, a tile of sf2_tile_words() at a
// time: thread t encodes words t, t + THREADS, ... -- every word once, perfectly balanced, coalesced 32-byte loads -- into shared memory (and, with PLANES, into
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

}  // namespace commet
