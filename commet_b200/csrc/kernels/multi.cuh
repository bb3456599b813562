// multi-GPU: merge of partial filters, owner-applied insert, region all-gather, over NVLink peer memory
// (part of the device code of commet_b200; kernels.cuh includes every part, capi.cu launches them)
#pragma once
#include "common.cuh"
#include "insert.cuh"

namespace commet {

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

}  // namespace commet
