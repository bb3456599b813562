// stage 1: the insert of a range of reads into the filter (direct, L2-blocked), filter transfer, merge of partial filters
// (part of the C-ABI library: included by capi.cu, in this order, into one translation unit)

// ---------------------------------------------------------- stage 1: index --
extern "C" int commet_index_begin(commet_ctx *c, int k)
{
    CKR(set_device(c));
    if (k < 1 || k > kMaxK) return fail("k=%d unsupported (1..%d)", k, kMaxK);
    uint64_t bytes = commet_filter_bytes(k);
    // at least one whole 2 MiB block of its own: smaller cudaMalloc allocations are sub-allocated by the
    // driver, and a CUDA IPC handle (commet_index_export) always maps the enclosing block
    const uint64_t blk = 2ull << 20;
    uint64_t cap = std::max<uint64_t>((bytes + blk - 1) & ~(blk - 1), blk);
    if (c->filter_cap < cap) {
        if (c->filter) { cudaFree(c->filter); c->filter = nullptr; c->filter_cap = 0; }
        cudaError_t e = cudaMalloc(&c->filter, cap);
        if (e != cudaSuccess)
            return fail("Index memory allocation impossible (%llu bytes for k=%d): %s",
                        (unsigned long long)cap, k, cudaGetErrorString(e));
        c->filter_cap = cap;
    }
    c->filter_bytes = bytes;
    c->k = k;
    CK(cudaMemsetAsync(c->filter, 0, std::max<uint64_t>((bytes + 255) & ~255ull, 256), c->stream));
    return 0;
}

// L2-blocked insert of stream positions [b0, b1): see kernels.cuh.  kmers_hint = upper bound of the
// k-mers in the range (0: unknown -> the number of positions).  Returns 1 if the direct path must be used.
static int index_range_binned(commet_ctx *c, commet_reads *r, uint64_t b0, uint64_t b1, uint64_t kmers_hint)
{
    const int k = c->k;
    const int n_bins = 1 << (k - kRecKeyBits);
    if (!c->bins) CK(cudaMalloc(&c->bins, 2048 * sizeof(unsigned long long)));
    if (!(c->s2_attr & 1u)) {                        // per device, once
        c->s2_attr |= 1u;
        CK(cudaFuncSetAttribute(k_bin_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
        CK(cudaFuncSetAttribute(k_bin_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 128 + 128) * 4));
    }
    unsigned long long *hist = c->bins, *base = c->bins + 512, *cursor = c->bins + 1100, *tile_counter = c->bins + 1700;
    uint64_t positions = b1 - b0;
    uint64_t kmers = kmers_hint ? std::min(kmers_hint, positions) : positions;
    // scratch: 4 records of 4 bytes per k-mer; bounded by what the device has free, else sub-ranges.  The
    // driver is only asked for the free memory when the buffer has to grow: cudaMemGetInfo takes device-wide
    // locks and was measured to block the host for tens of milliseconds between two launches.
    uint64_t need = 4 * kmers + 64;
    uint64_t parts = 1;
    uint64_t budget = c->recs_cap;
    if (const char *e = getenv("COMMET_B200_RECS_BUDGET")) {                   // tests: force sub-ranges
        uint64_t v = strtoull(e, nullptr, 10);
        if (v >= 4096) budget = v;
        else if (need > budget) budget = 0;
    } else if (need > budget) budget = 0;
    if (budget == 0) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        budget = (uint64_t)((free_b + c->recs_cap * 4) * 0.6) / 4;             // records
    }
    if (need > budget) {
        need = 4 * positions + 64;                   // sub-ranges are cut by position: no per-part k-mer count
        parts = (need + budget - 1) / budget;
        need = 4 * ((positions + parts - 1) / parts + 32) + 64;
    }
    if (c->recs_cap < need) {
        if (c->recs) { cudaFree(c->recs); c->recs = nullptr; c->recs_cap = 0; }
        if (cudaMalloc(&c->recs, need * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return 1;                                // no room for the record buffer: direct atomics
        }
        c->recs_cap = need;
    }
    for (uint64_t p = 0; p < parts; p++) {
        uint64_t s0 = b0 + positions * p / parts, s1 = b0 + positions * (p + 1) / parts;
        if (s1 <= s0) continue;
        CK(cudaMemsetAsync(hist, 0, 512 * sizeof(unsigned long long), c->stream));
        unsigned g = grid_for(c, s1 - s0 + 32, 256, 8);
        if (n_bins <= 128) {     // pair table: n_bins^2 + n_bins counters of dynamic shared memory
            const size_t sh = ((size_t)n_bins * n_bins + n_bins) * sizeof(unsigned int);
            k_bin_count<true><<<std::min(g, (unsigned)c->sm_count * env_or("COMMET_B200_COUNT_BPS", 3)), 256, sh, c->stream>>>(r->planes, s0, s1, k, n_bins, hist);
        } else
            k_bin_count<false><<<g, 256, n_bins * sizeof(unsigned int), c->stream>>>(r->planes, s0, s1, k, n_bins, hist);
        k_bin_scan<<<1, 32, 0, c->stream>>>(hist, n_bins, base, cursor, tile_counter);
        uint64_t n_tiles = (((s1 + 31) >> 5) - (s0 >> 5) + kScatTileWords - 1) / kScatTileWords;
        unsigned gs = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)c->sm_count * env_or("COMMET_B200_SCATTER_BPS", 2));
        k_bin_scatter<<<gs, kScatThreads, sizeof(ScatterSmem), c->stream>>>(r->planes, s0, s1, k, n_bins, cursor, c->recs);
        {
            int tile = 2048, bps = 8, pf = 1;
            if (const char *e = getenv("COMMET_B200_APPLY_TILE")) tile = atoi(e);
            if (const char *e = getenv("COMMET_B200_APPLY_BPS")) bps = atoi(e);
            if (const char *e = getenv("COMMET_B200_APPLY_PREFETCH")) pf = atoi(e);
            const unsigned ga = c->sm_count * bps;
            if (tile == 8192 && pf) k_bin_apply<8192, true><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, base, n_bins, tile_counter);
            else if (tile == 8192) k_bin_apply<8192, false><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, base, n_bins, tile_counter);
            else if (tile == 2048 && pf) k_bin_apply<2048, true><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, base, n_bins, tile_counter);
            else if (tile == 2048) k_bin_apply<2048, false><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, base, n_bins, tile_counter);
            else if (pf) k_bin_apply<4096, true><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, base, n_bins, tile_counter);
            else k_bin_apply<4096, false><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, base, n_bins, tile_counter);
        }
        c->launches += 4;
        CK(cudaGetLastError());
    }
    return 0;
}

// The second form of the L2-blocked insert (kernels.cuh: k_bin_scatter2 / k_bin_plan2 / k_bin_apply2): no histogram
// pass, records in slabs.  Same contract as index_range_binned.
template <int TW>
static void launch_scatter2(commet_ctx *c, commet_reads *r, uint64_t s0, uint64_t s1, int k, int n_bins, uint32_t *fill,
                            uint32_t max_q, uint32_t *n_slabs, unsigned bps)
{
    const size_t sh = scatter2_smem_bytes(TW, n_bins);
    if (!(c->s2_attr & (unsigned)TW)) {             // per device: the attribute belongs to the context's device
        cudaFuncSetAttribute(k_bin_scatter2<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem_bytes(TW, kMaxBins));
        c->s2_attr |= (unsigned)TW;
    }
    const uint64_t n_tiles = (((s1 + 31) >> 5) - (s0 >> 5) + TW - 1) / TW;
    const unsigned g = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)c->sm_count * bps);
    const uint32_t max_slabs = (uint32_t)std::min<uint64_t>(c->recs_cap >> kSlabLog2, 0xFFFFFFFFull);
    k_bin_scatter2<TW><<<g, kS2Threads, sh, c->stream>>>(r->planes, s0, s1, k, n_bins, fill, c->slab_table, max_q, n_slabs, max_slabs, c->recs);
}

// record pool, slab table and counters of the second form for one launch of up to `bound` records; returns the row
// length of the table in *max_q.  The counters and the table get whole 2 MiB blocks of their own: CUDA IPC maps the
// block an allocation lies in (commet_dist_open exports them to the other ranks).  1: no room (direct atomics).
static int ensure_insert_buffers(commet_ctx *c, int n_bins, uint64_t bound, uint32_t *max_q)
{
    if (!c->bins) CK(cudaMalloc(&c->bins, 2048 * sizeof(unsigned long long)));
    if (!c->bins2) CK(cudaMalloc(&c->bins2, 2u << 20));
    const uint64_t need = ((bound + kSlabRecs - 1) / kSlabRecs + (uint64_t)n_bins + 1) * kSlabRecs;
    if (c->recs_cap < need) {
        if (c->recs) { cudaFree(c->recs); c->recs = nullptr; c->recs_cap = 0; }
        if (cudaMalloc(&c->recs, need * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        c->recs_cap = need;
    }
    *max_q = (uint32_t)((bound + kSlabRecs - 1) / kSlabRecs + 1);
    const uint64_t table_entries = std::max<uint64_t>((uint64_t)n_bins * *max_q, (2u << 20) / sizeof(uint32_t));
    if (c->slab_table_cap < table_entries) {
        if (c->slab_table) { cudaFree(c->slab_table); c->slab_table = nullptr; c->slab_table_cap = 0; }
        CK(cudaMalloc(&c->slab_table, table_entries * sizeof(uint32_t)));
        c->slab_table_cap = table_entries;
    }
    return 0;
}

// records of stream positions [s0, s1) -> the context's slabs (fill[], table rows of max_q entries)
static int scatter_range(commet_ctx *c, commet_reads *r, uint64_t s0, uint64_t s1, int n_bins, uint32_t max_q)
{
    uint32_t *fill = c->bins2, *n_slabs = c->bins2 + 1030;
    const int tw = (int)env_or("COMMET_B200_S2_TW", 96);
    const unsigned sbps = env_or("COMMET_B200_SCATTER_BPS", tw <= 96 ? 3 : 2);
    CK(cudaMemsetAsync(c->bins2, 0, 2048 * sizeof(uint32_t), c->stream));
    CK(cudaMemsetAsync(c->slab_table, 0, (size_t)n_bins * max_q * sizeof(uint32_t), c->stream));
    if (s1 <= s0) return 0;                          // nothing to scatter: zeroed counters, no launch
    if (tw <= 64) launch_scatter2<64>(c, r, s0, s1, c->k, n_bins, fill, max_q, n_slabs, sbps);
    else if (tw <= 96) launch_scatter2<96>(c, r, s0, s1, c->k, n_bins, fill, max_q, n_slabs, sbps);
    else launch_scatter2<128>(c, r, s0, s1, c->k, n_bins, fill, max_q, n_slabs, sbps);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int TILE, int STAGES>
static void launch_apply3(commet_ctx *c, int pf, unsigned grid, uint32_t *fill, uint32_t *tbase, uint32_t max_q, int n_bins,
                          unsigned long long *tile_counter)
{
    const size_t sh = (size_t)TILE * 4 * STAGES;
    const unsigned bit = 0x100000u << (TILE / 4096);
    if (!(c->s2_attr & bit)) {
        cudaFuncSetAttribute(k_bin_apply3<TILE, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        cudaFuncSetAttribute(k_bin_apply3<TILE, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        c->s2_attr |= bit;
    }
    if (pf) k_bin_apply3<TILE, STAGES, true><<<grid, 256, sh, c->stream>>>(c->filter, c->recs, fill, tbase, c->slab_table, max_q, n_bins, tile_counter);
    else k_bin_apply3<TILE, STAGES, false><<<grid, 256, sh, c->stream>>>(c->filter, c->recs, fill, tbase, c->slab_table, max_q, n_bins, tile_counter);
}

static int index_range_binned2(commet_ctx *c, commet_reads *r, uint64_t b0, uint64_t b1, uint64_t kmers_hint)
{
    const int k = c->k;
    const int n_bins = 1 << (k - kRecKeyBits);
    const uint64_t positions = b1 - b0;
    const uint64_t kmers = kmers_hint ? std::min(kmers_hint, positions) : positions;
    // records of one launch: bounded by the 32-bit record counters and by what the device has room for (in slabs)
    const uint64_t limit = 0xE0000000ull;
    auto slabs_for = [&](uint64_t recs) { return (recs + kSlabRecs - 1) / kSlabRecs + (uint64_t)n_bins + 1; };
    uint64_t bound = 4 * kmers, parts = 1;
    uint64_t budget = c->recs_cap > ((uint64_t)n_bins + 1) * kSlabRecs ? c->recs_cap - ((uint64_t)n_bins + 1) * kSlabRecs : 0;   // records
    bool forced = false;
    if (const char *e = getenv("COMMET_B200_RECS_BUDGET")) {                   // tests: force sub-ranges
        uint64_t v = strtoull(e, nullptr, 10);
        if (v >= 4096) { budget = v; forced = true; }
    }
    if (!forced && slabs_for(bound) * kSlabRecs > c->recs_cap) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t room = (uint64_t)((free_b + c->recs_cap * 4) * 0.6) / 4;       // records
        budget = room > ((uint64_t)n_bins + 1) * kSlabRecs ? room - ((uint64_t)n_bins + 1) * kSlabRecs : 0;
        if (budget < kSlabRecs) return 1;                                            // no room: direct atomics
    }
    budget = std::min(budget, limit);
    if (bound > budget) {
        parts = (4 * positions + budget - 1) / budget;            // sub-ranges are cut by position: no per-part k-mer count
        bound = 4 * ((positions + parts - 1) / parts + 32);
    }
    uint32_t max_q = 0;
    {
        const int rc = ensure_insert_buffers(c, n_bins, bound, &max_q);
        if (rc != 0) return rc;
    }
    uint32_t *fill = c->bins2, *tbase = c->bins2 + 512;
    unsigned long long *tile_counter = c->bins + 1700;
    // COMMET_B200_APPLY_FORM=3 (A/B): record tiles through the bulk-copy engine (k_bin_apply3); measured equal to the
    // LDG form (both sit at the L2 lookup rate), which stays the default
    const int aform = (int)env_or("COMMET_B200_APPLY_FORM", 2);
    int tile = 2048, bps = aform == 3 ? 6 : 8, pf = 1;
    if (const char *e = getenv("COMMET_B200_APPLY_TILE")) tile = atoi(e);
    if (const char *e = getenv("COMMET_B200_APPLY_BPS")) bps = atoi(e);
    if (const char *e = getenv("COMMET_B200_APPLY_PREFETCH")) pf = atoi(e);
    for (uint64_t p = 0; p < parts; p++) {
        const uint64_t s0 = b0 + positions * p / parts, s1 = b0 + positions * (p + 1) / parts;
        if (s1 <= s0) continue;
        CKR(scatter_range(c, r, s0, s1, n_bins, max_q));
        const unsigned ga = c->sm_count * bps;
        if (aform == 3) {
            // record tiles through the bulk-copy engine into a ring of shared-memory stages
            if (tile == 4096) {
                k_bin_plan2<4096><<<1, 32, 0, c->stream>>>(fill, n_bins, tbase, tile_counter);
                launch_apply3<4096, 3>(c, pf, ga, fill, tbase, max_q, n_bins, tile_counter);
            } else {
                k_bin_plan2<2048><<<1, 32, 0, c->stream>>>(fill, n_bins, tbase, tile_counter);
                launch_apply3<2048, 4>(c, pf, ga, fill, tbase, max_q, n_bins, tile_counter);
            }
        } else if (tile == 4096) {
            k_bin_plan2<4096><<<1, 32, 0, c->stream>>>(fill, n_bins, tbase, tile_counter);
            if (pf) k_bin_apply2<4096, true><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, fill, tbase, c->slab_table, max_q, n_bins, tile_counter);
            else k_bin_apply2<4096, false><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, fill, tbase, c->slab_table, max_q, n_bins, tile_counter);
        } else {
            k_bin_plan2<2048><<<1, 32, 0, c->stream>>>(fill, n_bins, tbase, tile_counter);
            if (pf) k_bin_apply2<2048, true><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, fill, tbase, c->slab_table, max_q, n_bins, tile_counter);
            else k_bin_apply2<2048, false><<<ga, 256, 0, c->stream>>>(c->filter, c->recs, fill, tbase, c->slab_table, max_q, n_bins, tile_counter);
        }
        c->launches += 2;
        CK(cudaGetLastError());
    }
    return 0;
}

static int index_range(commet_ctx *c, commet_reads *r, uint64_t first, uint64_t count, uint64_t kmers_hint)
{
    if (c->k == 0) return fail("commet_index_add before commet_index_begin");
    if (first + count > r->n_reads) return fail("index range out of bounds");
    if (count == 0 || r->n_words == 0) return 0;
    CKR(prepare(c, r, c->k));
    uint64_t hb[2] = {0, r->n_bases};
    if (first != 0 || count != r->n_reads) {
        // a sub-range: its two stream offsets are read back (16 bytes) so the work can be sized
        CK(cudaMemcpyAsync(&hb[0], r->offs + first, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(&hb[1], r->offs + first + count, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if (hb[1] <= hb[0]) return 0;
    // filters larger than L2 (k >= 28: > 64 MiB): region passes (default) or the sort-based L2-blocked path
    if (c->binned_index && c->region_passes && c->k >= 28 && c->k - 1 - c->region_log2 >= 1 && c->k - 1 - c->region_log2 <= 10) {
        const int R = c->k - 1 - c->region_log2;            // region = 2^region_log2 bytes = top R key bits
        k_index_regions<<<c->sm_count * 8, 256, 0, c->stream>>>(c->filter, r->planes, hb[0], hb[1], c->k, R);
        c->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    if (c->binned_index && c->k >= 28 && c->k - kRecKeyBits <= 9) {
        int form = c->insert_form;
        if (const char *e = getenv("COMMET_B200_INSERT")) form = atoi(e);          // A/B (scripts/ab_index.py)
        int rc = form == 2 ? index_range_binned2(c, r, hb[0], hb[1], kmers_hint)
                                     : index_range_binned(c, r, hb[0], hb[1], kmers_hint);
        if (rc <= 0) return rc;
    }
    uint64_t positions = hb[1] - hb[0] + 32;
    k_index<<<grid_for(c, positions, 256, 8), 256, 0, c->stream>>>(c->filter, r->planes, hb[0], hb[1], c->k, nullptr);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int commet_index_add(commet_ctx *c, commet_reads *r, uint64_t first, uint64_t count)
{
    CKR(set_device(c));
    return index_range(c, r, first, count, 0);
}

extern "C" void *commet_index_filter_ptr(commet_ctx *c) { return c->filter; }

extern "C" int commet_index_download(commet_ctx *c, uint8_t *out, uint64_t bytes)
{
    CKR(set_device(c));
    if (bytes > c->filter_bytes) return fail("filter is %llu bytes", (unsigned long long)c->filter_bytes);
    CK(cudaMemcpyAsync(out, c->filter, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int commet_index_upload(commet_ctx *c, int k, const uint8_t *filter, uint64_t bytes)
{
    CKR(commet_index_begin(c, k));
    if (bytes != c->filter_bytes) return fail("filter for k=%d must be %llu bytes", k, (unsigned long long)c->filter_bytes);
    CK(cudaMemcpyAsync(c->filter, filter, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int commet_index_or(commet_ctx *c, const void *d_other, uint64_t offset, uint64_t bytes)
{
    CKR(set_device(c));
    if ((offset & 15) || offset + bytes > c->filter_cap) return fail("commet_index_or: bad range");
    uint64_t n_vec = (bytes + 15) / 16;
    if (n_vec == 0) return 0;
    k_or_into<<<grid_for(c, n_vec, 256, 8), 256, 0, c->stream>>>(
        reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(c->filter) + offset),
        reinterpret_cast<const uint4 *>(d_other), n_vec);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------- multi-GPU merge ----
extern "C" int commet_index_export(commet_ctx *c, uint8_t handle[COMMET_IPC_HANDLE_BYTES])
{
    CKR(set_device(c));
    if (!c->filter) return fail("commet_index_export before commet_index_begin");
    static_assert(sizeof(cudaIpcMemHandle_t) == COMMET_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->filter));
    memcpy(handle, &h, sizeof h);
    return 0;
}

extern "C" int commet_peer_open(commet_ctx *c, const uint8_t handle[COMMET_IPC_HANDLE_BYTES], void **d_filter)
{
    CKR(set_device(c));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CK(cudaIpcOpenMemHandle(d_filter, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int commet_peer_close(commet_ctx *c, void *d_filter)
{
    CKR(set_device(c));
    if (d_filter) CK(cudaIpcCloseMemHandle(d_filter));
    return 0;
}

extern "C" int commet_index_merge(commet_ctx *c, void *const *d_filters, int n_ranks, int rank)
{
    CKR(set_device(c));
    if (n_ranks < 1 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks)
        return fail("commet_index_merge: %d ranks (rank %d) unsupported (1..%d)", n_ranks, rank, kMaxPeers);
    if (!c->filter) return fail("commet_index_merge before commet_index_begin");
    if (n_ranks == 1) return 0;
    PeerFilters pf;
    for (int p = 0; p < kMaxPeers; p++) pf.f[p] = nullptr;
    for (int p = 0; p < n_ranks; p++) {
        pf.f[p] = p == rank ? reinterpret_cast<uint4 *>(c->filter) : static_cast<uint4 *>(d_filters[p]);
        if (!pf.f[p]) return fail("commet_index_merge: no filter mapped for rank %d", p);
    }
    uint64_t n_vec = std::max<uint64_t>(c->filter_bytes / 16, 1);      // filter_cap >= 256 bytes
    uint64_t v0 = n_vec * rank / n_ranks, v1 = n_vec * (rank + 1) / n_ranks;
    if (v1 <= v0) return 0;
    unsigned g = grid_for(c, v1 - v0, 256, 8);
    switch (n_ranks) {
    case 2: k_merge_peers<2><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    case 3: k_merge_peers<3><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    case 4: k_merge_peers<4><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    case 5: k_merge_peers<5><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    case 6: k_merge_peers<6><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    case 7: k_merge_peers<7><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    default: k_merge_peers<8><<<g, 256, 0, c->stream>>>(pf, rank, v0, v1); break;
    }
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}
