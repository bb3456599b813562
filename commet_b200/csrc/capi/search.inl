// stage 2: the search launches and the chunk loop (commet_index_and_search and its staged / resident forms)
// (part of the C-ABI library: included by capi.cu, in this order, into one translation unit)

// --------------------------------------------------------- stage 2: search --
static int search_launch(commet_ctx *c, commet_reads *r, int k, int t, uint32_t *d_tags, unsigned long long *d_counters)
{
    if (c->k != k || !c->filter) return fail("commet_search: no filter for k=%d (current k=%d)", k, c->k);
    if (r->n_reads == 0) return 0;
    CKR(prepare(c, r, k));
    // One read per thread (up to 512 blocks per SM): the cost of a read varies from a dozen probes (a copy, found
    // at once) to 2(L-k+1) (no k-mer in common), and a grid-stride loop over a grid of 8 blocks per SM left the
    // SMs unevenly loaded (measured: 27.4 ms with 1184 blocks, 22.8 ms with 9472, same kernel).
    unsigned bps = 512;
    if (const char *e = getenv("COMMET_B200_SEARCH_BPS")) bps = (unsigned)atoi(e);
    unsigned g = grid_for(c, r->n_reads, 256, bps);
#define COMMET_SEARCH(COUNT, BOTH) \
    k_search<COUNT, BOTH><<<g, 256, 0, c->stream>>>(c->filter, r->planes, r->offs, r->n_reads, k, t, d_tags, d_counters, r->sel)
    if (!c->count_probes && c->search_dynamic) {
        // persistent warps, reads handed out from a cursor (scratch[170]); search_dynamic = resident blocks per SM
        unsigned long long *cursor = c->scratch + 170;
        CK(cudaMemsetAsync(cursor, 0, sizeof *cursor, c->stream));
        const unsigned gd = (unsigned)std::min<uint64_t>((r->n_reads + 255) / 256, (uint64_t)c->sm_count * (unsigned)c->search_dynamic);
        if (c->search_dynamic >= 4)      // 64 registers (a few spilled), 4 resident blocks per SM
            k_search_dyn<4, 4><<<gd, 256, 0, c->stream>>>(c->filter, r->planes, r->offs, r->n_reads, k, t, d_tags, d_counters, r->sel, cursor);
        else                             // 73 registers, 3 resident blocks per SM
            k_search_dyn<4, 3><<<gd, 256, 0, c->stream>>>(c->filter, r->planes, r->offs, r->n_reads, k, t, d_tags, d_counters, r->sel, cursor);
    } else if (c->count_probes) COMMET_SEARCH(true, 0);
    else if (env_or("COMMET_B200_SEARCH_VARIANT", 0) == 44)      // A/B: round 1's shape -- 4 positions per strand and batch, 4 blocks per SM
        COMMET_SEARCH(false, 4);
    else if (c->search_both == 4 && k <= 30 && env_or("COMMET_B200_SEARCH_VARIANT", 0) != 25)
        // keys of at most 30 bits (filters up to 512 MiB, the L2-resident ones among them): 32-bit windows and keys, 40
        // registers, 6 resident blocks per SM -- the scan is latency-bound there and lives on resident warps
        // (profiles/r02_search_occupancy_ab.txt: 290 ms against 366 ms at k=27 = 71 % of the L2 random-sector ceiling)
        k_search<false, 2, 6, true><<<g, 256, 0, c->stream>>>(c->filter, r->planes, r->offs, r->n_reads, k, t, d_tags, d_counters, r->sel);
    else if (c->search_both == 4)
        // 2 positions per strand and batch (4 a-probes in flight per lane), 5 resident blocks per SM at 48 registers:
        // 18.6 against 19.2 ms at k=33
        k_search<false, 2, 5><<<g, 256, 0, c->stream>>>(c->filter, r->planes, r->offs, r->n_reads, k, t, d_tags, d_counters, r->sel);
    else if (c->search_both == 2) COMMET_SEARCH(false, 2);
    else if (c->search_both == 8) COMMET_SEARCH(false, 8);
    else if (c->search_both) COMMET_SEARCH(false, 4);
    else COMMET_SEARCH(false, 0);
#undef COMMET_SEARCH
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int commet_search_dev(commet_ctx *c, commet_reads *r, int k, int t, uint32_t *d_tags, uint64_t *d_counters)
{
    CKR(set_device(c));
    CK(cudaMemsetAsync(d_counters + 1, 0, sizeof(uint64_t), c->stream));
    return search_launch(c, r, k, t, d_tags, reinterpret_cast<unsigned long long *>(d_counters));
}

extern "C" int commet_search(commet_ctx *c, commet_reads *r, int k, int t, uint8_t *tags, uint64_t *n_found,
                             uint64_t *n_searched)
{
    CKR(set_device(c));
    uint64_t nb = r->n_reads / 8 + 1, nw = tag_words(r->n_reads);
    DevBuf d(c);
    if (d.alloc(nw * 4) != cudaSuccess) return fail("tag allocation failed");
    CK(cudaMemsetAsync(d.p, 0, nw * 4, c->stream));
    CK(cudaMemcpyAsync(d.p, tags, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->scratch, 0, 4 * sizeof(unsigned long long), c->stream));
    CKR(search_launch(c, r, k, t, d.as<uint32_t>(), c->scratch));
    unsigned long long cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(tags, d.p, nb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt, c->scratch, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (n_found) *n_found = cnt[0];
    if (n_searched) *n_searched = cnt[1];
    return 0;
}

// ------------------------------------------------------------- chunk loop ---
namespace {

// CUDA-event stopwatch over segments of the compute stream (index / search device time of the log lines)
struct SegTimer {
    std::vector<cudaEvent_t> ev;
    bool on = true;
    int begin(cudaStream_t st)
    {
        if (!on) return 0;
        if (ev.size() >= 512) { on = false; return 0; }
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        ev.push_back(e0);
        ev.push_back(e1);
        CK(cudaEventRecord(e0, st));
        return 0;
    }
    int end(cudaStream_t st)
    {
        if (!on || ev.empty()) return 0;
        CK(cudaEventRecord(ev.back(), st));
        return 0;
    }
    double total_ms()          // after a stream sync
    {
        double t = 0;
        for (size_t i = 0; on && i + 1 < ev.size(); i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) t += ms;
        }
        return t;
    }
    ~SegTimer() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};

}  // namespace

// per-read k-mer counts of a staged stream (device) and their sum (host; syncs the compute stream)
static int count_kmers(commet_ctx *c, commet_reads *r, int k, DevBuf &counts, unsigned long long *total)
{
    *total = 0;
    CKR(prepare(c, r, k));
    if (r->n_reads == 0) return 0;
    if (counts.alloc(r->n_reads * sizeof(uint32_t)) != cudaSuccess) return fail("allocation of k-mer counts failed");
    CK(cudaMemsetAsync(c->scratch + 150, 0, sizeof(unsigned long long), c->stream));
    k_kmer_counts<<<grid_for(c, r->n_reads, 256, 8), 256, 0, c->stream>>>(r->planes, r->offs, r->n_reads,
                                                                         counts.as<uint32_t>(), c->scratch + 150);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(total, c->scratch + 150, sizeof *total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// src/index_and_search.cpp:255-277 as ONE streaming pass over the index set, given as consecutive parts of
// its valid-read stream (one part for device-resident sets; a few for host sets, so that part i+1 crosses
// PCIe while part i is inserted).  The stop rule of index_reads (index_reads.h:48-49,60) is applied on the
// running k-mer count: as long as a whole part stays below max_kmer it is inserted without looking at
// per-read counts; only a part that contains a chunk boundary has its counts walked on the host.  A chunk
// closes after the read that reaches max_kmer, every query set is searched against it, and the next read is
// fetched-and-lost -- also when that read is the first one of the next part.
static int chunk_loop(commet_ctx *c, int k, int t, uint64_t max_kmer, const std::vector<commet_reads *> &parts,
                      int n_sets, commet_reads *const *queries, uint32_t *const *d_tags,
                      uint64_t *searched, uint64_t *shared, uint64_t *stats)
{
    // cnt[4s..4s+3]: found total, searched in the last chunk, filter tests, k-mer lookups -- sized from n_sets (the
    // reference takes any number of search sets, and Commet.py puts all the other samples into one -s file)
    DevBuf cnt_buf(c);
    const size_t n_cnt = 4 * (size_t)std::max(n_sets, 1);
    if (cnt_buf.alloc(n_cnt * sizeof(unsigned long long)) != cudaSuccess) return fail("counter allocation failed");
    unsigned long long *d_cnt = cnt_buf.as<unsigned long long>();
    CK(cudaMemsetAsync(d_cnt, 0, n_cnt * sizeof(unsigned long long), c->stream));
    uint64_t n_chunks = 0, n_indexed = 0, n_kmers = 0;
    uint64_t cum = 0, open_reads = 0;
    bool began = false, dirty = false, pending_drop = false;
    SegTimer t_index, t_search;
    const uint64_t clear_bytes = std::max<uint64_t>((commet_filter_bytes(k) + 255) & ~255ull, 256);

    auto open_filter = [&]() -> int {
        if (!began) { CKR(commet_index_begin(c, k)); began = true; }
        else if (dirty) CK(cudaMemsetAsync(c->filter, 0, clear_bytes, c->stream));
        dirty = false;
        return 0;
    };
    auto insert = [&](commet_reads *r, uint64_t first, uint64_t count, uint64_t kmers, uint64_t n_sel) -> int {
        CKR(open_filter());
        CKR(t_index.begin(c->stream));
        CKR(index_range(c, r, first, count, kmers));
        CKR(t_index.end(c->stream));
        n_indexed += n_sel;
        n_kmers += kmers;
        open_reads += n_sel;
        return 0;
    };
    auto close_chunk = [&]() -> int {
        CKR(open_filter());                 // a chunk without reads still owns an (empty) filter
        // (a query stream that is still crossing PCIe is encoded by search_launch right before its own scan: the scans of
        // the streams that have arrived do not wait for it)
        CKR(t_search.begin(c->stream));
        for (int s = 0; s < n_sets; s++) {
            CK(cudaMemsetAsync(d_cnt + 4 * s + 1, 0, sizeof(unsigned long long), c->stream));
            CKR(search_launch(c, queries[s], k, t, d_tags[s], d_cnt + 4 * s));
        }
        CKR(t_search.end(c->stream));
        n_chunks++;
        cum = 0;
        open_reads = 0;
        dirty = true;
        return 0;
    };

    // "reads" below are the SELECTED reads of a part (commet_reads_select); unselected ones carry no k-mer
    // (their W bits are cleared) and are invisible to the stop rule, exactly like reads the reference's
    // get_next_read skips (fasta_file.h:143-152)
    for (commet_reads *r : parts) {
        const uint64_t n = r->n_reads;
        if (n == 0 || sel_count(r, 0, n) == 0) continue;
        DevBuf counts(c);
        unsigned long long total = 0;
        CKR(count_kmers(c, r, k, counts, &total));
        trace("part: encode + W plane + k-mer counts queued, total read back (sync)");
        std::vector<uint32_t> cnt;          // fetched only when a chunk boundary falls inside this part
        uint64_t first = 0, rem = total;
        if (pending_drop) {                 // the read fetched and lost by the previous chunk (index_reads.h:60)
            while (first < n && !sel_get(r, first)) first++;
            uint32_t c0 = 0;
            CK(cudaMemcpyAsync(&c0, counts.as<uint32_t>() + first, sizeof c0, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            rem -= c0;
            first++;
            pending_drop = false;
        }
        while (first < n) {
            if (cum + rem < max_kmer) {     // the rest of the part fits in the open chunk
                CKR(insert(r, first, n - first, rem, sel_count(r, first, n)));
                trace("part: insert queued");
                cum += rem;
                break;
            }
            if (cnt.empty()) {
                cnt.resize(n);
                CK(cudaMemcpyAsync(cnt.data(), counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
                CK(cudaStreamSynchronize(c->stream));
            }
            uint64_t i = first, fed = 0, taken = 0;
            while (i < n && cum < max_kmer) {
                if (sel_get(r, i)) { cum += cnt[i]; fed += cnt[i]; taken++; }
                i++;
            }
            if (i > first) CKR(insert(r, first, i - first, fed, taken));
            rem -= fed;
            CKR(close_chunk());             // cum >= max_kmer here, because cum + rem was
            while (i < n && !sel_get(r, i)) i++;
            if (i < n) { rem -= cnt[i]; i++; } else pending_drop = true;
            first = i;
        }
    }
    if (open_reads > 0) CKR(close_chunk());
    trace("last chunk: searches queued");

    std::vector<unsigned long long> h(n_cnt);
    CK(cudaMemcpyAsync(h.data(), d_cnt, n_cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    trace("counters read back (sync)");
    uint64_t n_tests = 0, n_lookups = 0;
    for (int s = 0; s < n_sets; s++) {
        if (shared) shared[s] = h[4 * s];
        if (searched) searched[s] = h[4 * s + 1];
        n_tests += h[4 * s + 2];
        n_lookups += h[4 * s + 3];
    }
    if (stats) {
        stats[0] = n_chunks; stats[1] = n_indexed; stats[2] = n_kmers;
        stats[3] = (uint64_t)(t_index.total_ms() * 1e6); stats[4] = (uint64_t)(t_search.total_ms() * 1e6);
        stats[5] = n_tests; stats[6] = n_lookups; stats[7] = parts.size();
    }
    return 0;
}

extern "C" int commet_index_and_search_staged(commet_ctx *c, int k, int t, uint64_t max_kmer, commet_reads *index,
                                              int n_sets, commet_reads *const *queries, uint32_t *const *d_tags,
                                              uint64_t *searched, uint64_t *shared, uint64_t *stats)
{
    CKR(set_device(c));
    if (n_sets < 0) return fail("n_sets=%d unsupported", n_sets);
    if (k < 1 || k > kMaxK) return fail("k=%d unsupported (1..%d)", k, kMaxK);
    std::vector<commet_reads *> parts(1, index);
    return chunk_loop(c, k, t, max_kmer, parts, n_sets, queries, d_tags, searched, shared, stats);
}

// The same loop on resident streams with HOST outputs: what a persistent driver calls once per
// index_and_search round of Commet.py:186-240 (commet_b200/csrc/tools/commet_nxn.cpp).  Tag words live in the
// context's arena for the duration of the call; ones[s] is the device-side popcount of set s's tag vector
// (k_popcount), i.e. the number `bvop -i` would print for the .bv files of that set (Commet.py:252-271).
extern "C" int commet_index_and_search_resident(commet_ctx *c, int k, int t, uint64_t max_kmer, commet_reads *index,
                                                int n_sets, commet_reads *const *queries, uint8_t *const *tags,
                                                uint64_t *searched, uint64_t *shared, uint64_t *ones, uint64_t *stats)
{
    CKR(set_device(c));
    if (n_sets < 0) return fail("n_sets=%d unsupported", n_sets);
    if (k < 1 || k > kMaxK) return fail("k=%d unsupported (1..%d)", k, kMaxK);
    std::vector<uint32_t *> dt(n_sets, nullptr);
    int rc = 0;
    for (int s = 0; rc == 0 && s < n_sets; s++) {
        const uint64_t nw = tag_words(queries[s]->n_reads);
        if (c->arena.alloc((void **)&dt[s], nw * 4) != cudaSuccess) rc = fail("tag allocation failed");
        else if (cudaMemsetAsync(dt[s], 0, nw * 4, c->stream) != cudaSuccess) rc = fail("tag memset failed");
    }
    std::vector<commet_reads *> parts(1, index);
    if (rc == 0) rc = chunk_loop(c, k, t, max_kmer, parts, n_sets, queries, dt.data(), searched, shared, stats);
    if (rc == 0 && ones && n_sets > 0) {
        std::vector<const void *> pv(dt.begin(), dt.end());
        std::vector<uint64_t> nb(n_sets);
        for (int s = 0; s < n_sets; s++) nb[s] = queries[s]->n_reads;
        rc = commet_bv_popcount_batch_dev(c, pv.data(), nb.data(), n_sets, ones);
    }
    for (int s = 0; rc == 0 && s < n_sets; s++) {
        if (rc == 0 && tags && tags[s] &&
            cudaMemcpyAsync(tags[s], dt[s], queries[s]->n_reads / 8 + 1, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            rc = fail("tag download failed");
    }
    if (rc == 0 && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail("stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (int s = 0; s < n_sets; s++) if (dt[s]) c->arena.free(dt[s]);
    return rc;
}

// Read ranges of the parts a host-resident index set is uploaded in: 20 % / 30 % / 50 % of the bases, cut at
// read boundaries.  Growing parts keep the copy of part i+1 shorter than the insert of part i, so only the
// first (small) part's copy is exposed; few parts keep the number of sweeps of the filter low.
static std::vector<uint64_t> split_parts(const uint64_t *offs, uint64_t n_reads)
{
    std::vector<uint64_t> cuts(1, 0);
    const uint64_t n_bases = offs[n_reads];
    uint64_t min_part = 64ull << 20;
    if (const char *e = getenv("COMMET_B200_PART_BYTES")) min_part = std::max<uint64_t>(strtoull(e, nullptr, 10), 1);   // tests
    if (n_bases >= 4 * min_part) {
        std::vector<double> frac = {0.2, 0.5};
        if (const char *e = getenv("COMMET_B200_PART_FRACS")) {          // tuning: increasing cut positions in (0,1), comma separated
            frac.clear();
            for (const char *q = e; *q;) {
                char *end = nullptr;
                double f = strtod(q, &end);
                if (end == q) break;
                if (f > 0.0 && f < 1.0) frac.push_back(f);
                q = *end ? end + 1 : end;
            }
        }
        for (double f : frac) {
            uint64_t target = (uint64_t)(f * (double)n_bases);
            uint64_t r = (uint64_t)(std::lower_bound(offs, offs + n_reads + 1, target) - offs);
            if (r > cuts.back() && r < n_reads) cuts.push_back(r);
        }
    }
    cuts.push_back(n_reads);
    return cuts;
}

extern "C" int commet_index_and_search(commet_ctx *c, int k, int t, uint64_t max_kmer, const uint8_t *ibases,
                                       const uint64_t *ioffs, uint64_t n_index, int n_sets,
                                       const uint8_t *const *qbases, const uint64_t *const *qoffs,
                                       const uint64_t *n_query, uint8_t *const *tags, uint64_t *searched,
                                       uint64_t *shared, uint64_t *stats)
{
    CKR(set_device(c));
    if (n_sets < 0) return fail("n_sets=%d unsupported", n_sets);
    if (k < 1 || k > kMaxK) return fail("k=%d unsupported (1..%d)", k, kMaxK);
    if (ioffs[0] != 0) return fail("commet_index_and_search: ioffs[0] must be 0");
    std::vector<commet_reads *> parts;
    std::vector<uint32_t *> dt(n_sets, nullptr);
    // a large query set is uploaded (and searched) in a few parts cut at multiples of 32 reads -- their tag words are
    // disjoint ranges of the set's vector -- so that the search of part i runs while part i+1 still crosses PCIe
    std::vector<commet_reads *> vq;          // the parts of all sets, set after set
    std::vector<uint32_t *> vtags;
    std::vector<int> v_set;
    // every H2D copy is queued up front on the copy stream (index parts first); the host never waits for one
    std::vector<uint64_t> cuts = split_parts(ioffs, n_index);
    int rc = 0;
    HostTrace tr;
    g_trace = tr.on ? &tr : nullptr;
    trace("enter");
    for (size_t p = 0; rc == 0 && p + 1 < cuts.size(); p++) {
        commet_reads *r = nullptr;
        rc = reads_upload_async(c, ibases + ioffs[cuts[p]], ioffs + cuts[p], cuts[p + 1] - cuts[p], &r);
        if (rc == 0) parts.push_back(r);
        trace("index part: allocations + copies queued");
    }
    uint64_t q_part_bytes = 256ull << 20;
    if (const char *e = getenv("COMMET_B200_QUERY_PART_BYTES")) q_part_bytes = std::max<uint64_t>(strtoull(e, nullptr, 10), 1);     // tests
    const uint64_t max_q_parts = std::max(1u, env_or("COMMET_B200_QUERY_PARTS", 4));
    for (int s = 0; rc == 0 && s < n_sets; s++) {
        if (qoffs[s][0] != 0) { rc = fail("commet_index_and_search: qoffs[%d][0] must be 0", s); break; }
        const uint64_t nw = tag_words(n_query[s]);
        if (c->arena.alloc((void **)&dt[s], nw * 4) != cudaSuccess) { rc = fail("tag allocation failed"); break; }
        if (cudaMemsetAsync(dt[s], 0, nw * 4, c->stream) != cudaSuccess) { rc = fail("tag memset failed"); break; }
        const uint64_t n = n_query[s], bytes = qoffs[s][n];
        const uint64_t n_parts = std::max<uint64_t>(1, std::min<uint64_t>(max_q_parts, bytes / q_part_bytes));
        uint64_t a = 0;
        for (uint64_t p = 0; rc == 0 && p < n_parts; p++) {
            uint64_t b = n;
            if (p + 1 < n_parts) {
                const uint64_t target = bytes * (p + 1) / n_parts;
                b = (uint64_t)(std::lower_bound(qoffs[s], qoffs[s] + n + 1, target) - qoffs[s]) & ~31ull;
                b = std::min(std::max(b, a), n);
            }
            if (b == a && p + 1 < n_parts) continue;
            commet_reads *r = nullptr;
            rc = reads_upload_async(c, qbases[s] + qoffs[s][a], qoffs[s] + a, b - a, &r);
            if (rc == 0) {
                vq.push_back(r);
                vtags.push_back(dt[s] + a / 32);
                v_set.push_back(s);
            }
            a = b;
        }
    }
    trace("query sets: allocations + copies queued");
    const int nv = (int)vq.size();
    std::vector<uint64_t> v_searched(std::max(nv, 1), 0), v_shared(std::max(nv, 1), 0);
    if (rc == 0) rc = chunk_loop(c, k, t, max_kmer, parts, nv, vq.data(), vtags.data(), v_searched.data(), v_shared.data(), stats);
    if (rc == 0) {
        for (int s = 0; s < n_sets; s++) {
            if (searched) searched[s] = 0;
            if (shared) shared[s] = 0;
        }
        for (int v = 0; v < nv; v++) {
            if (searched) searched[v_set[v]] += v_searched[v];
            if (shared) shared[v_set[v]] += v_shared[v];
        }
    }
    for (int s = 0; rc == 0 && s < n_sets; s++)
        if (cudaMemcpyAsync(tags[s], dt[s], n_query[s] / 8 + 1, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            rc = fail("tag download failed");
    if (rc == 0 && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail("stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (commet_reads *r : parts) commet_reads_free(r);
    trace("tags downloaded (sync)");
    for (commet_reads *r : vq) commet_reads_free(r);
    for (int s = 0; s < n_sets; s++) if (dt[s]) c->arena.free(dt[s]);
    trace("freed");
    g_trace = nullptr;
    return rc;
}
