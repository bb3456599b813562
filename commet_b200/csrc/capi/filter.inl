// stage 3: filter_reads (selection on planes, fused staging + selection, host decisions for undecided reads)
// (part of the C-ABI library: included by capi.cu, in this order, into one translation unit)

// --------------------------------------------------- stage 3: filter_reads --
// exact shannon_index (filter_reads.cpp:265-306) from the device's counts, for
// the few reads whose device value lies within the log-implementation margin
// of the threshold: glibc's double log is what the reference calls.
static float shannon_from_counts(const unsigned int cnt[5], unsigned int len)
{
    float index = 0;
    for (int j = 0; j < 5; j++) {
        float f = (float)cnt[j] / (float)len;
        if (f != 0) index += (double)f * ::log((double)f) / ::log(2.0);
    }
    return fabsf(index);
}

// Shared tail of the two filter kernels: `launch` runs k_filter (bit-planes) or k_stage_filter (fused with the
// staging pass); then the undecided reads are settled, the -m cutoff located and the counters fetched.
template <class Launch>
static int filter_run(commet_ctx *c, uint64_t n, int64_t min_len, int64_t max_N, float min_shannon, int64_t max_reads,
                      uint32_t *d_bv, uint64_t *counters, Launch launch)
{
    uint64_t n_bv_words = tag_words(n);
    uint64_t n_blocks = std::max<uint64_t>((std::max(n, n_bv_words * 32) + kFilterBlock - 1) / kFilterBlock, 1);
    if (n_blocks > 0x7fffffffull) return fail("too many reads for one filter call");
    FilterParams fp;
    fp.min_len = min_len;
    fp.max_N = max_N == -1 ? 2147483647LL : max_N;      // -1: no limit; any other negative value drops every read (filter_reads.cpp:192)
    fp.min_shannon = min_shannon;
    fp.margin = 2e-5f;
    if (max_reads < -1) max_reads = 0;          // `selected < max_reads` is false at once: nothing kept
    const bool cut = max_reads >= 0 && (uint64_t)max_reads < n;
    DevBuf totals(c), classes(c), nb(c), patch(c);
    if (totals.alloc(n_blocks * 4 * sizeof(unsigned int)) != cudaSuccess || nb.alloc(sizeof(unsigned int)) != cudaSuccess)
        return fail("filter scratch allocation failed");
    // class bytes are needed to locate a -m cutoff and to patch undecided reads' totals
    if (classes.alloc(n ? n : 1) != cudaSuccess) return fail("filter class allocation failed");
    // Undecided reads (device value within `margin` of the threshold) come back as records of exact counts.  The
    // buffer starts at 2^20 records; a set with more of them -- dinucleotide repeats have H = 1.0 exactly, so `-e 1` on
    // a low-complexity-rich set makes every such read undecided -- is run again with a buffer of the size it asked for.
    unsigned int border_cap = 1u << 20, n_border = 0;
    if (const char *e = getenv("COMMET_B200_BORDER_CAP")) border_cap = std::max(1, atoi(e));        // tests
    std::vector<BorderRec> recs;
    for (;;) {
        DevBuf border(c);
        if (border.alloc((size_t)border_cap * sizeof(BorderRec)) != cudaSuccess) return fail("filter scratch allocation failed");
        CK(cudaMemsetAsync(nb.p, 0, sizeof(unsigned int), c->stream));
        CK(cudaMemsetAsync(totals.p, 0, n_blocks * 4 * sizeof(unsigned int), c->stream));
        launch((unsigned)n_blocks, fp, n_bv_words, classes.as<uint8_t>(), totals.as<unsigned int>(), border.as<BorderRec>(),
               border_cap, nb.as<unsigned int>());
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&n_border, nb.p, sizeof n_border, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (n_border > border_cap) { border_cap = n_border; continue; }
        if (n_border) {
            recs.resize(n_border);
            CK(cudaMemcpyAsync(recs.data(), border.p, (size_t)n_border * sizeof(BorderRec), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        break;
    }
    if (n_border) {
        // the decision depends on the five counts only: reads at an exact threshold share a handful of count tuples
        std::vector<uint8_t> cls(n_border);
        std::map<std::array<unsigned int, 5>, uint8_t> memo;
        for (unsigned int i = 0; i < n_border; i++) {
            const std::array<unsigned int, 5> key = {recs[i].cnt[0], recs[i].cnt[1], recs[i].cnt[2], recs[i].cnt[3], recs[i].cnt[4]};
            auto it = memo.find(key);
            if (it == memo.end())
                it = memo.emplace(key, (uint8_t)(shannon_from_counts(recs[i].cnt, recs[i].len) < min_shannon ? 3 : 0)).first;
            cls[i] = it->second;
        }
        DevBuf border(c);
        if (border.alloc((size_t)n_border * sizeof(BorderRec)) != cudaSuccess || patch.alloc(n_border) != cudaSuccess)
            return fail("patch allocation failed");
        CK(cudaMemcpyAsync(border.p, recs.data(), (size_t)n_border * sizeof(BorderRec), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(patch.p, cls.data(), n_border, cudaMemcpyHostToDevice, c->stream));
        k_filter_patch<<<(n_border + 255) / 256, 256, 0, c->stream>>>(border.as<BorderRec>(), patch.as<uint8_t>(), n_border,
                                                                     d_bv, classes.as<uint8_t>(), totals.as<unsigned int>());
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
    }
    unsigned long long *out = c->scratch + 128;
    k_filter_cutoff<<<1, 1024, 0, c->stream>>>(totals.as<unsigned int>(), n_blocks, classes.as<uint8_t>(), n,
                                               cut ? (long long)max_reads : -1LL, out);
    c->launches++;
    CK(cudaGetLastError());
    if (cut) {
        k_clear_from<<<grid_for(c, n_bv_words, 256, 8), 256, 0, c->stream>>>(d_bv, out + 4, n_bv_words);
        c->launches++;
        CK(cudaGetLastError());
    }
    unsigned long long h[5];
    CK(cudaMemcpyAsync(h, out, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (counters) for (int i = 0; i < 4; i++) counters[i] = h[i];
    return 0;
}

extern "C" int commet_filter_reads_staged(commet_ctx *c, commet_reads *r, int64_t min_len, int64_t max_N,
                                          float min_shannon, int64_t max_reads, uint32_t *d_bv, uint64_t *counters)
{
    CKR(set_device(c));
    CKR(flush_encode(c, r));
    const uint64_t n = r->n_reads;
    return filter_run(c, n, min_len, max_N, min_shannon, max_reads, d_bv, counters,
                      [&](unsigned n_blocks, const FilterParams &fp, uint64_t n_bv_words, uint8_t *classes, unsigned int *totals,
                          BorderRec *border, unsigned int border_cap, unsigned int *nb) {
                          k_filter<<<n_blocks, kFilterBlock, 0, c->stream>>>(r->planes, r->offs, n, fp, d_bv, n_bv_words, classes,
                                                                             totals, border, border_cap, nb);
                      });
}

extern "C" int commet_filter_reads_range(commet_ctx *c, commet_reads *r, uint64_t first, uint64_t count, int64_t min_len,
                                         int64_t max_N, float min_shannon, int64_t max_reads, uint8_t *bv,
                                         uint64_t *counters)
{
    CKR(set_device(c));
    if (first + count > r->n_reads) return fail("commet_filter_reads_range: range out of bounds");
    CKR(flush_encode(c, r));
    DevBuf d(c);
    const uint64_t nw = tag_words(count);
    if (d.alloc(nw * 4) != cudaSuccess) return fail("filter_reads: selection allocation failed");
    CKR(filter_run(c, count, min_len, max_N, min_shannon, max_reads, d.as<uint32_t>(), counters,
                   [&](unsigned n_blocks, const FilterParams &fp, uint64_t n_bv_words, uint8_t *classes, unsigned int *totals,
                       BorderRec *border, unsigned int border_cap, unsigned int *nb) {
                       k_filter<<<n_blocks, kFilterBlock, 0, c->stream>>>(r->planes, r->offs + first, count, fp, d.as<uint32_t>(),
                                                                          n_bv_words, classes, totals, border, border_cap, nb);
                   }));
    CK(cudaMemcpyAsync(bv, d.p, count / 8 + 1, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// the fused staging + selection kernel; n_blocks counts k_filter blocks of kFilterBlock reads (the unit of `totals`)
template <bool PLANES>
static void launch_stage_filter(commet_ctx *c, unsigned n_blocks, const uint8_t *d_bases, uint64_t readable, uint64_t n_bases,
                                const uint64_t *d_offs, uint64_t n_reads, uint4 *planes, const FilterParams &fp, uint32_t *d_bv,
                                uint64_t n_bv_words, uint8_t *classes, unsigned int *totals, BorderRec *border,
                                unsigned int border_cap, unsigned int *nb)
{
    // four blocks of 256 reads per SM: 6 % faster than two of 512 (profiles/r02_stage_filter_threads_ab.txt; the env selects the other)
    if (env_or("COMMET_B200_SF_THREADS", 256) == 512)
        k_stage_filter<PLANES, 512><<<n_blocks * (kFilterBlock / 512), 512, sf2_tile_words<512>() * 12, c->stream>>>(
                d_bases, readable, n_bases, d_offs, n_reads, planes, fp, d_bv, n_bv_words, classes, totals, border, border_cap, nb);
    else
        k_stage_filter<PLANES, 256><<<n_blocks * (kFilterBlock / 256), 256, sf2_tile_words<256>() * 12, c->stream>>>(
                d_bases, readable, n_bases, d_offs, n_reads, planes, fp, d_bv, n_bv_words, classes, totals, border, border_cap, nb);
}

extern "C" int commet_filter_reads_dev(commet_ctx *c, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                       int64_t min_len, int64_t max_N, float min_shannon, int64_t max_reads,
                                       uint32_t *d_bv, uint64_t *counters)
{
    CKR(set_device(c));
    if ((uintptr_t)d_bases & 15) return fail("commet_filter_reads_dev: d_bases must be 16-byte aligned");
    uint64_t n_bases = 0;
    CK(cudaMemcpyAsync(&n_bases, d_offs + n_reads, sizeof n_bases, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const uint64_t readable = (n_bases + 15) & ~15ull;
    return filter_run(c, n_reads, min_len, max_N, min_shannon, max_reads, d_bv, counters,
                      [&](unsigned n_blocks, const FilterParams &fp, uint64_t n_bv_words, uint8_t *classes, unsigned int *totals,
                          BorderRec *border, unsigned int border_cap, unsigned int *nb) {
                          launch_stage_filter<false>(c, n_blocks, d_bases, readable, n_bases, d_offs, n_reads, nullptr, fp, d_bv, n_bv_words,
                                                     classes, totals, border, border_cap, nb);
                      });
}

// The staging pass and the selection in one kernel: the ASCII bases are read ONCE, the bit-planes of the stream and the
// selection bits of filter_reads come out of the same pass (north_star stage 3).
extern "C" int commet_reads_from_device_filtered(commet_ctx *c, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                                 uint64_t n_bases, int64_t min_len, int64_t max_N, float min_shannon,
                                                 int64_t max_reads, uint32_t *d_bv, uint64_t *counters, commet_reads **out)
{
    if (!c || !d_offs || !out) return fail("commet_reads_from_device_filtered: null argument");
    CKR(set_device(c));
    if ((uintptr_t)d_bases & 15) return fail("commet_reads_from_device_filtered: d_bases must be 16-byte aligned");
    commet_reads *r = nullptr;
    CKR(reads_alloc(c, n_reads, n_bases, &r));
    CK(cudaMemcpyAsync(r->offs, d_offs, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
    const uint64_t readable = (n_bases + 15) & ~15ull;
    int rc = filter_run(c, n_reads, min_len, max_N, min_shannon, max_reads, d_bv, counters,
                        [&](unsigned n_blocks, const FilterParams &fp, uint64_t n_bv_words, uint8_t *classes, unsigned int *totals,
                            BorderRec *border, unsigned int border_cap, unsigned int *nb) {
                            launch_stage_filter<true>(c, n_blocks, d_bases, readable, n_bases, d_offs, n_reads, r->planes, fp, d_bv, n_bv_words,
                                                      classes, totals, border, border_cap, nb);
                        });
    if (rc != 0) { commet_reads_free(r); return rc; }
    *out = r;
    return 0;
}

// host entry: the bases go H2D and through the fused kernel; no bit-planes are built
extern "C" int commet_filter_reads(commet_ctx *c, const uint8_t *bases, const uint64_t *offs, uint64_t n_reads,
                                   int64_t min_len, int64_t max_N, float min_shannon, int64_t max_reads, uint8_t *bv,
                                   uint64_t *counters)
{
    CKR(set_device(c));
    if (offs[0] != 0) return fail("commet_filter_reads: offs[0] must be 0");
    const uint64_t n_bases = offs[n_reads], padded = (n_bases + 15) / 16 * 16 + 16;
    DevBuf d_bases(c), d_offs(c), d(c);
    uint64_t nw = tag_words(n_reads);
    if (d_bases.alloc(padded) != cudaSuccess || d_offs.alloc((n_reads + 1) * sizeof(uint64_t)) != cudaSuccess ||
        d.alloc(nw * 4) != cudaSuccess)
        return fail("filter_reads: device allocation for %llu bases failed", (unsigned long long)n_bases);
    CK(cudaMemsetAsync(d_bases.as<uint8_t>() + (padded - 32), 0, 32, c->stream));
    if (n_bases) CK(cudaMemcpyAsync(d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_offs.p, offs, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CKR(commet_filter_reads_dev(c, d_bases.as<uint8_t>(), d_offs.as<uint64_t>(), n_reads, min_len, max_N, min_shannon,
                                max_reads, d.as<uint32_t>(), counters));
    CK(cudaMemcpyAsync(bv, d.p, n_reads / 8 + 1, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
