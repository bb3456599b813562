// read staging (H2D copies, encode, W plane), read selection, the chunk plan of a staged set
// (part of the C-ABI library: included by capi.cu, in this order, into one translation unit)

// ----------------------------------------------------------- read staging ---
static int reads_alloc(commet_ctx *c, uint64_t n_reads, uint64_t n_bases, commet_reads **out)
{
    commet_reads *r = new commet_reads;
    r->ctx = c;
    r->n_reads = n_reads;
    r->n_bases = n_bases;
    r->n_words = (n_bases + 31) / 32;
    cudaError_t e = c->arena.alloc((void **)&r->planes, (r->n_words + 4) * sizeof(uint4));
    if (e == cudaSuccess) e = c->arena.alloc((void **)&r->offs, (n_reads + 1) * sizeof(uint64_t));
    if (e != cudaSuccess) {
        commet_reads_free(r);
        return fail("device allocation for %llu bases failed: %s", (unsigned long long)n_bases,
                    cudaGetErrorString(e));
    }
    CK(cudaMemsetAsync(r->planes + r->n_words, 0, 4 * sizeof(uint4), c->stream));
    *out = r;
    return 0;
}

static int launch_encode(commet_ctx *c, const uint8_t *d_bases_padded, commet_reads *r, uint64_t w0, uint64_t w1)
{
    if (w1 <= w0) return 0;
    k_encode<<<grid_for(c, w1 - w0, 256, 8), 256, 0, c->stream>>>(
        reinterpret_cast<const uint4 *>(d_bases_padded) + 2 * w0, r->planes + w0, w1 - w0);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int take_event(commet_ctx *c, cudaEvent_t *e)
{
    if (!c->ev_pool.empty()) {
        *e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return 0;
    }
    CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return 0;
}

constexpr uint64_t kUploadChunk = 32ull << 20;      // bytes of ASCII per H2D copy (multiple of 32)

// cudaMemcpyAsync from pageable memory is staged by the driver and blocks the HOST until the stream gets to
// it; only page-locked (or device/managed) sources can be queued ahead of time
static bool queueable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type != cudaMemoryTypeUnregistered;
}

// H2D copy on the copy stream.  Page-locked sources are handed to the DMA engine as they are.  Pageable sources
// go through a ring of four pinned 32 MB buffers owned by the context: the host thread fills slot i+1 while the
// DMA engine drains slot i.  (cudaMemcpyAsync on pageable memory does the same inside the driver, at a measured
// ~3 GB/s; pinning the whole source first costs ~0.5 s per GB.)
static void host_copy(uint8_t *dst, const uint8_t *src, uint64_t len)
{
    // one core moves ~4 GB/s out of pageable memory on the hosts measured; four keep the DMA engine busier
    constexpr int kThreads = 4;
    if (len < (8u << 20)) { memcpy(dst, src, len); return; }
    std::thread th[kThreads - 1];
    const uint64_t per = (len / kThreads + 4095) & ~4095ull;
    for (int i = 1; i < kThreads; i++) {
        const uint64_t a = std::min(len, per * i), b = std::min(len, per * (i + 1));
        th[i - 1] = std::thread([=]() { if (b > a) memcpy(dst + a, src + a, b - a); });
    }
    memcpy(dst, src, std::min(len, per));
    for (auto &t : th) t.join();
}

static int h2d_copy(commet_ctx *c, void *dst, const void *src, uint64_t bytes, bool pinned_src)
{
    if (bytes == 0) return 0;
    if (pinned_src) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        return 0;
    }
    if (!c->bounce[0]) {
        for (int i = 0; i < 4; i++) {
            CK(cudaHostAlloc((void **)&c->bounce[i], kUploadChunk, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&c->bounce_done[i], cudaEventDisableTiming));
        }
    }
    for (uint64_t off = 0; off < bytes; off += kUploadChunk) {
        const uint64_t len = std::min<uint64_t>(kUploadChunk, bytes - off);
        const unsigned slot = c->bounce_next++ & 3u;
        CK(cudaEventSynchronize(c->bounce_done[slot]));        // never-recorded events are complete
        host_copy(c->bounce[slot], static_cast<const uint8_t *>(src) + off, len);
        CK(cudaMemcpyAsync(static_cast<uint8_t *>(dst) + off, c->bounce[slot], len, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->bounce_done[slot], c->copy_stream));
    }
    return 0;
}

// copy stream: offsets, then the bases in chunks with one arrival event each
static int enqueue_copies(commet_ctx *c, commet_reads *r)
{
    const uint8_t *bases = r->h_bases;
    const uint64_t *offs = r->h_offs;
    r->h_bases = nullptr;
    r->h_offs = nullptr;
    // the copy stream may touch the allocations only after the compute stream has made them
    cudaEvent_t ready;
    CKR(take_event(c, &ready));
    CK(cudaEventRecord(ready, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, ready, 0));
    c->ev_pool.push_back(ready);
    CKR(h2d_copy(c, r->offs, offs, (r->n_reads + 1) * sizeof(uint64_t), queueable(offs)));
    r->chunk_words = kUploadChunk / 32;
    const bool pinned_bases = r->n_bases == 0 || queueable(bases);
    for (uint64_t b = 0; b < r->n_bases || b == 0; b += kUploadChunk) {
        uint64_t len = std::min(kUploadChunk, r->n_bases - b);
        if (len) CKR(h2d_copy(c, r->ascii + b, bases + b, len, pinned_bases));
        cudaEvent_t e;
        CKR(take_event(c, &e));
        CK(cudaEventRecord(e, c->copy_stream));
        r->chunk_ev.push_back(e);
        if (len == 0) break;
    }
    return 0;
}

// Queue the H2D copies of a host read stream on the copy stream; nothing is encoded yet and the host does
// not wait.  flush_encode() later enqueues, on the compute stream, the 2-bit encode of every chunk behind
// its arrival event -- so kernels already queued on the compute stream (the insert of the previous part)
// run while these bytes cross PCIe.  Pageable sources cannot be queued ahead (see queueable): their copies
// are issued by flush_encode, when the data is actually needed.
static int reads_upload_async(commet_ctx *c, const uint8_t *bases, const uint64_t *offs, uint64_t n_reads,
                              commet_reads **out)
{
    const uint64_t base = offs[0];              // a part of a larger stream: `bases` points at its first base
    uint64_t n_bases = offs[n_reads] - base;
    commet_reads *r = nullptr;
    CKR(reads_alloc(c, n_reads, n_bases, &r));
    r->offs_base = base;
    uint64_t padded = r->n_words * 32;
    if (c->arena.alloc((void **)&r->ascii, padded ? padded : 32) != cudaSuccess) {
        commet_reads_free(r);
        return fail("device allocation of %llu staging bytes failed", (unsigned long long)padded);
    }
    if (padded > n_bases) CK(cudaMemsetAsync(r->ascii + n_bases, 0, padded - n_bases, c->stream));
    r->h_bases = bases;
    r->h_offs = offs;
    if (queueable(offs) && (n_bases == 0 || queueable(bases))) CKR(enqueue_copies(c, r));
    *out = r;
    return 0;
}

// compute stream: wait for each chunk, encode it; then release the ASCII staging (stream-ordered)
static int flush_encode(commet_ctx *c, commet_reads *r)
{
    if (!r->ascii) return 0;
    if (r->h_offs) CKR(enqueue_copies(c, r));       // pageable source: copied now
    for (size_t i = 0; i < r->chunk_ev.size(); i++) {
        CK(cudaStreamWaitEvent(c->stream, r->chunk_ev[i], 0));
        if (i == 0 && r->offs_base) {           // the offsets travel before the first chunk of bases
            k_rebase<<<grid_for(c, r->n_reads + 1, 256, 8), 256, 0, c->stream>>>(r->offs, r->n_reads + 1, r->offs_base);
            c->launches++;
            r->offs_base = 0;
        }
        uint64_t w0 = i * r->chunk_words, w1 = std::min(r->n_words, w0 + r->chunk_words);
        CKR(launch_encode(c, r->ascii, r, w0, w1));
        c->ev_pool.push_back(r->chunk_ev[i]);
    }
    r->chunk_ev.clear();
    c->arena.free(r->ascii);           // the next owner's work is ordered behind the encodes just queued
    r->ascii = nullptr;
    return 0;
}

extern "C" int commet_reads_upload(commet_ctx *c, const uint8_t *bases, const uint64_t *offs,
                                   uint64_t n_reads, commet_reads **out)
{
    if (!c || !offs || !out) return fail("commet_reads_upload: null argument");
    if (offs[0] != 0) return fail("commet_reads_upload: offs[0] must be 0");
    CKR(set_device(c));
    commet_reads *r = nullptr;
    CKR(reads_upload_async(c, bases, offs, n_reads, &r));
    int rc = flush_encode(c, r);
    if (rc == 0 && cudaStreamSynchronize(c->stream) != cudaSuccess)
        rc = fail("encode failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != 0) { commet_reads_free(r); return rc; }
    *out = r;
    return 0;
}

// The same staging without waiting: the H2D copies are queued on the context's copy stream (page-locked sources;
// pageable ones are copied when first needed) and the 2-bit encode is enqueued the first time the stream is used
// (index, search, counts, filter), behind the arrival events of its chunks.  Streams uploaded this way cross
// PCIe in call order while kernels queued earlier run: a multi-GPU rank uploads its shard of the reference
// set, then its query set, and the query bytes travel during the insert and the merge.
extern "C" int commet_reads_upload_async(commet_ctx *c, const uint8_t *bases, const uint64_t *offs,
                                         uint64_t n_reads, commet_reads **out)
{
    if (!c || !offs || !out) return fail("commet_reads_upload_async: null argument");
    if (offs[0] != 0) return fail("commet_reads_upload_async: offs[0] must be 0");
    CKR(set_device(c));
    return reads_upload_async(c, bases, offs, n_reads, out);
}

extern "C" int commet_reads_from_device(commet_ctx *c, const uint8_t *d_bases, const uint64_t *d_offs,
                                        uint64_t n_reads, uint64_t n_bases, commet_reads **out)
{
    if (!c || !d_offs || !out) return fail("commet_reads_from_device: null argument");
    CKR(set_device(c));
    commet_reads *r = nullptr;
    CKR(reads_alloc(c, n_reads, n_bases, &r));
    CK(cudaMemcpyAsync(r->offs, d_offs, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
    int rc = 0;
    if (((uintptr_t)d_bases & 15) == 0) {
        // vector-aligned: whole 32-base words are encoded where they lie; a ragged last word goes through a
        // zero-padded 32-byte scratch
        const uint64_t full = n_bases / 32;
        rc = launch_encode(c, d_bases, r, 0, full);
        if (rc == 0 && full < r->n_words) {
            DevBuf tail(c);
            if (tail.alloc(32) != cudaSuccess) { commet_reads_free(r); return fail("staging allocation failed"); }
            CK(cudaMemsetAsync(tail.p, 0, 32, c->stream));
            CK(cudaMemcpyAsync(tail.p, d_bases + full * 32, n_bases - full * 32, cudaMemcpyDeviceToDevice, c->stream));
            k_encode<<<1, 32, 0, c->stream>>>(tail.as<uint4>(), r->planes + full, 1);
            c->launches++;
            CK(cudaGetLastError());
        }
        if (rc == 0) CK(cudaStreamSynchronize(c->stream));
    } else {
        uint64_t padded = r->n_words * 32;
        DevBuf ascii(c);
        if (ascii.alloc(padded) != cudaSuccess) { commet_reads_free(r); return fail("staging allocation failed"); }
        CK(cudaMemsetAsync(ascii.as<uint8_t>() + n_bases, 0, padded - n_bases, c->stream));
        if (n_bases) CK(cudaMemcpyAsync(ascii.p, d_bases, n_bases, cudaMemcpyDeviceToDevice, c->stream));
        rc = launch_encode(c, ascii.as<uint8_t>(), r, 0, r->n_words);
        if (rc == 0) CK(cudaStreamSynchronize(c->stream));
    }
    if (rc != 0) { commet_reads_free(r); return rc; }
    *out = r;
    return 0;
}

// A staged stream copied to another GPU of the same process over NVLink (cudaMemcpyPeerAsync): the planes are
// half a byte per base, so a set that was parsed, uploaded and encoded once reaches every other GPU at peer
// bandwidth instead of crossing PCIe again.  The H/L/V planes are immutable once encoded; the W plane and the
// selection are per-copy state (the clone starts with every read selected and no W plane).
extern "C" int commet_reads_clone(commet_ctx *c, const commet_reads *src, commet_reads **out)
{
    if (!c || !src || !out || !src->ctx) return fail("commet_reads_clone: null argument");
    if (src->ascii || !src->chunk_ev.empty()) return fail("commet_reads_clone: the source stream is still being uploaded");
    CKR(set_device(c));
    if (src->ctx->device != c->device) {
        // direct NVLink path; without peer access the copy is staged through host memory (PCIe twice)
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c->device, src->ctx->device) == cudaSuccess && can) {
            cudaError_t pe = cudaDeviceEnablePeerAccess(src->ctx->device, 0);
            if (pe != cudaSuccess) cudaGetLastError();      // already enabled: fine
        } else {
            cudaGetLastError();
        }
    }
    // the encode (and any kernel that still writes the source's planes) may be in flight on the source
    // context's compute stream: the copy waits for it
    {
        cudaEvent_t done;
        CK(cudaSetDevice(src->ctx->device));
        CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));    // not from the source's pool: another thread owns it
        CK(cudaEventRecord(done, src->ctx->stream));
        CK(cudaSetDevice(c->device));
        CK(cudaStreamWaitEvent(c->stream, done, 0));
        CK(cudaEventDestroy(done));                 // released by the runtime once the recorded work has completed
    }
    commet_reads *r = nullptr;
    CKR(reads_alloc(c, src->n_reads, src->n_bases, &r));
    cudaError_t e = cudaMemcpyPeerAsync(r->planes, c->device, src->planes, src->ctx->device, (src->n_words + 4) * sizeof(uint4), c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyPeerAsync(r->offs, c->device, src->offs, src->ctx->device, (src->n_reads + 1) * sizeof(uint64_t), c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        commet_reads_free(r);
        return fail("peer copy of a staged stream failed: %s", cudaGetErrorString(e));
    }
    *out = r;
    return 0;
}

extern "C" void commet_reads_free(commet_reads *r)
{
    if (!r) return;
    if (r->ctx) cudaSetDevice(r->ctx->device);
    if (r->ascii || !r->chunk_ev.empty()) {         // an upload that was never consumed: let its copies land first
        if (r->ctx) cudaStreamSynchronize(r->ctx->copy_stream);
        for (cudaEvent_t e : r->chunk_ev) { if (r->ctx) r->ctx->ev_pool.push_back(e); else cudaEventDestroy(e); }
        if (r->ascii) {
            if (r->ctx) r->ctx->arena.free(r->ascii); else cudaFree(r->ascii);
        }
    }
    if (r->ctx) {
        if (r->planes) r->ctx->arena.free(r->planes);
        if (r->offs) r->ctx->arena.free(r->offs);
        if (r->sel) r->ctx->arena.free(r->sel);
    }
    delete r;
}

extern "C" uint64_t commet_reads_count(const commet_reads *r) { return r ? r->n_reads : 0; }
extern "C" uint64_t commet_reads_bases(const commet_reads *r) { return r ? r->n_bases : 0; }

// W plane for k (cached per stream)
static int prepare(commet_ctx *c, commet_reads *r, int k)
{
    if (k < 1 || k > kMaxK) return fail("k=%d unsupported (1..%d)", k, kMaxK);
    CKR(flush_encode(c, r));
    if (r->k_prepared == k) return 0;
    if (r->n_words) {
        DevBuf S(c);
        if (S.alloc((r->n_words + 3) * sizeof(uint32_t)) != cudaSuccess) return fail("allocation of start marks failed");
        CK(cudaMemsetAsync(S.p, 0, (r->n_words + 3) * sizeof(uint32_t), c->stream));
        if (r->n_reads) {
            k_mark_starts<<<grid_for(c, r->n_reads, 256, 8), 256, 0, c->stream>>>(r->offs, r->n_reads, S.as<uint32_t>());
            c->launches++;
        }
        k_windows<<<grid_for(c, r->n_words, 256, 8), 256, 0, c->stream>>>(r->planes, S.as<uint32_t>(), r->n_words, k);
        c->launches++;
        if (r->sel && r->n_reads) {
            k_mask_unselected<<<grid_for(c, r->n_reads, 256, 8), 256, 0, c->stream>>>(r->planes, r->offs, r->n_reads, r->sel);
            c->launches++;
        }
        CK(cudaGetLastError());                    // S is released in stream order
    }
    r->k_prepared = k;
    return 0;
}

// ------------------------------------------------------------ read selection --
namespace {
inline bool sel_get(const commet_reads *r, uint64_t i)
{
    return r->h_sel.empty() || ((r->h_sel[i >> 3] >> (i & 7)) & 1u);
}
inline uint64_t sel_count(const commet_reads *r, uint64_t a, uint64_t b)      // selected reads in [a, b)
{
    if (r->h_sel.empty() || b <= a) return b > a ? b - a : 0;
    uint64_t n = 0, i = a;
    for (; i < b && (i & 7); i++) n += (r->h_sel[i >> 3] >> (i & 7)) & 1u;
    for (; i + 8 <= b; i += 8) n += (uint64_t)__builtin_popcount(r->h_sel[i >> 3]);
    for (; i < b; i++) n += (r->h_sel[i >> 3] >> (i & 7)) & 1u;
    return n;
}
}  // namespace

extern "C" int commet_reads_select(commet_ctx *c, commet_reads *r, const uint8_t *bv)
{
    if (!c || !r) return fail("commet_reads_select: null argument");
    CKR(set_device(c));
    r->k_prepared = 0;                              // the W plane depends on the selection
    if (!bv) {
        if (r->sel) { c->arena.free(r->sel); r->sel = nullptr; }
        r->h_sel.clear();
        r->n_selected = r->n_reads;
        return 0;
    }
    const uint64_t nb = r->n_reads / 8 + 1, nw = tag_words(r->n_reads);
    r->h_sel.assign(bv, bv + nb);
    if (r->n_reads & 7) r->h_sel[nb - 1] &= (uint8_t)((1u << (r->n_reads & 7)) - 1u);     // padding bits never select
    else r->h_sel[nb - 1] = 0;
    r->n_selected = 0;
    for (uint64_t i = 0; i < nb; i++) r->n_selected += (uint64_t)__builtin_popcount(r->h_sel[i]);
    if (!r->sel && c->arena.alloc((void **)&r->sel, nw * 4) != cudaSuccess) return fail("selection allocation failed");
    CK(cudaMemsetAsync(r->sel, 0, nw * 4, c->stream));
    // h_sel is owned by the stream object and outlives the copy; pageable source: staged by the driver
    CK(cudaMemcpyAsync(r->sel, r->h_sel.data(), nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" uint64_t commet_reads_selected(const commet_reads *r)
{
    if (!r) return 0;
    return r->sel ? r->n_selected : r->n_reads;
}

extern "C" int commet_reads_kmer_counts(commet_ctx *c, commet_reads *r, int k, uint32_t *counts)
{
    CKR(set_device(c));
    CKR(prepare(c, r, k));
    if (r->n_reads == 0) return 0;
    DevBuf d(c);
    if (d.alloc(r->n_reads * sizeof(uint32_t)) != cudaSuccess) return fail("allocation of k-mer counts failed");
    CK(cudaMemsetAsync(c->scratch + 150, 0, sizeof(unsigned long long), c->stream));
    k_kmer_counts<<<grid_for(c, r->n_reads, 256, 8), 256, 0, c->stream>>>(r->planes, r->offs, r->n_reads,
                                                                         d.as<uint32_t>(), c->scratch + 150);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, d.p, r->n_reads * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// sum of the per-read counts only (8 bytes come back): what a rank of the distributed placement contributes to
// the global "does the set reach max_kmer at all" test (commet_b200/multi.py)
extern "C" int commet_reads_kmer_total(commet_ctx *c, commet_reads *r, int k, uint64_t *total)
{
    if (!c || !r || !total) return fail("commet_reads_kmer_total: null argument");
    CKR(set_device(c));
    CKR(prepare(c, r, k));
    *total = 0;
    if (r->n_reads == 0) return 0;
    DevBuf d(c);
    if (d.alloc(r->n_reads * sizeof(uint32_t)) != cudaSuccess) return fail("allocation of k-mer counts failed");
    CK(cudaMemsetAsync(c->scratch + 150, 0, sizeof(unsigned long long), c->stream));
    k_kmer_counts<<<grid_for(c, r->n_reads, 256, 8), 256, 0, c->stream>>>(r->planes, r->offs, r->n_reads,
                                                                         d.as<uint32_t>(), c->scratch + 150);
    c->launches++;
    CK(cudaGetLastError());
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, c->scratch + 150, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *total = h;
    return 0;
}

// ------------------------------------------------------------- chunk plan ---
static int chunk_plan(commet_ctx *c, commet_reads *r, int k, uint64_t max_kmer,
                      std::vector<uint64_t> &bounds, uint64_t *n_indexed, uint64_t *n_kmers,
                      std::vector<uint64_t> *chunk_kmers = nullptr)
{
    bounds.clear();
    if (chunk_kmers) chunk_kmers->clear();
    uint64_t n = r->n_reads;
    if (n_indexed) *n_indexed = 0;
    if (n_kmers) *n_kmers = 0;
    if (n == 0) return 0;
    CKR(prepare(c, r, k));
    DevBuf d(c);
    if (d.alloc(n * sizeof(uint32_t)) != cudaSuccess) return fail("allocation of k-mer counts failed");
    CK(cudaMemsetAsync(c->scratch + 150, 0, sizeof(unsigned long long), c->stream));
    k_kmer_counts<<<grid_for(c, n, 256, 8), 256, 0, c->stream>>>(r->planes, r->offs, n, d.as<uint32_t>(), c->scratch + 150);
    c->launches++;
    CK(cudaGetLastError());
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, c->scratch + 150, sizeof total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (n_kmers) *n_kmers = total;
    if (total < max_kmer) {            // the limit is never reached: one chunk, nothing dropped
        const uint64_t n_sel = sel_count(r, 0, n);
        if (n_sel) {
            bounds.push_back(0);
            bounds.push_back(n);
            if (chunk_kmers) chunk_kmers->push_back(total);
        }
        if (n_indexed) *n_indexed = n_sel;
        return 0;
    }
    // index_reads.h:48-49,60 + index_and_search.cpp:255: walk the per-read counts (of the selected reads)
    std::vector<uint32_t> cnt(n);
    CK(cudaMemcpyAsync(cnt.data(), d.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    uint64_t i = 0, indexed = 0, kmers = 0;
    while (i < n) {
        uint64_t start = i, cum = 0, taken = 0;
        while (i < n && cum < max_kmer) {
            if (sel_get(r, i)) { cum += cnt[i]; taken++; }
            i++;
        }
        if (taken == 0) break;                 // only unselected reads were left
        bounds.push_back(start);
        bounds.push_back(i);
        indexed += taken;
        kmers += cum;
        if (chunk_kmers) chunk_kmers->push_back(cum);
        if (cum >= max_kmer) {                 // the next valid read is fetched, then lost
            while (i < n && !sel_get(r, i)) i++;
            if (i < n) i++;
        }
    }
    if (n_indexed) *n_indexed = indexed;
    if (n_kmers) *n_kmers = kmers;             // k-mers actually fed (lost reads excluded)
    return 0;
}

extern "C" int commet_chunk_plan(commet_ctx *c, commet_reads *r, int k, uint64_t max_kmer, uint64_t *bounds,
                                 uint64_t cap, uint64_t *n_chunks, uint64_t *n_indexed)
{
    CKR(set_device(c));
    std::vector<uint64_t> b;
    CKR(chunk_plan(c, r, k, max_kmer, b, n_indexed, nullptr));
    uint64_t nc = b.size() / 2;
    if (n_chunks) *n_chunks = nc;
    for (uint64_t i = 0; i < std::min(nc, cap) * 2; i++) bounds[i] = b[i];
    return 0;
}
