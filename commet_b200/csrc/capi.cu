// C-ABI of the B200-native Commet hot path: context, read staging, chunk
// loop, and the launchers of the kernels in kernels.cuh.  See
// include/commet_b200.h for the contract of every entry point and the
// reference interface it replaces.  There is no CPU fallback in this file:
// every data-path operation is a kernel launch on the context's stream.
#include "../../include/commet_b200.h"
#include "kernels.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace commet;

// ------------------------------------------------------------------ errors --
static thread_local std::string g_err;

static int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define CKR(call)                     \
    do {                              \
        int rc_ = (call);             \
        if (rc_ != 0) return rc_;     \
    } while (0)

// ------------------------------------------------------------------ trace ---
// COMMET_B200_TRACE=1: host wall-clock of the phases of the chunk loop on stderr (where does the HOST spend its
// time between the launches -- driver calls that block, allocations, syncs)
#include <chrono>
namespace {
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    HostTrace() : on(getenv("COMMET_B200_TRACE") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(const char *what)
    {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[commet trace] +%8.3f ms (%8.3f) %s\n", std::chrono::duration<double, std::milli>(now - t0).count(),
                std::chrono::duration<double, std::milli>(now - last).count(), what);
        last = now;
    }
};
thread_local HostTrace *g_trace = nullptr;
inline void trace(const char *what) { if (g_trace) g_trace->mark(what); }
}  // namespace

// ------------------------------------------------------------------ arena ---
// Device temporaries (ASCII staging, bit-planes, offsets, tags, counts) come from a context-owned cache of
// cudaMalloc blocks.  Every user of a block touches it on the context's compute stream, or on the copy stream
// behind an event recorded on the compute stream after the allocation, so handing a freed block to the next
// owner needs no device synchronisation: stream order already separates the two uses.  Steady-state calls
// therefore never enter the driver's allocator (cudaMallocAsync was measured to stall the host for 10-60 ms,
// sometimes 500 ms, when a call re-allocates its gigabyte-sized staging buffers).
namespace {
struct Arena {
    struct Block { void *p; size_t cap; bool used; };
    std::vector<Block> blocks;
    static size_t round_up(size_t bytes)
    {
        const size_t g = bytes >= (64u << 20) ? (2u << 20) : bytes >= (1u << 20) ? (256u << 10) : 4096;
        return (std::max<size_t>(bytes, 16) + g - 1) / g * g;
    }
    cudaError_t alloc(void **out, size_t bytes)
    {
        const size_t want = round_up(bytes);
        int best = -1;
        for (size_t i = 0; i < blocks.size(); i++)          // best fit, but never waste more than a fifth of a block:
            // a loose fit lets a small request take the block a later, larger request was sized for, and the
            // cache keeps re-shuffling (and calling cudaMalloc) for several calls before it settles
            if (!blocks[i].used && blocks[i].cap >= want && blocks[i].cap <= want + want / 4 + (1u << 20) &&
                (best < 0 || blocks[i].cap < blocks[best].cap))
                best = (int)i;
        if (best >= 0) {
            blocks[best].used = true;
            *out = blocks[best].p;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {                              // give the cached free blocks back and retry
            cudaGetLastError();
            trim();
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) { cudaGetLastError(); return e; }
        }
        blocks.push_back({p, want, true});
        *out = p;
        return cudaSuccess;
    }
    void free(void *p)
    {
        for (Block &b : blocks)
            if (b.p == p) { b.used = false; return; }
    }
    void trim()                                              // cudaFree synchronises the device: no block is in flight after it
    {
        size_t j = 0;
        for (size_t i = 0; i < blocks.size(); i++) {
            if (blocks[i].used) blocks[j++] = blocks[i];
            else cudaFree(blocks[i].p);
        }
        blocks.resize(j);
    }
};
}  // namespace

// ------------------------------------------------------------------ types ---
struct commet_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;    // compute stream: every kernel launch
    cudaStream_t copy_stream = nullptr;   // H2D staging copies, overlapped with kernels on `stream`
    std::vector<cudaEvent_t> ev_pool;     // recycled chunk-arrival events
    uint32_t *filter = nullptr;       // bloom_filter.h byte array, device
    uint64_t filter_cap = 0;          // allocated bytes
    uint64_t filter_bytes = 0;        // 2^(k-1)
    int k = 0;
    unsigned long long *scratch = nullptr;   // kScratch u64 of device counters
    uint64_t launches = 0;
    bool count_probes = false;        // instrumented search kernel (reference-semantics probe counts)
    int search_dynamic = 0;           // k_search_dyn: lanes take the next read when theirs is done (A/B; see kernels.cuh)
    int search_both = 4;              // both strands in one pass, this many positions per strand and batch (scan_both); 0: forward scan, then reverse (A/B)
    bool binned_index = true;         // L2-blocked insert for DRAM-resident filters
    bool region_passes = false;       // ... by region passes over the stream (false, default: sort keys by region first)
    int region_log2 = 26;             // bytes of filter one pass covers
    uint32_t *recs = nullptr;         // region-sorted key records of the L2-blocked insert
    uint64_t recs_cap = 0;            // capacity in records
    unsigned long long *bins = nullptr;   // hist[512] | base[513] | cursor[512] | tile counter
    int insert_form = 2;              // L2-blocked insert: 1 = histogram + scatter + apply, 2 = slab scatter + apply (kernels.cuh)
    uint32_t *bins2 = nullptr;        // second form: fill[512] | tbase[513] | slab counter
    uint32_t *slab_table = nullptr;   // second form: table[region][slab of the region] -> 1 + slab id
    uint64_t slab_table_cap = 0;      // entries
    unsigned s2_attr = 0;             // k_bin_scatter2<TW> instances whose shared-memory limit has been raised on this device
    Arena arena;                      // cached device temporaries (see Arena)
    // pinned bounce ring for H2D copies out of pageable host memory (see h2d_copy)
    uint8_t *bounce[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t bounce_done[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned bounce_next = 0;
};

struct commet_reads {
    commet_ctx *ctx = nullptr;
    uint64_t n_reads = 0, n_bases = 0, n_words = 0;
    uint4 *planes = nullptr;          // n_words + 4 (zero tail)
    uint64_t *offs = nullptr;         // n_reads + 1, device
    int k_prepared = 0;               // W plane valid for this k (0: none) and the current selection
    // read selection = the input boolean vectors of the set's files (commet_reads_select); null: every read
    uint32_t *sel = nullptr;          // device, ceil((n_reads/8+1)/4) words
    std::vector<uint8_t> h_sel;       // host copy (n_reads/8+1 bytes) for the chunk-boundary walk
    uint64_t n_selected = 0;
    // upload in flight: ASCII chunks arrive on the copy stream, each followed by an event; the
    // encode of a chunk is enqueued on the compute stream behind its event (flush_encode)
    uint8_t *ascii = nullptr;         // device staging of the ASCII bases (pool allocation)
    std::vector<cudaEvent_t> chunk_ev;
    uint64_t chunk_words = 0;         // plane words per chunk
    uint64_t offs_base = 0;           // subtracted from the uploaded offsets on the device (flush_encode)
    const uint8_t *h_bases = nullptr; // host source whose copies are not queued yet (pageable memory)
    const uint64_t *h_offs = nullptr;
};

namespace {

constexpr unsigned kGridBps = 8;       // blocks per SM of the streaming kernels' grids (see grid_for)
constexpr int kScratch = 256;         // [0,4): commet_search counters; [128,256): misc

struct DevBuf {                       // scoped, stream-ordered device temporary from the context's arena
    void *p = nullptr;
    commet_ctx *ctx;
    explicit DevBuf(commet_ctx *c) : ctx(c) {}
    DevBuf(const DevBuf &) = delete;
    ~DevBuf() { if (p) ctx->arena.free(p); }
    cudaError_t alloc(size_t bytes) { return ctx->arena.alloc(&p, bytes); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

inline unsigned env_or(const char *name, unsigned dflt)
{
    const char *e = getenv(name);
    return e && atoi(e) > 0 ? (unsigned)atoi(e) : dflt;
}

// Grid of a grid-stride kernel.  The SMs of a B200 do not all see the same memory bandwidth/latency (two dies), so
// a grid of exactly one resident wave -- every block an equal, static share -- finishes with its slowest SM
// (measured on random DRAM loads: 37.9 G/s with <= 8 blocks per SM, 49.7 G/s with 64).  Several waves of smaller
// shares let the hardware scheduler even it out.
inline unsigned grid_for(const commet_ctx *c, uint64_t items, unsigned block, unsigned blocks_per_sm)
{
    if (blocks_per_sm == 8) blocks_per_sm = env_or("COMMET_B200_GRID_BPS", kGridBps);
    uint64_t need = (items + block - 1) / block;
    uint64_t cap = (uint64_t)c->sm_count * blocks_per_sm;
    if (need < 1) need = 1;
    return (unsigned)std::min<uint64_t>(need, cap);
}

inline int set_device(const commet_ctx *c)
{
    CK(cudaSetDevice(c->device));
    return 0;
}

inline uint64_t tag_words(uint64_t n_reads) { return (n_reads / 8 + 1 + 3) / 4; }

}  // namespace

// ---------------------------------------------------------------- context ---
extern "C" const char *commet_last_error(void) { return g_err.c_str(); }
extern "C" int commet_abi_version(void) { return COMMET_B200_ABI_VERSION; }

extern "C" int commet_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int commet_ctx_create(int device, commet_ctx **out)
{
    if (!out) return fail("commet_ctx_create: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("commet_b200 needs a CUDA device (B200, sm_100a); none visible: %s -- there is no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail("device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    commet_ctx *c = new commet_ctx;
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    // Bloom probes and inserts touch ONE 32-byte sector per key: ask L2 not to pull the neighbouring
    // sectors of the 128-byte line from DRAM with it (ncu: 4x the algorithmic bytes otherwise)
    {
        size_t gran = 32;
        if (const char *e = getenv("COMMET_B200_L2_FETCH")) gran = (size_t)atoi(e);
        if (gran) { if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) cudaGetLastError(); }
    }
    if (const char *e = getenv("COMMET_B200_SEARCH_BOTH")) c->search_both = atoi(e);
    if (const char *e = getenv("COMMET_B200_SEARCH_DYNAMIC")) c->search_dynamic = atoi(e);
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&c->scratch, kScratch * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(c->scratch, 0, kScratch * sizeof(unsigned long long), c->stream));
    *out = c;
    return 0;
}

extern "C" void commet_ctx_destroy(commet_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->filter) cudaFree(c->filter);
    if (c->scratch) cudaFree(c->scratch);
    if (c->recs) cudaFree(c->recs);
    if (c->bins) cudaFree(c->bins);
    if (c->bins2) cudaFree(c->bins2);
    if (c->slab_table) cudaFree(c->slab_table);
    for (Arena::Block &b : c->arena.blocks) cudaFree(b.p);
    for (int i = 0; i < 4; i++) {
        if (c->bounce[i]) cudaFreeHost(c->bounce[i]);
        if (c->bounce_done[i]) cudaEventDestroy(c->bounce_done[i]);
    }
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int commet_ctx_sync(commet_ctx *c)
{
    CKR(set_device(c));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" void *commet_ctx_stream(commet_ctx *c) { return (void *)c->stream; }
extern "C" int commet_ctx_count_probes(commet_ctx *c, int on) { c->count_probes = on != 0; return 0; }
extern "C" int commet_ctx_binned_index(commet_ctx *c, int on)
{
    // 0: direct RED.OR; 1: keys sorted by region first (default); 101 / 102: the same, first / second form of the
    // L2-blocked insert (kernels.cuh) whatever the default is; 16..30: region passes with 2^on-byte regions
    if (on == 101 || on == 102) c->insert_form = on - 100;
    c->binned_index = on != 0;
    c->region_passes = on >= 16 && on <= 30;
    if (c->region_passes) c->region_log2 = on;
    return 0;
}
extern "C" uint64_t commet_ctx_launches(commet_ctx *c) { return c->launches; }

extern "C" void *commet_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void commet_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" uint64_t commet_filter_bytes(int k) { return (uint64_t)1 << (k - 1); }
extern "C" uint64_t commet_max_kmer(int k) { return (uint64_t)(1000000000.0 / pow(2, 33 - k)); }

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

// Shared tail of the two filter kernels: `launch` runs k_filter (bit-planes) or k_filter_ascii (fused with the
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

// ----------------------------------------------------------- stage 4: bvop --
extern "C" int commet_bvop_dev(commet_ctx *c, int op, const void *d_a, const void *d_b, void *d_out, uint64_t n_bytes)
{
    CKR(set_device(c));
    if (op < 0 || op > 3) return fail("unknown bv op %d", op);
    if (n_bytes == 0) return 0;
    if (((uintptr_t)d_a | (uintptr_t)d_out | (op == 3 ? 0 : (uintptr_t)d_b)) & 15) return fail("bvop buffers must be 16-byte aligned");
    uint64_t n_vec = n_bytes / 16;
    unsigned g = grid_for(c, std::max<uint64_t>(n_vec, 16), 256, 8);
    const uint4 *a = static_cast<const uint4 *>(d_a), *b = static_cast<const uint4 *>(d_b);
    uint4 *o = static_cast<uint4 *>(d_out);
    switch (op) {
    case 0: k_bvop<0><<<g, 256, 0, c->stream>>>(a, b, o, n_vec, n_bytes); break;
    case 1: k_bvop<1><<<g, 256, 0, c->stream>>>(a, b, o, n_vec, n_bytes); break;
    case 2: k_bvop<2><<<g, 256, 0, c->stream>>>(a, b, o, n_vec, n_bytes); break;
    default: k_bvop<3><<<g, 256, 0, c->stream>>>(a, a, o, n_vec, n_bytes); break;
    }
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// nb_one of several device-resident vectors: one kernel per vector, ONE read-back and ONE synchronisation for all
// (a count per call costs a D2H copy and a stream sync that dwarf the kernel: 125 MB are counted in 25 us)
extern "C" int commet_bv_popcount_batch_dev(commet_ctx *c, const void *const *d_bvs, const uint64_t *n_bits, int n, uint64_t *ones)
{
    CKR(set_device(c));
    if (n <= 0) return 0;
    if (!d_bvs || !n_bits || !ones) return fail("commet_bv_popcount_batch_dev: null argument");
    DevBuf tot(c);
    if (tot.alloc((size_t)n * sizeof(unsigned long long)) != cudaSuccess) return fail("popcount allocation failed");
    CK(cudaMemsetAsync(tot.p, 0, (size_t)n * sizeof(unsigned long long), c->stream));
    for (int i = 0; i < n; i++) {
        if ((uintptr_t)d_bvs[i] & 15) return fail("bv buffer must be 16-byte aligned");
        const uint64_t n_bytes = n_bits[i] / 8 + 1, n_vec = n_bytes / 16;
        k_popcount<<<grid_for(c, std::max<uint64_t>(n_vec, 16), 256, 8), 256, 0, c->stream>>>(static_cast<const uint4 *>(d_bvs[i]), n_vec,
                                                                                             n_bytes, tot.as<unsigned long long>() + i);
        c->launches++;
    }
    CK(cudaGetLastError());
    std::vector<unsigned long long> h(n);
    CK(cudaMemcpyAsync(h.data(), tot.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) ones[i] = h[i] > n_bits[i] ? n_bits[i] : h[i];      // boolean_vector.h:266-268
    return 0;
}

extern "C" int commet_bv_popcount_dev(commet_ctx *c, const void *d_bv, uint64_t n_bits, uint64_t *ones)
{
    uint64_t one = 0;
    CKR(commet_bv_popcount_batch_dev(c, &d_bv, &n_bits, 1, &one));
    if (ones) *ones = one;
    return 0;
}

extern "C" int commet_bvop(commet_ctx *c, int op, const uint8_t *a, const uint8_t *b, uint8_t *out, uint64_t n_bytes)
{
    CKR(set_device(c));
    if (n_bytes == 0) return 0;
    DevBuf da(c), db(c), dout(c);
    if (da.alloc(n_bytes) != cudaSuccess || dout.alloc(n_bytes) != cudaSuccess || (op != 3 && db.alloc(n_bytes) != cudaSuccess))
        return fail("bvop allocation failed");
    CK(cudaMemcpyAsync(da.p, a, n_bytes, cudaMemcpyHostToDevice, c->stream));
    if (op != 3) CK(cudaMemcpyAsync(db.p, b, n_bytes, cudaMemcpyHostToDevice, c->stream));
    CKR(commet_bvop_dev(c, op, da.p, db.p, dout.p, n_bytes));
    CK(cudaMemcpyAsync(out, dout.p, n_bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int commet_bv_popcount(commet_ctx *c, const uint8_t *bv, uint64_t n_bits, uint64_t *ones)
{
    CKR(set_device(c));
    uint64_t n_bytes = n_bits / 8 + 1;
    DevBuf d(c);
    if (d.alloc(n_bytes) != cudaSuccess) return fail("popcount allocation failed");
    CK(cudaMemcpyAsync(d.p, bv, n_bytes, cudaMemcpyHostToDevice, c->stream));
    return commet_bv_popcount_dev(c, d.p, n_bits, ones);
}

// ------------------------------------------------------------ measurement ---
extern "C" int commet_bench_random_sectors(commet_ctx *c, uint64_t bytes, uint64_t n_ops, int atomic_or, double *ns)
{
    CKR(set_device(c));
    if (bytes < 4096 || (bytes & (bytes - 1))) return fail("bytes must be a power of two >= 4096");
    DevBuf buf(c);
    if (buf.alloc(bytes) != cudaSuccess) return fail("allocation of %llu bytes failed", (unsigned long long)bytes);
    CK(cudaMemsetAsync(buf.p, 0, bytes, c->stream));
    uint64_t mask = bytes / 4 - 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    unsigned g = grid_for(c, n_ops / 4, 256, 64);  // several waves: see grid_for
    for (int rep = 0; rep < 2; rep++) {          // first pass warms up, second is timed
        if (rep == 1) CK(cudaEventRecord(e0, c->stream));
        if (atomic_or) k_random_sectors<true><<<g, 256, 0, c->stream>>>(buf.as<uint32_t>(), mask, n_ops, c->scratch + 144);
        else k_random_sectors<false><<<g, 256, 0, c->stream>>>(buf.as<uint32_t>(), mask, n_ops, c->scratch + 144);
        c->launches++;
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ns) *ns = (double)ms * 1e6;
    return 0;
}

// ------------------------------------------------------------- multi-GPU ----
#include "dist.inl"
