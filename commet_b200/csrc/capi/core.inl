// errors, trace, the context's device-memory arena, the context and read-stream types, context life cycle
// (part of the C-ABI library: included by capi.cu, in this order, into one translation unit)

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
