// Multi-GPU chunk loop (included at the end of capi.cu): src/index_and_search.cpp:255-277 with the index set dealt
// block-cyclically over `world` ranks, one GPU each.  Block b of `block` consecutive reads of the set's valid-read
// stream lives on rank b % world, so a rank parses, uploads and encodes only 1/world of the set and every chunk --
// a contiguous range of global reads -- is a contiguous, evenly sized range of LOCAL reads on every rank.
//
//   plan   the stop rule of index_reads (index_reads.h:48-49,60) needs the k-mer counts of all reads in global order.
//          The ranks exchange their local totals (8 bytes each); only if max_kmer is reached at all, their per-BLOCK
//          totals (computed on the device), which they walk together; the read at which a chunk closes is resolved by
//          the rank that owns its block from that block's per-read counts (one 256 KB download) and announced.
//   chunk  every rank clears its filter and inserts ITS reads of the chunk; barrier; ONE kernel per rank ORs slice
//          `rank` of every partial over NVLink peer memory and pushes the merged slice into every filter
//          (k_merge_peers); barrier; every rank probes its own query reads.  Bloom insertion is commutative and
//          idempotent: the merged filter is bit-identical to the single-GPU filter of the chunk.
//
// The ranks may be processes (one per GPU: the filters are mapped through CUDA IPC, the two collectives the loop needs
// -- barrier and all-gather of a few bytes -- are callbacks the host language provides, e.g. over torch.distributed)
// or threads of one process (commet_group_*: peer access, an in-process barrier).  commet_b200/multi.py keeps a
// Python mirror of the plan for the CPU tests.

namespace {

struct PeerInfo {                   // what the ranks exchange in commet_dist_open
    uint64_t pid;
    int32_t device;
    int32_t pad;
    uint64_t ptr;                   // the filter's device address (meaningful inside the owning process)
    uint8_t handle[COMMET_IPC_HANDLE_BYTES];
};

inline uint64_t local_index(uint64_t g, uint64_t world, uint64_t rank, uint64_t block)
{
    // number of reads with global index < g that live on `rank`
    const uint64_t cycle = block * world, full = g / cycle, o = g % cycle;
    const uint64_t lo = rank * block;
    return full * block + (o <= lo ? 0 : std::min(o - lo, block));
}

struct Stopwatch {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    uint64_t lap_ns()
    {
        auto t1 = std::chrono::steady_clock::now();
        uint64_t ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
        t0 = t1;
        return ns;
    }
};

}  // namespace

struct commet_dist {
    commet_ctx *ctx = nullptr;
    commet_comm comm{};
    int k = 0;
    std::vector<void *> peers;          // peer filters as seen from this device (entry `rank`: null)
    std::vector<char> via_ipc;          // opened with cudaIpcOpenMemHandle: closed in commet_dist_close
    // owner-applied insert (kernels.cuh): the peers' record pools, record counters and slab tables as seen from here
    struct PeerBuf { uint64_t raw = 0; void *mapped = nullptr; bool ipc = false; };
    std::vector<PeerBuf> p_recs, p_bins2, p_table;
    std::vector<uint32_t> p_max_q;
};

extern "C" int commet_dist_open(commet_ctx *c, const commet_comm *comm, int k, commet_dist **out)
{
    if (!c || !comm || !out) return fail("commet_dist_open: null argument");
    if (comm->world < 1 || comm->world > kMaxPeers || comm->rank < 0 || comm->rank >= comm->world)
        return fail("commet_dist_open: %d ranks (rank %d) unsupported (1..%d)", comm->world, comm->rank, kMaxPeers);
    if (comm->world > 1 && (!comm->barrier || !comm->all_gather)) return fail("commet_dist_open: the communicator has no callbacks");
    CKR(commet_index_begin(c, k));
    commet_dist *d = new commet_dist;
    d->ctx = c;
    d->comm = *comm;
    d->k = k;
    d->peers.assign(comm->world, nullptr);
    d->via_ipc.assign(comm->world, 0);
    d->p_recs.resize(comm->world);
    d->p_bins2.resize(comm->world);
    d->p_table.resize(comm->world);
    d->p_max_q.assign(comm->world, 0);
    if (comm->world > 1) {
        PeerInfo me{};
        me.pid = (uint64_t)getpid();
        me.device = c->device;
        me.ptr = (uint64_t)(uintptr_t)c->filter;
        int rc = commet_index_export(c, me.handle);
        std::vector<PeerInfo> all(comm->world);
        if (rc == 0 && comm->all_gather(comm->user, &me, all.data(), sizeof me) != 0) rc = fail("commet_dist_open: all_gather failed");
        for (int p = 0; rc == 0 && p < comm->world; p++) {
            if (p == comm->rank) continue;
            if (all[p].pid == me.pid) {
                // a thread of this process: the pointer is valid here; the other device must be accessible
                if (all[p].device != c->device) {
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, c->device, all[p].device) != cudaSuccess || !can) {
                        cudaGetLastError();
                        rc = fail("device %d cannot access the memory of device %d", c->device, all[p].device);
                        break;
                    }
                    cudaError_t pe = cudaDeviceEnablePeerAccess(all[p].device, 0);
                    if (pe != cudaSuccess) cudaGetLastError();          // already enabled: fine
                }
                d->peers[p] = (void *)(uintptr_t)all[p].ptr;
            } else {
                rc = commet_peer_open(c, all[p].handle, &d->peers[p]);
                if (rc == 0) d->via_ipc[p] = 1;
            }
        }
        // everybody has mapped everybody (or failed) before anyone may free its filter
        if (comm->barrier(comm->user) != 0 && rc == 0) rc = fail("commet_dist_open: barrier failed");
        if (rc != 0) { commet_dist_close(d); return rc; }
    }
    *out = d;
    return 0;
}

extern "C" void commet_dist_close(commet_dist *d)
{
    if (!d) return;
    if (d->ctx) {
        cudaSetDevice(d->ctx->device);
        cudaStreamSynchronize(d->ctx->stream);
        for (size_t p = 0; p < d->peers.size(); p++)
            if (d->peers[p] && d->via_ipc[p]) cudaIpcCloseMemHandle(d->peers[p]);
        for (auto *v : {&d->p_recs, &d->p_bins2, &d->p_table})
            for (auto &b : *v)
                if (b.mapped && b.ipc) cudaIpcCloseMemHandle(b.mapped);
    }
    delete d;
}

namespace {

// per-block sums of the per-read k-mer counts of a local shard (block j of the shard = its j-th run of `block` reads)
__global__ void __launch_bounds__(256)
k_block_sums(const uint32_t *__restrict__ counts, uint64_t n, uint64_t block, unsigned long long *__restrict__ sums)
{
    const uint64_t j = blockIdx.x;
    const uint64_t lo = j * block, hi = min(n, lo + block);
    unsigned long long s = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += counts[i];
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += ws[i];
        sums[j] = t;
    }
}

struct Resolved { uint64_t found, index, kmers; };      // found: 0 / 1; index: global read; kmers of the range

// sums of the per-read counts over local read ranges (lo, hi): one block per range
__global__ void __launch_bounds__(256)
k_range_sums(const uint32_t *__restrict__ counts, const uint64_t *__restrict__ ranges, unsigned long long *__restrict__ sums)
{
    const uint64_t lo = ranges[2 * blockIdx.x], hi = ranges[2 * blockIdx.x + 1];
    unsigned long long s = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += counts[i];
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += ws[i];
        sums[blockIdx.x] = t;
    }
}

// Where the per-read k-mer counts of a rank's shard live: on the device (the product's loop) or in a host array
// (commet_dist_plan_host: the same walk, for the CPU tests of the multi-rank logic).  The walk below only ever asks for
// per-block sums, the counts of one block, and sums over a few ranges.
struct DeviceCounts {
    commet_ctx *c;
    const uint32_t *counts;             // device
    uint64_t n_local;
    int block_sums(uint64_t block, uint64_t my_blocks, uint64_t *out)
    {
        DevBuf sums(c);
        if (sums.alloc(my_blocks * sizeof(unsigned long long)) != cudaSuccess) return fail("allocation of block sums failed");
        k_block_sums<<<(unsigned)my_blocks, 256, 0, c->stream>>>(counts, n_local, block, sums.as<unsigned long long>());
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, sums.p, my_blocks * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    int fetch(uint64_t l0, uint64_t l1, std::vector<uint32_t> &blk)
    {
        blk.resize(l1 - l0);
        if (l1 > l0) {
            CK(cudaMemcpyAsync(blk.data(), counts + l0, (l1 - l0) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        return 0;
    }
    int range_sums(const std::vector<uint64_t> &ranges, uint64_t *out)
    {
        const size_t n = ranges.size() / 2;
        DevBuf d_ranges(c), d_sums(c);
        if (d_ranges.alloc(ranges.size() * sizeof(uint64_t)) != cudaSuccess || d_sums.alloc(n * sizeof(unsigned long long)) != cudaSuccess)
            return fail("allocation of chunk sums failed");
        CK(cudaMemcpyAsync(d_ranges.p, ranges.data(), ranges.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        k_range_sums<<<(unsigned)n, 256, 0, c->stream>>>(counts, d_ranges.as<uint64_t>(), d_sums.as<unsigned long long>());
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, d_sums.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
};

struct HostCounts {
    const uint32_t *counts;             // host
    uint64_t n_local;
    int block_sums(uint64_t block, uint64_t my_blocks, uint64_t *out)
    {
        for (uint64_t j = 0; j < my_blocks; j++) {
            uint64_t acc = 0;
            for (uint64_t i = j * block; i < std::min(n_local, (j + 1) * block); i++) acc += counts[i];
            out[j] = acc;
        }
        return 0;
    }
    int fetch(uint64_t l0, uint64_t l1, std::vector<uint32_t> &blk)
    {
        blk.assign(counts + l0, counts + l1);
        return 0;
    }
    int range_sums(const std::vector<uint64_t> &ranges, uint64_t *out)
    {
        for (size_t j = 0; j < ranges.size() / 2; j++) {
            uint64_t acc = 0;
            for (uint64_t i = ranges[2 * j]; i < ranges[2 * j + 1]; i++) acc += counts[i];
            out[j] = acc;
        }
        return 0;
    }
};

// Chunk plan of the whole set from this rank's counts (see the header of this file).  plan: pairs (first, end).
// chunk_kmers[r * n_chunks + i]: k-mers of rank r's reads of chunk i (what its scatter will emit, times four).
// `total`: the sum of this rank's counts.
template <class Counts>
int dist_plan_walk(const commet_comm &cm, Counts &src, uint64_t total, uint64_t n_global, uint64_t block, uint64_t max_kmer,
                   std::vector<uint64_t> &plan, std::vector<uint64_t> &chunk_kmers)
{
    const uint64_t world = (uint64_t)cm.world, rank = (uint64_t)cm.rank;
    const uint64_t n_local = src.n_local;
    plan.clear();
    chunk_kmers.clear();
    std::vector<uint64_t> totals(world, 0);
    uint64_t mine = total;
    if (world > 1) { if (cm.all_gather(cm.user, &mine, totals.data(), sizeof mine) != 0) return fail("all_gather failed"); }
    else totals[0] = mine;
    if (n_global == 0) return 0;
    uint64_t sum = 0;
    for (uint64_t v : totals) sum += v;
    if (sum < max_kmer) {                       // the limit is never reached: one chunk, nothing lost
        plan.push_back(0);
        plan.push_back(n_global);
        chunk_kmers = totals;                   // one chunk: every rank's share is its total, already exchanged
        return 0;
    }
    // per-block totals: n_global / block numbers in all
    const uint64_t n_blocks = (n_global + block - 1) / block;
    const uint64_t my_blocks = (n_local + block - 1) / block, slots = (n_blocks + world - 1) / world;
    std::vector<uint64_t> mine_b(slots, 0), all_b(slots * world, 0);
    if (my_blocks) CKR(src.block_sums(block, my_blocks, mine_b.data()));
    if (world > 1) { if (cm.all_gather(cm.user, mine_b.data(), all_b.data(), slots * sizeof(uint64_t)) != 0) return fail("all_gather failed"); }
    else all_b = mine_b;
    std::vector<uint64_t> csT(n_blocks);          // inclusive prefix over the global blocks
    {
        uint64_t acc = 0;
        for (uint64_t b = 0; b < n_blocks; b++) {
            acc += all_b[(b % world) * slots + b / world];
            csT[b] = acc;
        }
    }
    std::vector<uint32_t> blk;                    // per-read counts of the one block being resolved
    // reads [g0, g1) of ONE block, `cum` k-mers already in the open chunk: the read at which the chunk closes, if any
    auto ask = [&](uint64_t g0, uint64_t g1, uint64_t cum, Resolved &ans) -> int {
        const uint64_t owner = (g0 / block) % world;
        Resolved r{0, 0, 0};
        if (owner == rank) {
            const uint64_t l0 = local_index(g0, world, rank, block), l1 = local_index(g1, world, rank, block);
            CKR(src.fetch(l0, l1, blk));
            uint64_t acc = 0;
            for (uint64_t i = 0; i < l1 - l0; i++) {
                acc += blk[i];
                if (!r.found && cum + acc >= max_kmer) { r.found = 1; r.index = g0 + i; }
            }
            r.kmers = acc;
        }
        if (world > 1) {
            std::vector<Resolved> all(world);
            if (cm.all_gather(cm.user, &r, all.data(), sizeof r) != 0) return fail("all_gather failed");
            ans = all[owner];
        } else ans = r;
        return 0;
    };
    uint64_t i = 0;
    while (i < n_global) {
        const uint64_t b = i / block;
        Resolved a{0, 0, 0};
        CKR(ask(i, std::min((b + 1) * block, n_global), 0, a));            // from read i to the end of its block
        if (!a.found) {
            // whole blocks after b: the first one in which the running count reaches max_kmer
            const uint64_t target = (max_kmer - a.kmers) + csT[b];
            const uint64_t b2 = (uint64_t)(std::lower_bound(csT.begin(), csT.end(), target) - csT.begin());
            if (b2 >= n_blocks) {                                          // the limit is not reached again
                plan.push_back(i);
                plan.push_back(n_global);
                break;
            }
            const uint64_t cum = a.kmers + csT[b2 - 1] - csT[b];
            CKR(ask(b2 * block, std::min((b2 + 1) * block, n_global), cum, a));
            if (!a.found) return fail("chunk plan: block %llu does not close the chunk it should", (unsigned long long)b2);
        }
        plan.push_back(i);
        plan.push_back(a.index + 1);
        i = a.index + 2;                                                   // read index+1 is fetched and lost (index_reads.h:60)
    }
    // every rank's k-mers of every chunk
    const size_t n_chunks = plan.size() / 2;
    std::vector<uint64_t> mine_c(n_chunks, 0);
    if (n_chunks && n_local) {
        std::vector<uint64_t> ranges(2 * n_chunks);
        for (size_t ci = 0; ci < n_chunks; ci++) {
            ranges[2 * ci] = local_index(plan[2 * ci], world, rank, block);
            ranges[2 * ci + 1] = local_index(plan[2 * ci + 1], world, rank, block);
        }
        CKR(src.range_sums(ranges, mine_c.data()));
    }
    chunk_kmers.assign(world * n_chunks, 0);
    if (world > 1 && n_chunks) {
        if (cm.all_gather(cm.user, mine_c.data(), chunk_kmers.data(), n_chunks * sizeof(uint64_t)) != 0) return fail("all_gather failed");
    } else chunk_kmers = mine_c;
    return 0;
}

int dist_plan(commet_dist *d, commet_reads *shard, uint64_t n_global, uint64_t block, uint64_t max_kmer,
              std::vector<uint64_t> &plan, std::vector<uint64_t> &chunk_kmers)
{
    commet_ctx *c = d->ctx;
    const commet_comm &cm = d->comm;
    plan.clear();
    chunk_kmers.clear();
    const uint64_t n_local = local_index(n_global, (uint64_t)cm.world, (uint64_t)cm.rank, block);
    if (shard->n_reads != n_local)
        return fail("rank %d holds %llu reads, its blocks of %llu reads are %llu", cm.rank, (unsigned long long)shard->n_reads,
                    (unsigned long long)n_global, (unsigned long long)n_local);
    DevBuf counts(c);
    unsigned long long total = 0;
    CKR(count_kmers(c, shard, d->k, counts, &total));
    DeviceCounts src{c, counts.as<uint32_t>(), n_local};
    return dist_plan_walk(cm, src, (uint64_t)total, n_global, block, max_kmer, plan, chunk_kmers);
}

// The regions of the filter dealt to the ranks by their record counts (fills[p * n_bins + b]: records rank p holds for region
// b): largest region first, to the rank with the fewest records so far -- every rank computes the same assignment from
// the same numbers.  (+ 1 per region: empty regions are dealt evenly too.)
void deal_regions(const uint32_t *fills, int world, int n_bins, int *owner)
{
    std::vector<uint64_t> tot(n_bins, 0), load(world, 0);
    for (int b = 0; b < n_bins; b++)
        for (int p = 0; p < world; p++) tot[b] += fills[(size_t)p * n_bins + b];
    std::vector<int> order(n_bins);
    for (int b = 0; b < n_bins; b++) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tot[x] > tot[y]; });
    for (int b : order) {
        int best = 0;
        for (int p = 1; p < world; p++)
            if (load[p] < load[best]) best = p;
        owner[b] = best;
        load[best] += tot[b] + 1;
    }
}

struct InsertInfo {                 // what the ranks exchange before an owner-applied insert
    uint64_t pid;
    int32_t device;
    uint32_t max_q;
    uint64_t recs, bins2, table;    // device addresses inside the owning process
    uint8_t h_recs[COMMET_IPC_HANDLE_BYTES], h_bins2[COMMET_IPC_HANDLE_BYTES], h_table[COMMET_IPC_HANDLE_BYTES];
};

int map_peer_buf(commet_ctx *c, commet_dist::PeerBuf &b, uint64_t raw, const uint8_t *handle, bool same_process)
{
    if (b.raw == raw && b.mapped) return 0;                    // unchanged since the last call
    if (b.mapped && b.ipc) CK(cudaIpcCloseMemHandle(b.mapped));
    b = commet_dist::PeerBuf();
    if (same_process) {
        b.mapped = (void *)(uintptr_t)raw;
    } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof h);
        CK(cudaIpcOpenMemHandle(&b.mapped, h, cudaIpcMemLazyEnablePeerAccess));
        b.ipc = true;
    }
    b.raw = raw;
    return 0;
}

// collective: every rank's record pool, counters and slab table (sized for `bound` records here) mapped everywhere
int dist_connect_insert(commet_dist *d, int n_bins, uint64_t bound, uint32_t *max_q)
{
    commet_ctx *c = d->ctx;
    const commet_comm &cm = d->comm;
    int rc = ensure_insert_buffers(c, n_bins, bound, max_q);
    if (rc > 0) rc = fail("no room for the record pool of %llu records", (unsigned long long)bound);
    InsertInfo me{};
    me.pid = (uint64_t)getpid();
    me.device = c->device;
    me.max_q = *max_q;
    if (rc == 0) {
        me.recs = (uint64_t)(uintptr_t)c->recs;
        me.bins2 = (uint64_t)(uintptr_t)c->bins2;
        me.table = (uint64_t)(uintptr_t)c->slab_table;
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, c->recs) == cudaSuccess) memcpy(me.h_recs, &h, sizeof h); else { cudaGetLastError(); rc = fail("cudaIpcGetMemHandle failed"); }
        if (rc == 0 && cudaIpcGetMemHandle(&h, c->bins2) == cudaSuccess) memcpy(me.h_bins2, &h, sizeof h); else if (rc == 0) { cudaGetLastError(); rc = fail("cudaIpcGetMemHandle failed"); }
        if (rc == 0 && cudaIpcGetMemHandle(&h, c->slab_table) == cudaSuccess) memcpy(me.h_table, &h, sizeof h); else if (rc == 0) { cudaGetLastError(); rc = fail("cudaIpcGetMemHandle failed"); }
    }
    if (rc != 0) me.recs = 0;                                   // tells the others that this rank cannot take part
    std::vector<InsertInfo> all(cm.world);
    if (cm.all_gather(cm.user, &me, all.data(), sizeof me) != 0) return fail("all_gather failed");
    for (int p = 0; p < cm.world; p++)
        if (all[p].recs == 0) return rc != 0 ? rc : fail("rank %d could not allocate its record pool", p);
    for (int p = 0; p < cm.world; p++) {
        d->p_max_q[p] = all[p].max_q;
        if (p == cm.rank) {
            d->p_recs[p].mapped = c->recs; d->p_bins2[p].mapped = c->bins2; d->p_table[p].mapped = c->slab_table;
            continue;
        }
        const bool same = all[p].pid == me.pid;
        CKR(map_peer_buf(c, d->p_recs[p], all[p].recs, all[p].h_recs, same));
        CKR(map_peer_buf(c, d->p_bins2[p], all[p].bins2, all[p].h_bins2, same));
        CKR(map_peer_buf(c, d->p_table[p], all[p].table, all[p].h_table, same));
    }
    return 0;
}

}  // namespace

extern "C" int commet_dist_index_and_search(commet_dist *d, int t, uint64_t max_kmer, commet_reads *shard, uint64_t n_global,
                                            uint64_t block, int n_sets, commet_reads *const *queries, uint32_t *const *d_tags,
                                            uint64_t *searched, uint64_t *shared, uint64_t *stats)
{
    if (!d || !shard) return fail("commet_dist_index_and_search: null argument");
    commet_ctx *c = d->ctx;
    const commet_comm &cm = d->comm;
    const int k = d->k;
    CKR(set_device(c));
    if (n_sets < 0) return fail("n_sets=%d unsupported", n_sets);
    if (block == 0) return fail("commet_dist_index_and_search: block must be positive");
    if (shard->sel) return fail("commet_dist_index_and_search: read selections are not supported on a sharded index set");
    if (c->k != k || !c->filter) return fail("commet_dist_index_and_search: the context's filter was re-created since commet_dist_open");
    const uint64_t world = (uint64_t)cm.world, rank = (uint64_t)cm.rank;
    Stopwatch sw;
    uint64_t ns_plan = 0, ns_index = 0, ns_merge = 0, ns_wait = 0;
    std::vector<uint64_t> plan, chunk_kmers;
    if (max_kmer == 0) {
        // nothing is ever inserted (index_reads.h:48: 0 < 0 is false) and every call loses one read: n_global searches
        // of an empty filter tag nothing; one of them gives the same vectors and counters
        if (n_global) { plan.push_back(0); plan.push_back(0); chunk_kmers.assign(world, 0); }
    } else {
        CKR(dist_plan(d, shard, n_global, block, max_kmer, plan, chunk_kmers));
    }
    const size_t n_chunks = plan.size() / 2;
    // ---- how the chunks' filters are built: owner-applied records (kernels.cuh) when the filter has at least one 32 MiB
    // region per rank and every rank's records of a chunk fit the 32-bit record counters; else every rank builds a
    // partial filter and the partials are merged (k_merge_peers).  The decision is taken on exchanged numbers: the same
    // everywhere.
    bool owner_mode = false;
    uint32_t max_q = 0;
    const int n_bins = k >= kRecKeyBits + 2 && k - kRecKeyBits <= 9 ? 1 << (k - kRecKeyBits) : 0;
    uint64_t tile_bound = 0;
    if (world > 1 && n_chunks) {
        uint64_t mine_max = 0, any_max = 0, chunk_max = 0;
        for (size_t ci = 0; ci < n_chunks; ci++) {
            uint64_t tot = 0;
            for (uint64_t r = 0; r < world; r++) { any_max = std::max(any_max, chunk_kmers[r * n_chunks + ci]); tot += chunk_kmers[r * n_chunks + ci]; }
            chunk_max = std::max(chunk_max, tot);
            mine_max = std::max(mine_max, chunk_kmers[rank * n_chunks + ci]);
        }
        const char *mode = getenv("COMMET_B200_DIST_MODE");
        int form = c->insert_form;
        if (const char *e = getenv("COMMET_B200_INSERT")) form = atoi(e);
        owner_mode = !(mode && std::string(mode) == "merge") && c->binned_index && !c->region_passes && form == 2 && n_bins >= 2 &&
                     k >= 28 && n_bins >= (int)world && 4 * any_max < 0xE0000000ull && chunk_max > 0;
        if (owner_mode) {
            CKR(dist_connect_insert(d, n_bins, 4 * mine_max + 64, &max_q));
            tile_bound = 4 * chunk_max / 2048 + (uint64_t)n_bins + 1;
        }
    }
    ns_plan = sw.lap_ns();
    DevBuf cnt_buf(c);
    const size_t n_cnt = 4 * (size_t)std::max(n_sets, 1);
    if (cnt_buf.alloc(n_cnt * sizeof(unsigned long long)) != cudaSuccess) return fail("counter allocation failed");
    unsigned long long *d_cnt = cnt_buf.as<unsigned long long>();
    CK(cudaMemsetAsync(d_cnt, 0, n_cnt * sizeof(unsigned long long), c->stream));
    // (the query streams are prepared by their first search: one that is still crossing PCIe must not hold up the inserts)
    const uint64_t filter_bytes = commet_filter_bytes(k);
    const uint64_t clear_bytes = std::max<uint64_t>((filter_bytes + 255) & ~255ull, 256);
    uint64_t indexed_here = 0;
    SegTimer t_search, t_gather;
    DevBuf tiles_buf(c);
    PeerInsert pi{};
    PeerFilters pf{};
    cudaEvent_t ev_gather = nullptr;
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { if (*e) cudaEventDestroy(*e); } } ev_guard{&ev_gather};
    if (owner_mode) {
        if (tiles_buf.alloc(tile_bound * sizeof(OwnerTile)) != cudaSuccess) return fail("allocation of the tile list failed");
        for (int p = 0; p < cm.world; p++) {
            pi.recs[p] = static_cast<const uint32_t *>(d->p_recs[p].mapped);
            pi.fill[p] = static_cast<const uint32_t *>(d->p_bins2[p].mapped);
            pi.table[p] = static_cast<const uint32_t *>(d->p_table[p].mapped);
            pi.max_q[p] = d->p_max_q[p];
            pf.f[p] = p == cm.rank ? reinterpret_cast<uint4 *>(c->filter) : static_cast<uint4 *>(d->peers[p]);
        }
        CK(cudaEventCreateWithFlags(&ev_gather, cudaEventDisableTiming));
        CKR(prepare(c, shard, k));
    }
    // owner tables: own_bins[n_bins] | fills[pairs] | tbase[pairs + 1] | gather list[n_bins], pairs = (own region, source rank):
    // the regions are dealt by record count, so a rank may own more than n_bins / world of them -- sized for any dealing
    const size_t nb1 = (size_t)std::max(n_bins, 1), cap_pairs = nb1 * world;
    const size_t o_fills = nb1, o_tbase = o_fills + cap_pairs, o_glist = o_tbase + cap_pairs + 1, n_meta = o_glist + nb1;
    DevBuf meta_buf(c);
    if (owner_mode && meta_buf.alloc(n_meta * sizeof(uint32_t)) != cudaSuccess) return fail("allocation of the owner tables failed");
    std::vector<uint32_t> h_fills((size_t)world * nb1), h_meta(n_meta);
    const uint64_t region_bytes = 1ull << kRegionLog2;
    for (size_t ci = 0; ci + 1 < plan.size(); ci += 2) {
        const uint64_t lo = local_index(plan[ci], world, rank, block), hi = local_index(plan[ci + 1], world, rank, block);
        if (owner_mode) {
            uint32_t *d_own = meta_buf.as<uint32_t>(), *d_fills = d_own + o_fills, *d_tbase = d_own + o_tbase, *d_glist = d_own + o_glist;
            unsigned long long *tile_counter = c->bins + 1700;
            if (ci > 0) {
                // my regions of the previous chunk are still being pulled by the peers: nobody clears before everybody has them
                CK(cudaEventSynchronize(ev_gather));
                sw.lap_ns();
                if (cm.barrier(cm.user) != 0) return fail("barrier failed");
                ns_wait += sw.lap_ns();
            }
            sw.lap_ns();
            uint64_t hb[2] = {0, 0};
            if (hi > lo) {
                CK(cudaMemcpyAsync(&hb[0], shard->offs + lo, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
                CK(cudaMemcpyAsync(&hb[1], shard->offs + hi, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                indexed_here += hi - lo;
            }
            CKR(scatter_range(c, shard, hb[0], hb[1], n_bins, max_q));      // an empty range leaves zeroed counters
            uint32_t overflow = 0;
            CK(cudaMemcpyAsync(&overflow, c->bins2 + 1031, sizeof overflow, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));          // my records are complete in my slabs ...
            if (overflow) return fail("record pool too small for chunk %zu", ci / 2);
            ns_index += sw.lap_ns();
            if (cm.barrier(cm.user) != 0) return fail("barrier failed");      // ... and so are everybody else's
            ns_wait += sw.lap_ns();
            // ---- the regions are dealt by their record counts: every rank reads every rank's counters (n_bins numbers each)
            // and computes the same assignment -- largest region first, to the rank with the fewest records so far
            for (int p = 0; p < cm.world; p++)
                CK(cudaMemcpyAsync(h_fills.data() + (size_t)p * n_bins, pi.fill[p], (size_t)n_bins * sizeof(uint32_t), cudaMemcpyDefault, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            std::vector<int> owner(n_bins, 0);
            deal_regions(h_fills.data(), cm.world, n_bins, owner.data());
            uint32_t n_own = 0, n_list = 0;
            for (int b = 0; b < n_bins; b++)
                if (owner[b] == cm.rank) h_meta[n_own++] = (uint32_t)b;
            const uint32_t n_pairs = n_own * (uint32_t)world;
            uint32_t acc = 0;
            for (uint32_t i = 0; i < n_own; i++)
                for (int p = 0; p < cm.world; p++) {
                    const uint32_t f = h_fills[(size_t)p * n_bins + h_meta[i]];
                    h_meta[o_fills + i * world + p] = f;
                    h_meta[o_tbase + i * world + p] = acc;
                    acc += (f + 2047) / 2048;
                }
            h_meta[o_tbase + n_pairs] = acc;
            if (acc > tile_bound) return fail("tile list too small (%u tiles, room for %llu)", acc, (unsigned long long)tile_bound);
            // the regions to pull, round-robin over their owners (concurrent pieces come from different peers)
            {
                std::vector<std::vector<uint32_t>> by_owner(world);
                for (int b = 0; b < n_bins; b++)
                    if (owner[b] != cm.rank) by_owner[owner[b]].push_back((uint32_t)b | ((uint32_t)owner[b] << 16));
                for (size_t j = 0;; j++) {
                    bool any = false;
                    for (uint64_t p = 0; p < world; p++)
                        if (j < by_owner[p].size()) { h_meta[o_glist + n_list++] = by_owner[p][j]; any = true; }
                    if (!any) break;
                }
            }
            CK(cudaMemcpyAsync(d_own, h_meta.data(), n_meta * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemsetAsync(tile_counter, 0, sizeof(unsigned long long), c->stream));
            for (uint32_t i = 0; i < n_own; i++)
                CK(cudaMemsetAsync(reinterpret_cast<uint8_t *>(c->filter) + (uint64_t)h_meta[i] * region_bytes, 0, region_bytes, c->stream));
            // the records of MY regions, from every rank's slabs (peer reads over NVLink inside the apply kernel)
            if (acc) {
                k_owner_tiles<2048><<<c->sm_count * 4, 256, 0, c->stream>>>(pi, cm.world, cm.rank, d_own, (int)n_pairs, d_fills, d_tbase, tiles_buf.as<OwnerTile>());
                k_owner_apply<2048, true><<<c->sm_count * env_or("COMMET_B200_APPLY_BPS", 8), 256, 0, c->stream>>>(
                    c->filter, tiles_buf.as<OwnerTile>(), d_tbase, (int)n_pairs, tile_counter);
                c->launches += 2;
                CK(cudaGetLastError());
            }
            CK(cudaStreamSynchronize(c->stream));          // my regions are final ...
            ns_index += sw.lap_ns();
            if (cm.barrier(cm.user) != 0) return fail("barrier failed");      // ... and so are everybody else's
            ns_wait += sw.lap_ns();
            CKR(t_gather.begin(c->stream));
            if (n_list) {
                k_gather_regions<<<c->sm_count * 8, 256, 0, c->stream>>>(pf, cm.rank, d_glist, n_list, (uint32_t)(region_bytes / 16));
                c->launches++;
                CK(cudaGetLastError());
            }
            CKR(t_gather.end(c->stream));
            CK(cudaEventRecord(ev_gather, c->stream));
        } else {
            CK(cudaMemsetAsync(c->filter, 0, clear_bytes, c->stream));
            sw.lap_ns();
            if (hi > lo) {
                CKR(index_range(c, shard, lo, hi - lo, chunk_kmers.size() == world * n_chunks ? chunk_kmers[rank * n_chunks + ci / 2] : 0));
                indexed_here += hi - lo;
            }
            if (world > 1) {
                CK(cudaStreamSynchronize(c->stream));          // my partial filter is complete on the device ...
                ns_index += sw.lap_ns();
                if (cm.barrier(cm.user) != 0) return fail("barrier failed");      // ... and so is everybody else's
                ns_wait += sw.lap_ns();
                CKR(commet_index_merge(c, d->peers.data(), cm.world, cm.rank));
                CK(cudaStreamSynchronize(c->stream));          // my merged slice has landed in every filter ...
                ns_merge += sw.lap_ns();
                if (cm.barrier(cm.user) != 0) return fail("barrier failed");      // ... and so has everybody else's
                ns_wait += sw.lap_ns();
            }
        }
        CKR(t_search.begin(c->stream));
        for (int s = 0; s < n_sets; s++) {
            CK(cudaMemsetAsync(d_cnt + 4 * s + 1, 0, sizeof(unsigned long long), c->stream));
            CKR(search_launch(c, queries[s], k, t, d_tags[s], d_cnt + 4 * s));
        }
        CKR(t_search.end(c->stream));
    }
    std::vector<unsigned long long> h(n_cnt);
    CK(cudaMemcpyAsync(h.data(), d_cnt, n_cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (world == 1) ns_index += sw.lap_ns();
    // nobody clears or frees its filter while a peer may still be merging into / out of it
    if (world > 1 && cm.barrier(cm.user) != 0) return fail("barrier failed");
    uint64_t n_tests = 0, n_lookups = 0;
    for (int s = 0; s < n_sets; s++) {
        if (shared) shared[s] = h[4 * s];
        if (searched) searched[s] = h[4 * s + 1];
        n_tests += h[4 * s + 2];
        n_lookups += h[4 * s + 3];
    }
    if (stats) {
        stats[0] = max_kmer == 0 ? n_global : plan.size() / 2;
        stats[1] = indexed_here;
        stats[2] = ns_plan;
        stats[3] = ns_index;
        stats[4] = (uint64_t)(t_search.total_ms() * 1e6);
        stats[5] = owner_mode ? (uint64_t)(t_gather.total_ms() * 1e6) : ns_merge;
        stats[6] = ns_wait;
        stats[7] = n_tests;
        stats[8] = n_lookups;
        stats[9] = plan.size() >= 2 ? plan[plan.size() - 2] : 0;       // the last chunk (global reads): what the filter holds now
        stats[10] = plan.size() >= 2 ? plan[plan.size() - 1] : 0;
        stats[11] = owner_mode ? 1 : 0;
    }
    return 0;
}

// ---- host-only entry points to the multi-rank logic (no device, no context): the CPU tests drive them over gloo ------
// The chunk plan of a set whose per-read k-mer counts are dealt block-cyclically over the ranks: `counts` = this rank's
// counts in local order.  bounds: n_chunks pairs (first, end); chunk_kmers[r * n_chunks + i].  Collective.
extern "C" int commet_dist_plan_host(const commet_comm *comm, const uint32_t *counts, uint64_t n_local, uint64_t n_global,
                                     uint64_t block, uint64_t max_kmer, uint64_t *bounds, uint64_t cap_chunks, uint64_t *n_chunks,
                                     uint64_t *chunk_kmers)
{
    if (!comm || comm->world < 1 || comm->rank < 0 || comm->rank >= comm->world || block == 0 || !n_chunks)
        return fail("commet_dist_plan_host: bad arguments");
    if (comm->world > 1 && (!comm->all_gather || !comm->barrier)) return fail("commet_dist_plan_host: callbacks missing");
    if (n_local != local_index(n_global, (uint64_t)comm->world, (uint64_t)comm->rank, block))
        return fail("rank %d holds %llu counts, its blocks of %llu reads are %llu", comm->rank, (unsigned long long)n_local,
                    (unsigned long long)n_global,
                    (unsigned long long)local_index(n_global, (uint64_t)comm->world, (uint64_t)comm->rank, block));
    HostCounts src{counts, n_local};
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_local; i++) total += counts[i];
    std::vector<uint64_t> plan, ck;
    CKR(dist_plan_walk(*comm, src, total, n_global, block, max_kmer, plan, ck));
    *n_chunks = plan.size() / 2;
    if (*n_chunks > cap_chunks) return fail("commet_dist_plan_host: %llu chunks, room for %llu", (unsigned long long)*n_chunks, (unsigned long long)cap_chunks);
    if (bounds) memcpy(bounds, plan.data(), plan.size() * sizeof(uint64_t));
    if (chunk_kmers) memcpy(chunk_kmers, ck.data(), ck.size() * sizeof(uint64_t));
    return 0;
}

// owner[b] of every filter region from the record counts fills[p * n_bins + b] of all ranks (what every rank computes
// before an owner-applied insert)
extern "C" int commet_dist_deal_regions(const uint32_t *fills, int world, int n_bins, int *owner)
{
    if (!fills || !owner || world < 1 || n_bins < 1) return fail("commet_dist_deal_regions: bad arguments");
    deal_regions(fills, world, n_bins, owner);
    return 0;
}

// ------------------------------------------------- several GPUs of ONE process ----
// commet_index_and_search over G devices: a thread per device runs the loop above on its block-cyclic shard of the
// index set and on its slice of every query set (reads [n r / G, n (r+1) / G) cut at multiples of 32, so that the
// slices' tag words are disjoint byte ranges of the set's vector).  The contexts live as long as the group.
struct commet_group {
    std::vector<commet_ctx *> ctx;
    // in-process collectives
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t generation = 0;
    bool failed = false;                // a rank gave up: the others must not wait for it
    std::vector<uint8_t> buf;
};

namespace {

struct GroupRank { commet_group *g; int rank; };

int group_barrier(void *user)
{
    GroupRank *gr = static_cast<GroupRank *>(user);
    commet_group *g = gr->g;
    std::unique_lock<std::mutex> lk(g->m);
    const uint64_t gen = g->generation;
    if (++g->arrived == (int)g->ctx.size()) {
        g->arrived = 0;
        g->generation++;
        g->cv.notify_all();
    } else {
        g->cv.wait(lk, [&] { return g->generation != gen || g->failed; });
    }
    return g->failed ? -1 : 0;
}

int group_all_gather(void *user, const void *in, void *out, uint64_t bytes)
{
    GroupRank *gr = static_cast<GroupRank *>(user);
    commet_group *g = gr->g;
    const size_t world = g->ctx.size();
    {
        std::lock_guard<std::mutex> lk(g->m);
        if (g->buf.size() < world * bytes) g->buf.resize(world * bytes);
    }
    if (group_barrier(user) != 0) return -1;        // the buffer has its size and nobody still reads the previous exchange
    memcpy(g->buf.data() + (size_t)gr->rank * bytes, in, bytes);
    if (group_barrier(user) != 0) return -1;
    memcpy(out, g->buf.data(), world * bytes);
    return group_barrier(user);
}

}  // namespace

extern "C" int commet_group_create(const int *devices, int n_dev, commet_group **out)
{
    if (!devices || !out || n_dev < 1 || n_dev > kMaxPeers) return fail("commet_group_create: 1..%d devices", kMaxPeers);
    commet_group *g = new commet_group;
    g->ctx.assign(n_dev, nullptr);
    std::vector<std::string> errs(n_dev);
    std::vector<std::thread> th;
    for (int r = 0; r < n_dev; r++)
        th.emplace_back([&, r]() { if (commet_ctx_create(devices[r], &g->ctx[r]) != 0) errs[r] = commet_last_error(); });
    for (auto &t : th) t.join();
    for (int r = 0; r < n_dev; r++)
        if (!g->ctx[r]) {
            std::string e = errs[r];
            commet_group_destroy(g);
            return fail("%s", e.c_str());
        }
    *out = g;
    return 0;
}

extern "C" void commet_group_destroy(commet_group *g)
{
    if (!g) return;
    for (commet_ctx *c : g->ctx) commet_ctx_destroy(c);
    delete g;
}

extern "C" int commet_group_size(const commet_group *g) { return g ? (int)g->ctx.size() : 0; }

extern "C" int commet_group_index_and_search(commet_group *g, int k, int t, uint64_t max_kmer, const uint8_t *ibases,
                                             const uint64_t *ioffs, uint64_t n_index, int n_sets,
                                             const uint8_t *const *qbases, const uint64_t *const *qoffs,
                                             const uint64_t *n_query, uint8_t *const *tags, uint64_t *searched,
                                             uint64_t *shared, uint64_t *stats)
{
    if (!g || !ioffs) return fail("commet_group_index_and_search: null argument");
    if (n_sets < 0) return fail("n_sets=%d unsupported", n_sets);
    if (k < 1 || k > kMaxK) return fail("k=%d unsupported (1..%d)", k, kMaxK);
    const int world = (int)g->ctx.size();
    if (world == 1)
        return commet_index_and_search(g->ctx[0], k, t, max_kmer, ibases, ioffs, n_index, n_sets, qbases, qoffs, n_query, tags,
                                       searched, shared, stats);
    uint64_t block = 1 << 16;
    if (const char *e = getenv("COMMET_B200_DIST_BLOCK")) block = std::max<uint64_t>(strtoull(e, nullptr, 10), 1);      // tests
    std::vector<std::string> errs(world);
    std::vector<std::vector<uint64_t>> r_searched(world, std::vector<uint64_t>(n_sets, 0)), r_shared(world, std::vector<uint64_t>(n_sets, 0));
    std::vector<std::vector<uint64_t>> r_stats(world, std::vector<uint64_t>(16, 0));
    std::vector<GroupRank> ranks(world);
    auto body = [&](int r) -> int {
        commet_ctx *c = g->ctx[r];
        CKR(set_device(c));
        // ---- this rank's shard of the index set: its blocks, contiguous on the device ------------------------------
        const uint64_t n_local = local_index(n_index, (uint64_t)world, (uint64_t)r, block);
        std::vector<uint64_t> loffs(n_local + 1, 0);
        {
            uint64_t l = 0;
            for (uint64_t b = (uint64_t)r; b * block < n_index; b += (uint64_t)world) {
                const uint64_t g0 = b * block, g1 = std::min(n_index, g0 + block);
                for (uint64_t i = g0; i < g1; i++, l++) loffs[l + 1] = loffs[l] + (ioffs[i + 1] - ioffs[i]);
            }
        }
        commet_reads *shard = nullptr;
        CKR(reads_alloc(c, n_local, loffs[n_local], &shard));
        struct Cleanup {
            std::vector<commet_reads *> rs; std::vector<void *> bufs; commet_ctx *c; commet_dist *d = nullptr;
            ~Cleanup() { commet_dist_close(d); for (auto *x : rs) commet_reads_free(x); for (void *p : bufs) c->arena.free(p); }
        } cl;
        cl.c = c;
        cl.rs.push_back(shard);
        {
            const uint64_t padded = shard->n_words * 32;
            if (c->arena.alloc((void **)&shard->ascii, padded ? padded : 32) != cudaSuccess) return fail("device allocation of %llu staging bytes failed", (unsigned long long)padded);
            if (padded > shard->n_bases) CK(cudaMemsetAsync(shard->ascii + shard->n_bases, 0, padded - shard->n_bases, c->stream));
            cudaEvent_t ready;
            CKR(take_event(c, &ready));
            CK(cudaEventRecord(ready, c->stream));
            CK(cudaStreamWaitEvent(c->copy_stream, ready, 0));
            c->ev_pool.push_back(ready);
            CKR(h2d_copy(c, shard->offs, loffs.data(), (n_local + 1) * sizeof(uint64_t), false));
            const bool pinned = n_index == 0 || queueable(ibases);
            uint64_t l = 0;
            for (uint64_t b = (uint64_t)r; b * block < n_index; b += (uint64_t)world) {
                const uint64_t g0 = b * block, g1 = std::min(n_index, g0 + block);
                CKR(h2d_copy(c, shard->ascii + loffs[l], ibases + ioffs[g0], ioffs[g1] - ioffs[g0], pinned));
                l += g1 - g0;
            }
            cudaEvent_t e;
            CKR(take_event(c, &e));
            CK(cudaEventRecord(e, c->copy_stream));
            shard->chunk_ev.push_back(e);
            shard->chunk_words = std::max<uint64_t>(shard->n_words, 1);      // one encode over the whole shard
        }
        // ---- this rank's slice of every query set --------------------------------------------------------------------
        std::vector<commet_reads *> q(n_sets, nullptr);
        std::vector<uint32_t *> dt(n_sets, nullptr);
        std::vector<uint64_t> a0(n_sets), a1(n_sets);
        for (int s = 0; s < n_sets; s++) {
            const uint64_t n = n_query[s];
            a0[s] = (n * (uint64_t)r / (uint64_t)world) & ~31ull;
            a1[s] = r + 1 == world ? n : ((n * (uint64_t)(r + 1) / (uint64_t)world) & ~31ull);
            CKR(reads_upload_async(c, qbases[s] + qoffs[s][a0[s]], qoffs[s] + a0[s], a1[s] - a0[s], &q[s]));
            cl.rs.push_back(q[s]);
            const uint64_t nw = tag_words(a1[s] - a0[s]);
            if (c->arena.alloc((void **)&dt[s], nw * 4) != cudaSuccess) return fail("tag allocation failed");
            cl.bufs.push_back(dt[s]);
            CK(cudaMemsetAsync(dt[s], 0, nw * 4, c->stream));
        }
        // ---- the loop ---------------------------------------------------------------------------------------------------
        commet_comm cm;
        cm.world = world;
        cm.rank = r;
        cm.user = &ranks[r];
        cm.barrier = group_barrier;
        cm.all_gather = group_all_gather;
        CKR(commet_dist_open(c, &cm, k, &cl.d));
        CKR(commet_dist_index_and_search(cl.d, t, max_kmer, shard, n_index, block, n_sets, q.data(), dt.data(), r_searched[r].data(),
                                         r_shared[r].data(), r_stats[r].data()));
        for (int s = 0; s < n_sets; s++) {
            const uint64_t m = a1[s] - a0[s];
            const uint64_t bytes = r + 1 == world ? (n_query[s] / 8 + 1) - a0[s] / 8 : m / 8;
            if (bytes) CK(cudaMemcpyAsync(tags[s] + a0[s] / 8, dt[s], bytes, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    };
    // a rank that fails marks the group: the others' collectives return an error instead of waiting for it
    g->failed = false;
    g->arrived = 0;
    std::vector<std::thread> th;
    std::vector<int> rcs(world, 0);
    for (int r = 0; r < world; r++) {
        ranks[r] = GroupRank{g, r};
        th.emplace_back([&, r]() {
            rcs[r] = body(r);
            if (rcs[r] != 0) {
                errs[r] = commet_last_error();
                std::lock_guard<std::mutex> lk(g->m);       // release the ranks waiting for this one
                g->failed = true;
                g->cv.notify_all();
            }
        });
    }
    for (auto &x : th) x.join();
    for (int r = 0; r < world; r++)          // the rank that failed first has the message that matters
        if (rcs[r] != 0 && errs[r].find("barrier failed") == std::string::npos && errs[r].find("all_gather failed") == std::string::npos)
            return fail("GPU %d: %s", g->ctx[r]->device, errs[r].c_str());
    for (int r = 0; r < world; r++)
        if (rcs[r] != 0) return fail("GPU %d: %s", g->ctx[r]->device, errs[r].c_str());
    for (int s = 0; s < n_sets; s++) {
        if (searched) searched[s] = 0;
        if (shared) shared[s] = 0;
        for (int r = 0; r < world; r++) {
            if (searched) searched[s] += r_searched[r][s];
            if (shared) shared[s] += r_shared[r][s];
        }
    }
    if (stats) {
        // same layout as commet_index_and_search: chunks, indexed, k-mers (not counted here), index ns, search ns (device
        // time of the slowest rank), tests, lookups, GPUs
        for (int i = 0; i < 8; i++) stats[i] = 0;
        stats[0] = r_stats[0][0];
        for (int r = 0; r < world; r++) {
            stats[1] += r_stats[r][1];
            stats[3] = std::max(stats[3], r_stats[r][3] + r_stats[r][5] + r_stats[r][6]);
            stats[4] = std::max(stats[4], r_stats[r][4]);
            stats[5] += r_stats[r][7];
            stats[6] += r_stats[r][8];
        }
        stats[7] = (uint64_t)world;
    }
    return 0;
}
