/*
 * commet_b200.h -- C-ABI of the B200-native Commet hot path.
 *
 * The reference (pierrepeterlongo/commet) has no plugin/FFI API: its tools
 * #include header-only classes.  This ABI is what a maintainer would bind in
 * place of those headers; each entry point names the reference interface it
 * replaces (file:line relative to the reference root).  INTEGRATION.md shows
 * the reference-side call sites.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message is in
 *     commet_last_error() (thread-local).  There is NO CPU fallback: without
 *     a CUDA device commet_ctx_create fails.
 *   - plain pointers are caller-owned HOST memory; `d_` pointers are DEVICE
 *     memory of the context's device.
 *   - a read stream is `bases` (concatenated sequence bytes, no separators)
 *     plus `offs[0..n_reads]` byte offsets; read i = bases[offs[i]..offs[i+1]).
 *     It is the "valid read stream" of a set: what FileManager::
 *     get_next_read_to_compare yields (include/file_manager.h:88-112).
 *   - boolean vectors use the .bv payload layout of include/boolean_vector.h:
 *     n/8+1 bytes, bit i of the vector = bit (i%8) of byte i/8.
 *   - one context = one GPU = one CUDA stream; a context is not thread-safe.
 */
#ifndef COMMET_B200_H_
#define COMMET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMMET_B200_ABI_VERSION 1

typedef struct commet_ctx commet_ctx;
typedef struct commet_reads commet_reads;   /* device-resident, 2-bit encoded read stream */

/* ---- context ------------------------------------------------------------ */
const char *commet_last_error(void);
int commet_abi_version(void);
int commet_device_count(void);
int commet_ctx_create(int device, commet_ctx **out);
void commet_ctx_destroy(commet_ctx *ctx);
int commet_ctx_sync(commet_ctx *ctx);
/* CUDA stream of the context as a void* (cudaStream_t); for event timing by the host language. */
void *commet_ctx_stream(commet_ctx *ctx);
/* on != 0: search kernels also count the filter byte tests and k-mer lookups the REFERENCE would perform
 * (stats[5], stats[6] of commet_index_and_search*); a slightly slower instrumented kernel. */
int commet_ctx_count_probes(commet_ctx *ctx, int on);
/* How filters larger than L2 (k >= 28) are fed; the filter bits are the same in every mode.
 *   0: direct RED.OR into the DRAM-resident filter
 *   1: (default) keys partitioned by 32 MiB filter region into a record pool, then applied region by region so that
 *      every RED.OR hits L2.  Two forms exist: 102 (default) = records in slabs handed out on demand, no histogram
 *      pass; 101 = histogram, exclusive scan, scatter into exact ranges (round 1).  101 / 102 select a form explicitly.
 *   16..30: region passes: 2^R passes over the stream, pass r inserting only the keys of the r-th 2^mode-byte
 *      region (region membership = bit-parallel match on the first R bases); measured slower, kept for A/B */
int commet_ctx_binned_index(commet_ctx *ctx, int on);
/* number of kernels launched by this context since creation */
uint64_t commet_ctx_launches(commet_ctx *ctx);

/* pinned host staging memory (cudaHostAlloc) for read buffers */
void *commet_host_alloc(size_t bytes);
void commet_host_free(void *p);

/* ---- parameters fixed by the reference ---------------------------------- */
/* filter size in bytes = 2^(k-1)            include/bloom_filter.h:73-76 */
uint64_t commet_filter_bytes(int k);
/* max k-mers per index chunk = (unsigned long)(1e9 / 2^(33-k))
 *                                            src/index_and_search.cpp:73,146 */
uint64_t commet_max_kmer(int k);

/* ---- read staging: Alphabet + HashKey's per-base coding ------------------
 * Replaces Alphabet::is_in (include/alphabet.h:44-58) and the per-char
 * branches of HashKey::add/rv_add (include/hash_key.h:65-125): bases are
 * turned on the device into bit-planes H (G,T), L (C,T), V (ACGTacgt).
 * upload: H2D copy from host memory + encode.  from_device: encode only. */
int commet_reads_upload(commet_ctx *ctx, const uint8_t *bases, const uint64_t *offs,
                        uint64_t n_reads, commet_reads **out);
/* upload without waiting: copies are queued on the context's copy stream in call order and the encode happens
 * when the stream is first used, so later uploads overlap kernels queued earlier.  `bases` and `offs` must stay
 * valid and unchanged until the stream has been used by a call that synchronises (or commet_ctx_sync after one
 * that does not); page-locked memory (commet_host_alloc) is copied by DMA without host involvement. */
int commet_reads_upload_async(commet_ctx *ctx, const uint8_t *bases, const uint64_t *offs,
                              uint64_t n_reads, commet_reads **out);
int commet_reads_from_device(commet_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offs,
                             uint64_t n_reads, uint64_t n_bases, commet_reads **out);
/* copy of a staged stream on another GPU of this process, over NVLink peer copies (no second parse/upload/encode);
 * the source must be completely staged (commet_reads_upload returns after the encode) and must not be freed or
 * re-selected concurrently; the clone starts with every read selected. */
int commet_reads_clone(commet_ctx *ctx, const commet_reads *src, commet_reads **out);
void commet_reads_free(commet_reads *r);
uint64_t commet_reads_count(const commet_reads *r);
uint64_t commet_reads_bases(const commet_reads *r);

/* Read selection = the input boolean vector of a read file (ReadFile::bv, include/fasta_file.h:143-152,
 * include/fastq_file.h:132-180: get_next_read skips reads whose bit is 0).  A stream staged with ALL the records
 * of a set's files can be re-used under different vectors without another parse or upload: unselected reads are
 * neither indexed, nor counted by the stop rule, nor searched, nor tagged.  bv: n_reads/8+1 bytes in the .bv
 * payload layout over the stream's records (host memory, copied); NULL selects every read again.  This is what
 * lets one process run all N^2-1 index_and_search rounds of Commet.py:186-240 on resident sets. */
int commet_reads_select(commet_ctx *ctx, commet_reads *r, const uint8_t *bv);
uint64_t commet_reads_selected(const commet_reads *r);

/* Number of k-mers index_reads would feed for each read (windows of k
 * consecutive ACGTacgt chars, include/index_reads.h:52-58). counts: n_reads u32. */
int commet_reads_kmer_counts(commet_ctx *ctx, commet_reads *r, int k, uint32_t *counts);
/* their sum only (selected reads), without the per-read download */
int commet_reads_kmer_total(commet_ctx *ctx, commet_reads *r, int k, uint64_t *total);

/* ---- chunk plan: the stop rule of index_reads -----------------------------
 * include/index_reads.h:48-49,60 + src/index_and_search.cpp:255: a chunk ends
 * after the read at which the cumulative k-mer count reaches max_kmer; the
 * next read is fetched and lost.  Fills bounds[2*c] = first read, bounds[2*c+1]
 * = one past the last read of chunk c (cap = capacity in chunks); *n_chunks =
 * number of chunks (may exceed cap: call again with a larger buffer);
 * *n_indexed = reads indexed over all chunks. */
int commet_chunk_plan(commet_ctx *ctx, commet_reads *r, int k, uint64_t max_kmer,
                      uint64_t *bounds, uint64_t cap, uint64_t *n_chunks, uint64_t *n_indexed);

/* ---- stage 1: BloomFilter + index_reads ----------------------------------
 * begin: BloomFilter::BloomFilter (include/bloom_filter.h:61-81): 2^(k-1)
 *        zeroed bytes on the device (re-used across calls with the same k).
 * add:   index_reads' inner loop (include/index_reads.h:49-61) for reads
 *        [first, first+count): HashKey::add + BloomFilter::feed
 *        (include/bloom_filter.h:112-118) for every k-mer.
 * filter_ptr: device address of the bloom_filter.h-layout byte array. */
int commet_index_begin(commet_ctx *ctx, int k);
int commet_index_add(commet_ctx *ctx, commet_reads *r, uint64_t first, uint64_t count);
void *commet_index_filter_ptr(commet_ctx *ctx);
/* copy the filter to host memory (tests) */
int commet_index_download(commet_ctx *ctx, uint8_t *out, uint64_t bytes);
/* copy a host-built filter to the device (tests: probe a reference-built filter) */
int commet_index_upload(commet_ctx *ctx, int k, const uint8_t *filter, uint64_t bytes);
/* multi-GPU merge step: filter |= d_other (word-wise OR of a peer's partial
 * filter, n bytes from byte offset `offset`); d_other may be peer memory. */
int commet_index_or(commet_ctx *ctx, const void *d_other, uint64_t offset, uint64_t bytes);

/* ---- multi-GPU merge of partial filters (north_star; no reference equivalent) ---------
 * One process per GPU.  Each rank indexes its shard of a chunk into its own filter, then:
 *   export:  CUDA IPC handle of this context's filter (valid until the next index_begin with another k)
 *   open:    map a peer rank's filter into this process (peer access over NVLink)
 *   merge:   ONE kernel: pull slice `rank` of every peer's partial, OR, push the merged slice into every
 *            rank's filter.  d_filters[n_ranks]: entry `rank` is ignored (own filter).  The caller brackets
 *            the call with a barrier over all ranks on both sides (partials complete / pushes complete). */
#define COMMET_IPC_HANDLE_BYTES 64
int commet_index_export(commet_ctx *ctx, uint8_t handle[COMMET_IPC_HANDLE_BYTES]);
int commet_peer_open(commet_ctx *ctx, const uint8_t handle[COMMET_IPC_HANDLE_BYTES], void **d_filter);
int commet_peer_close(commet_ctx *ctx, void *d_filter);
int commet_index_merge(commet_ctx *ctx, void *const *d_filters, int n_ranks, int rank);

/* ---- multi-GPU: the chunk loop of src/index_and_search.cpp:255-277 with the index set dealt over the ranks ----
 * One rank = one GPU.  Block b of `block` consecutive reads of the index set's valid-read stream lives on rank
 * b % world: a rank stages only its own blocks (its SHARD, reads in global order), and every chunk is a contiguous
 * range of local reads on every rank.  Per chunk every rank turns its reads into key records; every region of the
 * filter is applied by ONE owner rank, which reads its peers' records out of their memory, and the finished regions
 * are gathered by everybody (or, for small k / COMMET_B200_DIST_MODE=merge: every rank inserts into its own filter
 * and the partial filters are merged by commet_index_merge); then every rank probes its own query streams.  The stop rule of
 * index_reads (include/index_reads.h:48-49,60) is evaluated on k-mer counts the ranks exchange: totals, then
 * per-block totals, then the one read that closes each chunk -- the plan equals commet_chunk_plan's on the whole set.
 *
 * The ranks may be processes or threads; the two collectives the loop needs are callbacks (return 0 on success):
 * barrier(user) returns once every rank has called it; all_gather(user, in, out, bytes) gives every rank the
 * `bytes`-byte contributions of all ranks in rank order (out: world * bytes).  Filters of ranks in other processes
 * are mapped through CUDA IPC, those of threads of this process through peer access. */
typedef struct commet_comm {
    int world, rank;
    void *user;
    int (*barrier)(void *user);
    int (*all_gather)(void *user, const void *in, void *out, uint64_t bytes);
} commet_comm;
typedef struct commet_dist commet_dist;
/* collective: allocates the context's filter for k and maps every peer's.  The context must not get another
 * filter (commet_index_begin / commet_index_and_search* with another k) while the handle is open. */
int commet_dist_open(commet_ctx *ctx, const commet_comm *comm, int k, commet_dist **out);
/* collective.  shard: this rank's blocks of the index set (n_global reads in all); queries / d_tags / searched /
 * shared: this rank's own query streams, as in commet_index_and_search_staged.  stats[12]: chunks, reads indexed
 * by this rank, then host nanoseconds spent in the plan, the inserts, [4] = device ns of the searches, the merges
 * (device ns of the slice all-gathers when [11] is set), the barriers; filter tests and k-mer lookups (with
 * commet_ctx_count_probes); first and end read of the last chunk (what the filter holds when the call returns);
 * [11] = 1 when the chunks' filters were built by owner-applied records + slice all-gather, 0 for partial filters +
 * merge (COMMET_B200_DIST_MODE=merge forces the latter). */
int commet_dist_index_and_search(commet_dist *d, int t, uint64_t max_kmer, commet_reads *shard, uint64_t n_global,
                                 uint64_t block, int n_sets, commet_reads *const *queries, uint32_t *const *d_tags,
                                 uint64_t *searched, uint64_t *shared, uint64_t *stats);
void commet_dist_close(commet_dist *d);
/* The host logic of that loop on its own -- no device, no context -- so that it can be driven on CPU-only hosts (the
 * world_size-2/3 gloo tests; a planner that only has the counts).
 * commet_dist_plan_host: collective; the chunk plan (include/index_reads.h:41-63 on the whole set) from per-read k-mer
 * counts dealt block-cyclically over the ranks: `counts` = this rank's, in local order (n_local of them).  bounds:
 * *n_chunks pairs (first read, end read) in global numbering -- the read after `end` is the one fetched and lost;
 * chunk_kmers[r * *n_chunks + i] = k-mers rank r contributes to chunk i.  Fails if there are more than cap_chunks.
 * commet_dist_deal_regions: the owner rank of every filter region from all ranks' record counts
 * (fills[p * n_bins + b]), as every rank computes it before an owner-applied insert. */
int commet_dist_plan_host(const commet_comm *comm, const uint32_t *counts, uint64_t n_local, uint64_t n_global,
                          uint64_t block, uint64_t max_kmer, uint64_t *bounds, uint64_t cap_chunks, uint64_t *n_chunks,
                          uint64_t *chunk_kmers);
int commet_dist_deal_regions(const uint32_t *fills, int world, int n_bins, int *owner);

/* The same over several GPUs of ONE process, behind the signature of commet_index_and_search: a thread per
 * device stages its shard of the index set and its slice of every query set (contiguous reads, cut at multiples
 * of 32) from the host buffers, runs the loop above and writes its byte range of every tag vector.  What the
 * drop-in index_and_search executable calls when COMMET_B200_GPUS asks for more than one GPU.  stats[8] as
 * commet_index_and_search (stats[2], the k-mer count, is 0; stats[7] = GPUs used). */
typedef struct commet_group commet_group;
int commet_group_create(const int *devices, int n_dev, commet_group **out);
void commet_group_destroy(commet_group *g);
int commet_group_size(const commet_group *g);
int commet_group_index_and_search(commet_group *g, int k, int t, uint64_t max_kmer, const uint8_t *ibases,
                                  const uint64_t *ioffs, uint64_t n_index, int n_sets,
                                  const uint8_t *const *qbases, const uint64_t *const *qoffs,
                                  const uint64_t *n_query, uint8_t *const *tags, uint64_t *searched,
                                  uint64_t *shared, uint64_t *stats);

/* ---- stage 2: search_reads ------------------------------------------------
 * include/search_reads.h:34-87 against the context's current filter: for
 * every read whose bit in `tags` is 0 (FileManager skips tagged reads,
 * include/file_manager.h:99): forward greedy scan, then reverse-complement
 * scan; sets the read's bit when >= t non-overlapping k-mers are found.
 * tags: n_reads/8+1 bytes, read AND written.  *n_found = newly tagged reads
 * (search_reads' return value), *n_searched = reads scanned (:39,44).
 * _dev: tags live on the device as ceil(n/32) u32 words; counters[0]+=found,
 * counters[1]=searched (device u64[4]; [2],[3] += tests, lookups when probe counting is on); nothing is
 * copied to the host. */
int commet_search(commet_ctx *ctx, commet_reads *r, int k, int t, uint8_t *tags,
                  uint64_t *n_found, uint64_t *n_searched);
int commet_search_dev(commet_ctx *ctx, commet_reads *r, int k, int t, uint32_t *d_tags,
                      uint64_t *d_counters);

/* ---- the chunk loop of index_and_search's main ------------------------------
 * src/index_and_search.cpp:241-277: while reads remain: index one chunk,
 * search every query set against it.  tags[s]: n_query[s]/8+1 bytes each,
 * zero-initialised by the callee (file_bvs.set_all_false, file_manager.h:169).
 * searched[s] = last chunk's count, shared[s] = sum over chunks, exactly the
 * numbers of the "[indexed X, searched Y, shared Z]" log line (:286).
 * stats (optional, u64[8]): [0] chunks [1] indexed reads [2] indexed k-mers
 * [3] ns index kernels [4] ns search kernels (device time, CUDA events)
 * [5] filter byte tests [6] k-mer lookups (only with commet_ctx_count_probes)
 * [7] parts the index set was uploaded in (host entry point: part i+1 crosses PCIe while part i is inserted). */
int commet_index_and_search(commet_ctx *ctx, int k, int t, uint64_t max_kmer,
                            const uint8_t *ibases, const uint64_t *ioffs, uint64_t n_index,
                            int n_sets, const uint8_t *const *qbases,
                            const uint64_t *const *qoffs, const uint64_t *n_query,
                            uint8_t *const *tags, uint64_t *searched, uint64_t *shared,
                            uint64_t *stats);
/* same loop on already-staged streams (device-resident inputs) */
int commet_index_and_search_staged(commet_ctx *ctx, int k, int t, uint64_t max_kmer,
                                   commet_reads *index, int n_sets, commet_reads *const *queries,
                                   uint32_t *const *d_tags, uint64_t *searched, uint64_t *shared,
                                   uint64_t *stats);

/* same loop on resident streams (with their current commet_reads_select vectors) and HOST outputs: tags[s] =
 * n_reads/8+1 bytes over ALL records of set s (unselected reads stay 0), ones[s] = device-side popcount of that
 * vector = what `bvop -i` prints for the set's .bv files (src/bvop.cpp:155-160, Commet.py:252-271).  One call =
 * one index_and_search invocation of Commet.py:197,220,233 without parse, upload or process start. */
int commet_index_and_search_resident(commet_ctx *ctx, int k, int t, uint64_t max_kmer, commet_reads *index,
                                     int n_sets, commet_reads *const *queries, uint8_t *const *tags,
                                     uint64_t *searched, uint64_t *shared, uint64_t *ones, uint64_t *stats);

/* ---- stage 3: filter_reads --------------------------------------------------
 * The per-read selection of src/filter_reads.cpp:181-205: length < min_len ->
 * drop; #non-ACGTacgt > max_N -> drop (max_N == -1: no limit; any other
 * negative value drops every read, as the reference's signed compare does); shannon_index
 * (:265-306, float/double mixed precision reproduced exactly) < min_shannon
 * -> drop; stop after max_reads selected (<0: all) and clear every later bit
 * (untag_last_reads, include/read_file.h:76-81).  bv: n_reads/8+1 bytes out.
 * counters[4] = removed by length, by N, by shannon, selected.
 * Reads must be non-empty: the reference loop stops at the first empty read
 * (:188); callers truncate the stream there. */
int commet_filter_reads(commet_ctx *ctx, const uint8_t *bases, const uint64_t *offs,
                        uint64_t n_reads, int64_t min_len, int64_t max_N, float min_shannon,
                        int64_t max_reads, uint8_t *bv, uint64_t *counters);
int commet_filter_reads_staged(commet_ctx *ctx, commet_reads *r, int64_t min_len, int64_t max_N,
                               float min_shannon, int64_t max_reads, uint32_t *d_bv,
                               uint64_t *counters);
/* the same selection on records [first, first+count) of a staged stream (one FILE of a set staged as a whole):
 * bv (host) gets count/8+1 bytes, bit i = record first+i. */
int commet_filter_reads_range(commet_ctx *ctx, commet_reads *r, uint64_t first, uint64_t count, int64_t min_len,
                              int64_t max_N, float min_shannon, int64_t max_reads, uint8_t *bv,
                              uint64_t *counters);
/* the selection fused into the staging pass: ASCII bases already on the device are read once (16-byte vector
 * loads) and only the selection bits are written.  d_bases: 16-byte aligned, readable up to the next multiple of
 * 16 bytes; d_bv: ceil((n_reads/8+1)/4) u32 words.  commet_filter_reads is this after an H2D copy. */
int commet_filter_reads_dev(commet_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                            int64_t min_len, int64_t max_N, float min_shannon, int64_t max_reads,
                            uint32_t *d_bv, uint64_t *counters);

/* the staging pass and the selection in ONE pass over ASCII bases already on the device: the stream's bit-planes
 * (as commet_reads_from_device) and the selection bits (as commet_filter_reads_dev) from a single read of the bases.
 * Same alignment contract as commet_filter_reads_dev. */
int commet_reads_from_device_filtered(commet_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                      uint64_t n_bases, int64_t min_len, int64_t max_N, float min_shannon,
                                      int64_t max_reads, uint32_t *d_bv, uint64_t *counters, commet_reads **out);

/* ---- stage 4: bvop ------------------------------------------------------------
 * BooleanVector::full_and/or/and_not/not (include/boolean_vector.h:418-462)
 * over ALL n_bytes = n/8+1 payload bytes (padding bits included), and nb_one
 * (:244-270): popcount of all bytes clamped to n_bits. */
enum { COMMET_BV_AND = 0, COMMET_BV_OR = 1, COMMET_BV_ANDNOT = 2, COMMET_BV_NOT = 3 };
int commet_bvop(commet_ctx *ctx, int op, const uint8_t *a, const uint8_t *b, uint8_t *out,
                uint64_t n_bytes);
int commet_bv_popcount(commet_ctx *ctx, const uint8_t *bv, uint64_t n_bits, uint64_t *ones);
int commet_bvop_dev(commet_ctx *ctx, int op, const void *d_a, const void *d_b, void *d_out,
                    uint64_t n_bytes);
int commet_bv_popcount_dev(commet_ctx *ctx, const void *d_bv, uint64_t n_bits, uint64_t *ones);
/* nb_one of n device-resident vectors with one read-back and one synchronisation for all of them */
int commet_bv_popcount_batch_dev(commet_ctx *ctx, const void *const *d_bvs, const uint64_t *n_bits, int n, uint64_t *ones);

/* ---- measurement helpers (bench.py / profiles) ---------------------------------
 * random 32-byte-sector gather / atomic-OR ceilings over a `bytes`-sized
 * buffer: n_ops random accesses, returns elapsed ns of the kernel. */
int commet_bench_random_sectors(commet_ctx *ctx, uint64_t bytes, uint64_t n_ops, int atomic_or,
                                double *ns);

#ifdef __cplusplus
}
#endif
#endif /* COMMET_B200_H_ */
