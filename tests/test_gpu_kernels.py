"""GPU parity: every kernel stage through the C-ABI against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import commet_b200
    from commet_b200 import build
    build.build_lib()
    c = commet_b200.Context(0)
    yield c
    c.close()


DIRT = [dict(), dict(p_N=0.02), dict(p_N=0.01, p_lower=0.3, p_other=0.01)]


def oracle_filter(k, stream, first=0, count=None):
    bases, offs = stream
    n = len(offs) - 1
    count = n - first if count is None else count
    f = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)
    sub_offs = offs[first:first + count + 1]
    # index exactly reads [first, first+count): no stop rule (max_kmer huge)
    o2 = (sub_offs - sub_offs[0]).astype(np.uint64)
    b2 = bases[int(sub_offs[0]):int(sub_offs[-1])]
    oracle.index_chunk(f, k, b2, o2, 0, 1 << 62)
    return f


@pytest.mark.parametrize("k", [1, 2, 5, 8, 13, 16, 20, 24])
@pytest.mark.parametrize("dirt", range(3))
def test_index_filter_bit_exact(ctx, k, dirt):
    rng = np.random.default_rng(100 * k + dirt)
    reads = H.make_ref_set(rng, 300, max(1, k - 3), 3 * k + 40, **DIRT[dirt])
    stream = H.to_stream(reads)
    rs = ctx.stage(*stream)
    ctx.index_reads(rs, k)
    got = ctx.filter_download(k)
    assert np.array_equal(got, oracle_filter(k, stream))
    # a sub-range of reads
    ctx.index_reads(rs, k, 17, 101)
    assert np.array_equal(ctx.filter_download(k), oracle_filter(k, stream, 17, 101))


@pytest.mark.parametrize("k,budget", [(28, None), (29, None), (31, None), (28, "5000"), (30, "100000")])
@pytest.mark.parametrize("mode", [101, 102, 26, 22])
def test_index_l2_blocked_path_bit_exact(ctx, k, budget, mode, monkeypatch):
    """Filters larger than L2 are fed region by region -- by sorting key records by region first (mode 1, also
    when the record buffer forces several sub-ranges) or by region passes over the stream (modes 26 and 22 = 64 MiB
    and 4 MiB regions, i.e. few and many passes): same bits as the oracle and as the direct RED.OR path."""
    if budget:
        if mode < 100:
            pytest.skip("the record budget only concerns the sorted path")
        monkeypatch.setenv("COMMET_B200_RECS_BUDGET", budget)
    rng = np.random.default_rng(k)
    reads = H.make_ref_set(rng, 4000, 20, 150, **DIRT[k % 3])
    reads += [b"A" * 200, b"T" * 90, b"ACGT" * 40, b"N" * 50, b"G" * (k - 1), b"C" * k]      # skewed regions
    stream = H.to_stream(reads)
    rs = ctx.stage(*stream)
    exp = oracle_filter(k, stream)
    try:
        ctx.binned_index(mode)
        ctx.index_reads(rs, k)
        got = ctx.filter_download(k)
        assert np.array_equal(got, exp)
        ctx.binned_index(0)
        ctx.index_reads(rs, k, 5, 3000)
        direct = ctx.filter_download(k)
        ctx.binned_index(mode)
        ctx.index_reads(rs, k, 5, 3000)
        assert np.array_equal(ctx.filter_download(k), direct)
    finally:
        ctx.binned_index(102)


@pytest.mark.parametrize("k", [3, 9, 12, 17, 21])
def test_kmer_counts(ctx, k):
    rng = np.random.default_rng(k)
    reads = H.make_ref_set(rng, 500, 1, 4 * k, p_N=0.03, p_lower=0.2)
    stream = H.to_stream(reads)
    rs = ctx.stage(*stream)
    got = ctx.kmer_counts(rs, k)
    exp = []
    for r in reads:
        _, size = oracle.keys(r, k)
        exp.append(int((size >= k).sum()))
    assert got.tolist() == exp


@pytest.mark.parametrize("seed", range(24))
def test_search_against_reference_built_filter(ctx, seed):
    """The filter is built by the ORACLE and uploaded: isolates search_reads."""
    rng = np.random.default_rng(7000 + seed)
    k = int(rng.integers(6, 25))
    t = int(rng.integers(0, 5))
    L = int(rng.integers(k, 5 * k))
    dirt = DIRT[seed % 3]
    ref = H.make_ref_set(rng, 200, max(1, L - 20), L + 20, **dirt)
    qry = H.make_query_set(rng, ref, 400, max(1, L - 20), L + 20, frac_shared=0.6, sub_rate=0.03, **dirt)
    f = oracle_filter(k, H.to_stream(ref))
    qs = H.to_stream(qry)
    exp = np.zeros(len(qry), dtype=np.uint8)
    exp[::7] = 1                                   # pre-tagged reads must be skipped
    st = oracle.search(f, k, t, *qs, exp)
    ctx.filter_upload(k, f)
    rs = ctx.stage(*qs)
    tags = np.zeros(len(qry) // 8 + 1, dtype=np.uint8)
    pre = np.zeros(len(qry), dtype=np.uint8); pre[::7] = 1
    tags[:] = oracle.tags_to_bv(pre)
    found, searched = ctx.search_reads(rs, k, t, tags)
    assert np.array_equal(tags, oracle.tags_to_bv(exp)), (k, t)
    assert (found, searched) == (st["found"], st["searched"])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 27, 31, 32, 33, 34, 35])
def test_search_small_and_large_k(ctx, k):
    """k below the probe batch (the k-jump after a hit lands inside the batch that found it) and k up to 35 (a 16 GiB
    filter), t = 0..3, reads shorter than k, both the one-pass two-strand scan and the reference-order scan (probe
    counting on) against the oracle at EVERY k -- its 2^(k-1)-byte filter is calloc'ed, so only the touched pages
    exist.  Low-complexity references keep small-k filters from saturating."""
    rng = np.random.default_rng(300 + k)
    L = max(3 * k, 12)
    alphabet = np.frombuffer(b"AC" if k < 8 else b"ACGT", dtype=np.uint8)
    ref = [alphabet[rng.integers(0, len(alphabet), size=int(rng.integers(max(1, k - 2), L)))].tobytes() for _ in range(40 if k < 28 else 300)]
    qry = H.make_query_set(rng, ref, 300, max(1, k - 3), L, frac_shared=0.5, sub_rate=0.05, p_N=0.03)
    for t in range(0, 4):
        exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(qry)], 1 << 60)
        for count in (False, True):
            ctx.count_probes(count)
            tags, info = ctx.index_and_search(k, t, H.to_stream(ref), [H.to_stream(qry)], 1 << 60)
            assert np.array_equal(tags[0], oracle.tags_to_bv(exp_tags[0])), (k, t, count)
            assert info["shared"] == exp["shared"]
            if count:
                assert info["tests"] == exp["tests"] and info["lookups"] == exp["lookups"], (k, t)
                ref_order = tags[0].copy()
            else:
                one_pass = tags[0].copy()
        ctx.count_probes(False)
        assert np.array_equal(one_pass, ref_order), (k, t)        # the two scan orders agree at every k


@pytest.mark.parametrize("seed", range(30))
def test_index_and_search_chunk_loop(ctx, seed):
    """Full chunk loop incl. the dropped read at every chunk boundary and the log counters."""
    rng = np.random.default_rng(8000 + seed)
    k = int(rng.integers(8, 19))
    t = int(rng.integers(0, 4))
    L = int(rng.integers(k, 4 * k))
    dirt = DIRT[seed % 3]
    maxk = oracle.max_kmer(k)
    n_ref = int(min(3000, max(20, 4 * maxk // max(1, (L - k + 1)) + 5)))
    ref = H.make_ref_set(rng, n_ref, max(1, L - 10), L + 10, **dirt)
    queries = [H.make_query_set(rng, ref, int(rng.integers(1, 400)), max(1, L - 10), L + 10, **dirt)
               for _ in range(int(rng.integers(1, 4)))]
    exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries])
    tags, info = ctx.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries])
    assert info["chunks"] == exp["chunks"] and info["indexed"] == exp["indexed"] and info["kmers"] == exp["kmers"]
    for s in range(len(queries)):
        assert np.array_equal(tags[s], oracle.tags_to_bv(exp_tags[s])), (seed, k, t, s)
    assert info["searched"] == exp["searched"] and info["shared"] == exp["shared"]


@pytest.mark.parametrize("seed", range(24))
def test_chunk_loop_over_uploaded_parts(ctx, seed, monkeypatch):
    """Host index sets are uploaded in three parts (20/30/50 % of the bases) that are inserted while the next
    one crosses PCIe: chunk boundaries and the fetched-and-lost read must come out the same wherever they fall
    relative to the part boundaries (inside a part, on its last read, on the first read of the next part).  Large query
    sets are uploaded and searched in up to four parts cut at multiples of 32 reads (forced small here too)."""
    monkeypatch.setenv("COMMET_B200_PART_BYTES", str(int(np.random.default_rng(seed).integers(50, 3000))))
    monkeypatch.setenv("COMMET_B200_QUERY_PART_BYTES", str(int(np.random.default_rng(seed + 100).integers(40, 4000))))
    rng = np.random.default_rng(9000 + seed)
    k = int(rng.integers(8, 19))
    t = int(rng.integers(1, 3))
    L = int(rng.integers(k, 4 * k))
    n_ref = int(rng.integers(30, 1500))
    ref = H.make_ref_set(rng, n_ref, max(1, L - 10), L + 10, **DIRT[seed % 3])
    total_kmers = sum(max(0, len(r) - k + 1) for r in ref)
    # from "never reached" to dozens of chunks, including limits that land exactly on a read's last k-mer
    maxk = [None, max(1, total_kmers // 2), max(1, total_kmers // 7), max(1, total_kmers // 40), 1][seed % 5]
    queries = [H.make_query_set(rng, ref, int(rng.integers(1, 600)), max(1, L - 10), L + 10, **DIRT[seed % 3])
               for _ in range(2)]
    exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    tags, info = ctx.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    assert info["chunks"] == exp["chunks"] and info["indexed"] == exp["indexed"] and info["kmers"] == exp["kmers"]
    for s in range(len(queries)):
        assert np.array_equal(tags[s], oracle.tags_to_bv(exp_tags[s])), (seed, k, t, s)
    assert info["searched"] == exp["searched"] and info["shared"] == exp["shared"]


@pytest.mark.parametrize("seed", range(24))
def test_selection_equals_compacted_streams(ctx, seed):
    """commet_reads_select: a stream staged with ALL records of a file, restricted by an input boolean vector,
    must behave exactly like the reference's valid-read stream (fasta_file.h:143-152): chunk boundaries, the
    fetched-and-lost read, the counters and every tag bit equal the oracle run on the compacted streams."""
    import torch
    rng = np.random.default_rng(12000 + seed)
    k = int(rng.integers(8, 19))
    t = int(rng.integers(1, 3))
    L = int(rng.integers(k, 4 * k))
    ref = H.make_ref_set(rng, int(rng.integers(30, 1500)), max(1, L - 10), L + 10, **DIRT[seed % 3])
    queries = [H.make_query_set(rng, ref, int(rng.integers(1, 300)), max(1, L - 10), L + 10, **DIRT[seed % 3])
               for _ in range(2)]
    p_sel = [0.5, 0.9, 0.1, 0.0][seed % 4]
    masks = [rng.random(len(x)) < (p_sel if i == 0 else 0.7) for i, x in enumerate([ref] + queries)]
    if seed % 6 == 5:
        masks[1][:] = False                           # a query set with no valid read at all
    sub = [[r for r, m in zip(x, mk) if m] for x, mk in zip([ref] + queries, masks)]
    total_kmers = sum(max(0, len(r) - k + 1) for r in sub[0])
    maxk = [None, max(1, total_kmers // 2), max(1, total_kmers // 9), 1][(seed // 4) % 4]
    exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(sub[0]), [H.to_stream(q) for q in sub[1:]], maxk)

    streams = [ctx.stage(*H.to_stream(x)) for x in [ref] + queries]
    for st, mk in zip(streams, masks):
        ctx.select(st, np.packbits(np.concatenate([mk, np.zeros(8, bool)]), bitorder="little")[:len(mk) // 8 + 1])
    d_tags = [torch.zeros((len(q) // 8 + 4) // 4 + 1, dtype=torch.int32, device="cuda") for q in queries]
    info = ctx.index_and_search_staged(k, t, streams[0], streams[1:], [x.data_ptr() for x in d_tags], maxk)
    ctx.sync()
    assert info["chunks"] == exp["chunks"] and info["indexed"] == exp["indexed"] and info["kmers"] == exp["kmers"]
    assert info["searched"] == exp["searched"] and info["shared"] == exp["shared"]
    for s, q in enumerate(queries):
        got = np.unpackbits(d_tags[s].cpu().numpy().view(np.uint8), bitorder="little")[:len(q)].astype(bool)
        want = np.zeros(len(q), bool)
        want[np.flatnonzero(masks[s + 1])] = np.asarray(exp_tags[s], dtype=bool)[:int(masks[s + 1].sum())]
        assert np.array_equal(got, want), (seed, s)
    # back to "every read": the same streams, no vector
    for st in streams:
        ctx.select(st, None)
    exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    for x in d_tags:
        x.zero_()
    info = ctx.index_and_search_staged(k, t, streams[0], streams[1:], [x.data_ptr() for x in d_tags], maxk)
    ctx.sync()
    assert info["chunks"] == exp["chunks"] and info["indexed"] == exp["indexed"] and info["shared"] == exp["shared"]
    for st in streams:
        st.free()


@pytest.mark.parametrize("seed", range(6))
def test_probe_counts_match_reference_semantics(ctx, seed):
    """N_probes: the instrumented kernel counts exactly the byte tests / lookups the reference does."""
    rng = np.random.default_rng(8100 + seed)
    k = int(rng.integers(9, 16))            # small k: dense filter, many a/b/c passes
    t = int(rng.integers(1, 4))
    ref = H.make_ref_set(rng, 3000, 40, 90, **DIRT[seed % 3])
    qry = H.make_query_set(rng, ref, 800, 40, 90, **DIRT[seed % 3])
    exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(qry)])
    ctx.count_probes(True)
    try:
        tags, info = ctx.index_and_search(k, t, H.to_stream(ref), [H.to_stream(qry)])
    finally:
        ctx.count_probes(False)
    assert np.array_equal(tags[0], oracle.tags_to_bv(exp_tags[0]))
    assert (info["tests"], info["lookups"]) == (exp["tests"], exp["lookups"])


def test_chunk_plan_matches_oracle_walk(ctx):
    rng = np.random.default_rng(5)
    k = 14
    reads = H.make_ref_set(rng, 2000, 10, 80, p_N=0.02)
    stream = H.to_stream(reads)
    rs = ctx.stage(*stream)
    chunks, indexed = ctx.chunk_plan(rs, k)
    f = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)
    pos, exp, tot = 0, [], 0
    while pos < len(reads):
        nxt, ni, _ = oracle.index_chunk(f, k, *stream, pos, oracle.max_kmer(k))
        exp.append((pos, pos + ni)); tot += ni; pos = nxt
    assert chunks == exp and indexed == tot and len(chunks) > 3


def test_empty_and_degenerate_inputs(ctx):
    k, t = 11, 2
    empty = (np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    one = H.to_stream([b"ACGTACGTACGTACGTACGT"])
    short = H.to_stream([b"ACG", b"A", b"ACGTACGTAC"])      # all shorter than k
    tags, info = ctx.index_and_search(k, t, empty, [one])
    assert info["chunks"] == 0 and tags[0].tolist() == [0]
    tags, info = ctx.index_and_search(k, t, one, [empty, short, one])
    e_tags, e = oracle.index_and_search(k, t, one, [empty, short, one])
    assert [x.tolist() for x in tags] == [oracle.tags_to_bv(x).tolist() for x in e_tags]
    assert info["searched"] == e["searched"] and info["shared"] == e["shared"]


def test_k33_dram_resident_filter(ctx):
    """k=33 (4 GiB filter, 33-bit keys): bit-exact against the oracle on 60k reads."""
    rng = np.random.default_rng(33)
    ref = H.make_ref_set(rng, 30000, 100)
    qry = H.make_query_set(rng, ref, 30000, 100)
    e_tags, e = oracle.index_and_search(33, 2, H.to_stream(ref), [H.to_stream(qry)])
    tags, info = ctx.index_and_search(33, 2, H.to_stream(ref), [H.to_stream(qry)])
    assert np.array_equal(tags[0], oracle.tags_to_bv(e_tags[0]))
    assert info["shared"] == e["shared"] and 0.4 < info["shared"][0] / len(qry) < 0.6


@pytest.mark.parametrize("seed", range(30))
def test_filter_reads(ctx, seed):
    rng = np.random.default_rng(6000 + seed)
    n = int(rng.integers(1, 5000))
    reads = []
    for _ in range(n):
        L = int(rng.integers(1, 200))
        kind = rng.random()
        if kind < 0.15:
            r = np.full(L, ord(rng.choice(list("ACGTacgtN"))), dtype=np.uint8)
        elif kind < 0.3:
            r = np.tile(np.frombuffer(bytes(rng.choice([b"AC", b"AT", b"ACGT", b"AAC", b"AACG"])), dtype=np.uint8), L)[:L]
        else:
            r = H.dirty(rng, H.random_read(rng, L), p_N=float(rng.choice([0, 0.05])),
                        p_lower=float(rng.choice([0, 0.4])), p_other=float(rng.choice([0, 0.02])))
        reads.append(r.tobytes())
    kw = {}
    if rng.random() < 0.8: kw["min_len"] = int(rng.integers(0, 120))
    if rng.random() < 0.6: kw["max_N"] = int(rng.integers(0, 6))
    if rng.random() < 0.85: kw["min_shannon"] = float(np.float32(rng.choice([1.0, 1.5, 2.0, 0.5, float(rng.uniform(0, 2.1))])))
    if rng.random() < 0.5: kw["max_reads"] = int(rng.choice([0, 1, n // 3, n, n + 5, 1024, 1023, 1025]))
    stream = H.to_stream(reads)
    exp_bv, exp_cnt = oracle.filter_reads(*stream, **kw)
    bv, cnt = ctx.filter_reads(*stream, **kw)               # one pass over the ASCII bases, no bit-planes (k_stage_filter<false>)
    assert cnt == exp_cnt, (seed, kw)
    assert np.array_equal(bv, exp_bv), (seed, kw)
    import torch                                             # the same selection on an already staged stream (k_filter)
    rs = ctx.stage(*stream)
    d_bv = torch.zeros((n // 8 + 1 + 3) // 4, dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    cnt2 = ctx.filter_reads_staged(rs, d_bv.data_ptr(), **kw)
    assert cnt2 == exp_cnt, (seed, kw)
    assert np.array_equal(d_bv.cpu().numpy().view(np.uint8)[:n // 8 + 1], exp_bv), (seed, kw)
    # staging and selection in ONE pass (k_stage_filter<true>): the same bits, and bit-planes that index like k_encode's
    bases_d = torch.zeros((len(stream[0]) + 15) // 16 * 16 + 16, dtype=torch.uint8, device="cuda:0")
    bases_d[:len(stream[0])] = torch.as_tensor(stream[0], device="cuda:0")
    if seed % 3 == 0:
        bases_d[len(stream[0]):] = ord("A")                   # bytes past the end of the stream are not bases
    offs_d = torch.as_tensor(stream[1].astype(np.int64), device="cuda:0")
    d_bv3 = torch.zeros_like(d_bv)
    torch.cuda.synchronize()
    rs3, cnt3 = ctx.stage_device_filtered(bases_d.data_ptr(), offs_d.data_ptr(), n, len(stream[0]), d_bv3.data_ptr(), **kw)
    assert cnt3 == exp_cnt, (seed, kw)
    assert np.array_equal(d_bv3.cpu().numpy().view(np.uint8)[:n // 8 + 1], exp_bv), (seed, kw)
    k = int(rng.integers(1, 25))
    ctx.index_reads(rs, k)
    f_encode = ctx.filter_download(k)
    ctx.index_reads(rs3, k)
    assert np.array_equal(ctx.filter_download(k), f_encode), (seed, k)
    assert np.array_equal(ctx.kmer_counts(rs3, k), ctx.kmer_counts(rs, k))
    rs.free()
    rs3.free()


@pytest.mark.parametrize("cap", ["64", None])
def test_filter_reads_more_undecided_reads_than_the_record_buffer(ctx, cap, monkeypatch):
    """dinucleotide repeats have H = 1.0 exactly: with -e 1 every one of them sits on the threshold and is re-decided on
    the host from its exact counts.  More of them than the device buffer holds (2^20; 64 here for the small case) run the
    pass again with a buffer of the size asked for -- the reference has no such limit."""
    rng = np.random.default_rng(99)
    n = 3000 if cap else (1 << 20) + 50_000
    if cap:
        monkeypatch.setenv("COMMET_B200_BORDER_CAP", cap)
    pats = [np.frombuffer(p, dtype=np.uint8) for p in (b"AC", b"GT", b"TA", b"CG", b"AG")]
    reads = []
    for i in range(n):
        L = 2 * int(rng.integers(5, 20))
        reads.append(np.tile(pats[i % 5], L // 2).tobytes() if i % 7 else H.random_read(rng, L).tobytes())
    stream = H.to_stream(reads)
    for kw in (dict(min_shannon=1.0), dict(min_shannon=1.0, min_len=20, max_reads=n // 2)):
        exp_bv, exp_cnt = oracle.filter_reads(*stream, **kw)
        bv, cnt = ctx.filter_reads(*stream, **kw)
        assert cnt == exp_cnt and np.array_equal(bv, exp_bv), kw
    assert exp_cnt["selected"] > 0


def test_filter_reads_long_reads(ctx):
    """reads far longer than a warp step and than the 16-bit partial counters of the fused kernel"""
    rng = np.random.default_rng(77)
    reads = [H.dirty(rng, H.random_read(rng, L), p_N=0.01).tobytes() for L in (70000, 1, 600000, 129, 128, 127, 5000)]
    reads.append(b"A" * 300000)
    stream = H.to_stream(reads)
    for kw in (dict(min_len=100, max_N=800, min_shannon=1.5), dict(max_N=6000), dict(min_shannon=0.5)):
        exp_bv, exp_cnt = oracle.filter_reads(*stream, **kw)
        bv, cnt = ctx.filter_reads(*stream, **kw)
        assert cnt == exp_cnt and np.array_equal(bv, exp_bv), kw


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 127, 128, 129, 4099, 1_000_003])
def test_bvop_and_popcount(ctx, n):
    import commet_b200 as cb
    rng = np.random.default_rng(n)
    a = rng.integers(0, 256, size=n // 8 + 1).astype(np.uint8)
    b = rng.integers(0, 256, size=n // 8 + 1).astype(np.uint8)
    for op, oop in ((cb.BV_AND, oracle.BV_AND), (cb.BV_OR, oracle.BV_OR), (cb.BV_ANDNOT, oracle.BV_ANDNOT),
                    (cb.BV_NOT, oracle.BV_NOT)):
        exp = oracle.bvop(oop, a, b)
        got = ctx.bvop(op, a, b)
        assert np.array_equal(got, exp)
        assert ctx.nb_one(got, n) == oracle.nb_one(exp, n)
    with pytest.raises(cb.CommetError):
        ctx.bvop(cb.BV_AND, a, np.zeros(a.size + 1, np.uint8))


def test_popcount_of_several_vectors_in_one_call(ctx):
    """commet_bv_popcount_batch_dev: nb_one (all n/8+1 bytes, clamped to n: boolean_vector.h:244-270) of vectors of
    different sizes with one read-back"""
    import torch
    rng = np.random.default_rng(12)
    sizes = [0, 1, 7, 8, 129, 4099, 1_000_003, 64]
    vecs = [rng.integers(0, 256, size=n // 8 + 1).astype(np.uint8) for n in sizes]
    vecs[3][:] = 0xFF                                    # 16 set bits in a vector of 8: the count is clamped
    dev = [torch.zeros((v.size + 15) // 16 * 16, dtype=torch.uint8, device="cuda:0") for v in vecs]
    for d, v in zip(dev, vecs):
        d[:v.size] = torch.as_tensor(v, device="cuda:0")
    torch.cuda.synchronize()
    got = ctx.nb_one_device_batch([d.data_ptr() for d in dev], sizes)
    assert got == [oracle.nb_one(v, n) for v, n in zip(vecs, sizes)]
    assert got[3] == 8


def test_bvop_known_answer_not_counts_padding(ctx):
    """NOT of a 2000/10000 vector reports 8008 (padding bits flipped, clamped popcount)."""
    import commet_b200 as cb
    n = 10000
    tags = np.zeros(n, np.uint8); tags[:2000] = 1
    bv = oracle.tags_to_bv(tags)
    out = ctx.bvop(cb.BV_NOT, bv)
    assert ctx.nb_one(out, n) == 8008 and out[-1] == 0xFF


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("seed", range(4))
def test_upload_async_streams_behave_like_uploaded_ones(ctx, seed, pinned):
    """commet_reads_upload_async: copies queued in call order, encode at first use -- same bits as the synchronous
    upload, whichever entry point touches the stream first (counts, chunk plan, insert, search, filter)."""
    import torch
    rng = np.random.default_rng(12000 + seed)
    k, t = int(rng.integers(9, 19)), 2
    ref = H.make_ref_set(rng, 2500, 30, 130, p_N=0.01)
    qry = H.make_query_set(rng, ref, 1200, 30, 130, p_N=0.01)
    rs, qs = H.to_stream(ref), H.to_stream(qry)
    if pinned:
        def pin(a, dt):
            tt = torch.from_numpy(a.view(dt).copy()).pin_memory()
            return tt, tt.numpy().view(a.dtype)
        keep_r, rb = pin(rs[0], np.uint8); keep_ro, ro = pin(rs[1], np.int64)
        keep_q, qb = pin(qs[0], np.uint8); keep_qo, qo = pin(qs[1], np.int64)
        rs, qs = (rb, ro), (qb, qo)
    idx = ctx.stage_async(*rs)
    q = ctx.stage_async(*qs)
    order = seed % 4
    if order == 0:
        assert np.array_equal(ctx.kmer_counts(idx, k), ctx.kmer_counts(ctx.stage(*rs), k))
    elif order == 1:
        assert ctx.kmer_total(idx, k) == int(ctx.kmer_counts(ctx.stage(*rs), k).sum())
    elif order == 2:
        bv, cnt = ctx.filter_reads_range(q, 0, len(qry), min_len=40, max_N=1, min_shannon=1.2)
        ebv, ecnt = oracle.filter_reads(*H.to_stream(qry), min_len=40, max_N=1, min_shannon=1.2)
        assert np.array_equal(bv, ebv) and cnt == ecnt
    exp_tags, exp = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(qry)])
    tags = torch.zeros((len(qry) // 8 + 1 + 3) // 4, dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    info = ctx.index_and_search_staged(k, t, idx, [q], [tags.data_ptr()])
    ctx.sync()
    got = tags.cpu().numpy().view(np.uint8)[:len(qry) // 8 + 1]
    assert np.array_equal(got, oracle.tags_to_bv(exp_tags[0]))
    assert info["chunks"] == exp["chunks"] and info["shared"] == exp["shared"] and info["searched"] == exp["searched"]
    idx.free()
    q.free()
