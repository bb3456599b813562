"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle needs minutes there):

C2 (2 sets x 10 M reads x 100 bp, k=33, t=2, 4 GiB DRAM-resident filter)
  * a Bloom filter has no false negatives: a set searched against its own index, and its reverse complement,
    are tagged completely; a second search over tagged reads scans nothing (file_manager.h:99)
  * the L2-blocked (sorted) insert and the direct RED.OR insert build the same 4 GiB, and every key the ORACLE
    computes for sampled reads is set at the bloom_filter.h bit position
  * sampled query reads: the oracle's search_reads over the downloaded GPU filter gives the GPU's tag bits
  * the host entry point (index set uploaded in parts) and the staged entry point give the same vector
  * chunking (index_reads' stop rule) at this size: self-search misses at most the reads lost at the boundaries
C5 (bit vectors of 1e9 reads; filter_reads over 10 M reads with constructed classes)
  * inclusion-exclusion, involution and annihilation identities of the vector operators, counts by construction
"""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

N, L, K, T = 10_000_000, 100, 33, 2


class _DevMem:
    """raw device memory as a torch tensor (CUDA array interface)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


@pytest.fixture(scope="module")
def env():
    import torch
    import commet_b200
    from commet_b200 import build
    import bench
    build.build_lib()
    ctx = commet_b200.Context(0)
    dev = torch.device("cuda", 0)
    ref, qry, offs = bench.make_sets_torch(N, L, seed=0, device=dev)
    torch.cuda.synchronize()
    e = dict(torch=torch, ctx=ctx, dev=dev, ref=ref, qry=qry, offs=offs, lib=commet_b200)
    yield e
    ctx.close()


def _tag_words(n):
    return (n // 8 + 1 + 3) // 4


def _staged_run(e, index_d, queries_d, maxk=None):
    """index_and_search on device-resident ASCII sets; returns (info, [u8 tag payload tensors])"""
    torch, ctx = e["torch"], e["ctx"]
    torch.cuda.synchronize()
    idx = ctx.stage_device(index_d.data_ptr(), e["offs"].data_ptr(), N, N * L)
    qs = [ctx.stage_device(q.data_ptr(), e["offs"].data_ptr(), N, N * L) for q in queries_d]
    tags = [torch.zeros(_tag_words(N), dtype=torch.int32, device=e["dev"]) for _ in queries_d]
    torch.cuda.synchronize()
    info = ctx.index_and_search_staged(K, T, idx, qs, [t.data_ptr() for t in tags], maxk)
    ctx.sync()
    return info, tags, idx, qs


def _revcomp(e, x):
    torch = e["torch"]
    comp = torch.zeros(256, dtype=torch.uint8, device=e["dev"])
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    return comp[x.view(N, L).flip(1).long()].reshape(-1).contiguous()


def test_c2_no_false_negatives_and_skip_tagged(env):
    e, torch, ctx = env, env["torch"], env["ctx"]
    rc = _revcomp(e, e["ref"])
    info, tags, idx, qs = _staged_run(e, e["ref"], [e["ref"], rc])
    assert info["chunks"] == 1 and info["indexed"] == N and info["kmers"] == N * (L - K + 1)
    assert info["searched"] == [N, N] and info["shared"] == [N, N]
    for t in tags:
        assert ctx.nb_one_device(t.data_ptr(), N) == N
        by = t.view(torch.uint8)[:N // 8 + 1]
        assert int(by[:N // 8].min()) == 0xFF and int(by[N // 8]) == (1 << (N % 8)) - 1      # padding bits stay 0
    # every read is tagged: a second search scans and finds nothing
    cnt = torch.zeros(4, dtype=torch.int64, device=e["dev"])
    torch.cuda.synchronize()
    ctx.search_reads_device(qs[0], K, T, tags[0].data_ptr(), cnt.data_ptr())
    ctx.sync()
    assert cnt.tolist()[:2] == [0, 0]
    idx.free()
    for q in qs:
        q.free()


def test_c2_filter_sorted_insert_equals_direct_and_oracle_keys(env):
    e, torch, ctx = env, env["torch"], env["ctx"]
    F = oracle.filter_bytes(K)
    torch.cuda.synchronize()
    idx = ctx.stage_device(e["ref"].data_ptr(), e["offs"].data_ptr(), N, N * L)
    ctx.binned_index(102)                 # second form of the L2-blocked insert (slabs, no histogram pass): the default
    ctx.index_reads(idx, K)
    ctx.sync()
    sorted_f = torch.as_tensor(_DevMem(ctx.filter_ptr, F), device=e["dev"]).clone()
    torch.cuda.synchronize()              # the copy runs on torch's stream: it must land before the context clears the filter
    try:
        for mode in (101, 0):             # first form (histogram + scatter + apply), then direct RED.OR
            ctx.binned_index(mode)
            ctx.index_reads(idx, K)
            ctx.sync()
            other_f = torch.as_tensor(_DevMem(ctx.filter_ptr, F), device=e["dev"])
            assert torch.equal(sorted_f, other_f), mode
            del other_f
    finally:
        ctx.binned_index(102)
    # at most one bit per key; a, b, c keys are uniform over 2^33 positions (~6.54e8 distinct each), d = a|b is not
    ones = ctx.nb_one_device(ctx.filter_ptr, 8 * (F - 1))
    kmers = N * (L - K + 1)
    assert 3 * 640_000_000 < ones <= 4 * kmers
    # the oracle's keys of sampled reads sit at BloomFilter::feed's positions (bloom_filter.h:112-118)
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(N, size=300, replace=False))
    seqs = e["ref"].view(N, L)[torch.as_tensor(pick, device=e["dev"])].cpu().numpy()
    byte_idx, masks = [], []
    for s in seqs:
        ks, size = oracle.keys(s.tobytes(), K)
        full = ks[size >= K]                                   # one row of (a, b, c, d) per k-mer
        assert len(full) == L - K + 1
        for j in range(4):
            key = full[:, j]
            byte_idx.append((key >> np.uint64(1)).astype(np.int64))
            masks.append(np.where(key & np.uint64(1), 8 >> j, 128 >> j).astype(np.uint8))
    byte_idx, masks = np.concatenate(byte_idx), np.concatenate(masks)
    got = sorted_f[torch.as_tensor(byte_idx, device=e["dev"])].cpu().numpy()
    assert np.all(got & masks == masks)
    idx.free()


def test_c2_sampled_reads_match_oracle_search_and_host_path(env):
    e, torch, ctx = env, env["torch"], env["ctx"]
    info, tags, idx, qs = _staged_run(e, e["ref"], [e["qry"]])
    shared = info["shared"][0]
    assert 0.4 < shared / N < 0.6 and info["searched"] == [N]
    dev_bv = tags[0].view(torch.uint8)[:N // 8 + 1].cpu().numpy()
    assert ctx.nb_one(dev_bv, N) == shared == int(np.unpackbits(dev_bv, bitorder="little")[:N].sum())
    # the oracle's search_reads over the GPU-built 4 GiB filter, on sampled query reads
    filt = ctx.filter_download(K)
    rng = np.random.default_rng(9)
    pick = np.sort(rng.choice(N, size=4000, replace=False))
    seqs = e["qry"].view(N, L)[torch.as_tensor(pick, device=e["dev"])].cpu().numpy()
    offs = np.arange(len(pick) + 1, dtype=np.uint64) * L
    otags = np.zeros(len(pick), dtype=np.uint8)
    oracle.search(filt, K, T, seqs.reshape(-1), offs, otags)
    gtags = np.unpackbits(dev_bv, bitorder="little")[pick]
    assert np.array_equal(otags, gtags)
    assert 0.3 < otags.mean() < 0.7
    del filt
    idx.free()
    qs[0].free()
    # the host entry point (pageable buffers, index set uploaded in parts) gives the same vector
    ref_h, qry_h = e["ref"].cpu().numpy(), e["qry"].cpu().numpy()
    offs_h = np.arange(N + 1, dtype=np.uint64) * L
    htags, hinfo = ctx.index_and_search(K, T, (ref_h, offs_h), [(qry_h, offs_h)])
    assert hinfo["parts"] > 1 and hinfo["shared"] == [shared] and hinfo["chunks"] == 1
    assert np.array_equal(htags[0], dev_bv)


def test_c2_chunked_self_search(env):
    e, torch, ctx = env, env["torch"], env["ctx"]
    kmers = N * (L - K + 1)
    maxk = kmers // 4 + 1                                   # 4 chunks, 3 reads fetched and lost (index_reads.h:48-60)
    info, tags, idx, qs = _staged_run(e, e["ref"], [e["ref"]], maxk=maxk)
    assert info["chunks"] == 4 and info["indexed"] == N - 3
    assert N - 3 <= info["shared"][0] <= N
    assert ctx.nb_one_device(tags[0].data_ptr(), N) == info["shared"][0]
    # reads tagged by an earlier chunk are not scanned again: the last chunk sees about a quarter of the set
    assert info["searched"][0] < 0.3 * N
    idx.free()
    qs[0].free()


def test_c5_vector_identities_at_1e9_bits(env):
    e, torch, ctx, lib = env, env["torch"], env["ctx"], env["lib"]
    n = 1_000_000_000
    nb = n // 8 + 1
    g = torch.Generator(device=e["dev"])
    g.manual_seed(7)
    words = (nb + 7) // 8
    a = torch.randint(-2**63, 2**63 - 1, (words,), generator=g, device=e["dev"], dtype=torch.int64).view(torch.uint8)
    b = torch.randint(-2**63, 2**63 - 1, (words,), generator=g, device=e["dev"], dtype=torch.int64).view(torch.uint8)
    out = torch.empty_like(a)
    torch.cuda.synchronize()

    def ones(x):
        return ctx.nb_one_device(x.data_ptr(), n)

    def op(code, x, y, dst):
        ctx.bvop_device(code, x.data_ptr(), y.data_ptr() if y is not None else None, dst.data_ptr(), nb)
        ctx.sync()

    # nb_one counts ALL n/8+1 payload bytes, padding bits included, then clamps to n (boolean_vector.h:244-270);
    # here n % 8 == 0, so the last byte is pure (random) padding
    na, nbb = ones(a), ones(b)
    assert abs(na - n / 2) < 5e5 and abs(nbb - n / 2) < 5e5           # fair bits
    ref = int(sum(int(torch.count_nonzero((a[:nb] >> s) & 1)) for s in range(8)))     # independent count (torch)
    assert na == ref
    op(lib.BV_AND, a, b, out)
    n_and = ones(out)
    assert torch.equal(out[:nb], a[:nb] & b[:nb])
    op(lib.BV_OR, a, b, out)
    n_or = ones(out)
    assert n_and + n_or == na + nbb                                    # inclusion-exclusion
    op(lib.BV_ANDNOT, a, b, out)
    assert ones(out) == na - n_and
    op(lib.BV_ANDNOT, a, a, out)
    assert ones(out) == 0 and int(out[:nb].max()) == 0
    op(lib.BV_NOT, a, None, out)
    assert ones(out) == 8 * nb - na                                    # NOT flips the padding bits too
    tmp = torch.empty_like(a)
    op(lib.BV_NOT, out, None, tmp)
    assert torch.equal(tmp[:nb], a[:nb])                               # involution over all payload bytes


def test_c5_filter_reads_constructed_classes_10m(env):
    e, torch, ctx = env, env["torch"], env["ctx"]
    dev = e["dev"]
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    n = N
    reads = e["ref"].view(N, L).clone()
    cls = torch.randint(0, 20, (n,), generator=g, device=dev)
    # class 0: homopolymer (Shannon 0); class 1: dinucleotide repeat (Shannon 1.0 < 1.5); class 2: 3..10 N (> max_N=2);
    # class 3: 1..2 N (kept); others: random reads (Shannon ~2)
    reads[cls == 0] = ord("A")
    di = torch.tensor(list(b"AC" * (L // 2)), dtype=torch.uint8, device=dev)
    reads[cls == 1] = di
    pos = torch.arange(L, device=dev)[None, :]
    n_many = torch.randint(3, 11, (n,), generator=g, device=dev)[:, None]
    n_few = torch.randint(1, 3, (n,), generator=g, device=dev)[:, None]
    reads = torch.where((cls == 2)[:, None] & (pos < n_many), torch.full_like(reads, ord("N")), reads)
    reads = torch.where((cls == 3)[:, None] & (pos < n_few), torch.full_like(reads, ord("N")), reads)
    reads = reads.reshape(-1).contiguous()
    bv = torch.zeros(_tag_words(n), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    cnt = ctx.filter_reads_device(reads.data_ptr(), e["offs"].data_ptr(), n, bv.data_ptr(), min_len=66, max_N=2,
                                  min_shannon=1.5)
    ctx.sync()
    c = [int((cls == i).sum()) for i in range(4)]
    assert cnt["rm_length"] == 0 and cnt["rm_N"] == c[2] and cnt["rm_shannon"] == c[0] + c[1]
    assert cnt["selected"] == n - c[0] - c[1] - c[2] == ctx.nb_one_device(bv.data_ptr(), n)
    bits = bv.view(torch.uint8)[:n // 8 + 1]
    sel = torch.stack([(bits >> s) & 1 for s in range(8)], dim=1).reshape(-1)[:n].bool()
    assert torch.equal(sel, cls >= 3)
    # and the oracle agrees on a sample of every class
    pick = torch.cat([torch.nonzero(cls == i)[:50, 0] for i in range(5)]).sort().values
    seqs = reads.view(n, L)[pick].cpu().numpy()
    ebv, _ = oracle.filter_reads(seqs.reshape(-1), np.arange(len(pick) + 1, dtype=np.uint64) * L, min_len=66, max_N=2,
                                 min_shannon=1.5)
    assert np.array_equal(np.unpackbits(ebv, bitorder="little")[:len(pick)].astype(bool), sel[pick].cpu().numpy())
    # a min-length above the read length removes everything by length (tested first: filter_reads.cpp:189-191)
    cnt = ctx.filter_reads_device(reads.data_ptr(), e["offs"].data_ptr(), n, bv.data_ptr(), min_len=L + 1, max_N=2,
                                  min_shannon=1.5)
    assert cnt["rm_length"] == n and cnt["selected"] == 0


def test_c4_shape_many_chunks_eight_query_sets(env):
    """C4 (one reference set of many chunks against 8 query sets) scaled down with a smaller k so that the stop rule
    still triggers: 300 k reference reads x 150 bp = 3.8e7 k-mers = 10 chunks of max_kmer(25) = 3 906 250
    with 9 reads fetched and lost, 8 query sets of 10 k reads -- every vector and counter against the oracle (the
    filter is 11 % full per key, so false-positive tags occur and must match too)."""
    ctx = env["ctx"]
    rng = np.random.default_rng(44)
    k, t, Lr, n_ref, n_q = 25, 2, 150, 300_000, 10_000
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    ref = acgt[rng.integers(0, 4, size=(n_ref, Lr))]
    ref[rng.random(n_ref) < 0.01, 70] = ord("N")            # some reads carry fewer k-mers: uneven chunk lengths
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    queries = []
    for s in range(8):
        q = acgt[rng.integers(0, 4, size=(n_q, Lr))]
        cp = rng.random(n_q) < 0.5                          # half of the reads are copies of reference reads ...
        src = ref[rng.integers(0, n_ref, size=n_q)]
        rc = rng.random(n_q) < 0.5                          # ... half of those reverse-complemented, 1 % substitutions
        src = np.where(rc[:, None], comp[src[:, ::-1]], src)
        mut = rng.random((n_q, Lr)) < 0.01
        src = np.where(mut, acgt[rng.integers(0, 4, size=(n_q, Lr))], src)
        q = np.where(cp[:, None], src, q).astype(np.uint8)
        queries.append((q.reshape(-1), np.arange(n_q + 1, dtype=np.uint64) * Lr))
    istream = (ref.reshape(-1), np.arange(n_ref + 1, dtype=np.uint64) * Lr)
    exp_tags, exp = oracle.index_and_search(k, t, istream, queries)
    assert exp["chunks"] == 10 and exp["indexed"] == n_ref - 9
    tags, info = ctx.index_and_search(k, t, istream, queries)
    assert info["chunks"] == exp["chunks"] and info["indexed"] == exp["indexed"] and info["kmers"] == exp["kmers"]
    assert info["searched"] == exp["searched"] and info["shared"] == exp["shared"]
    for s in range(8):
        assert np.array_equal(tags[s], oracle.tags_to_bv(exp_tags[s])), s
        assert 0.4 < info["shared"][s] / n_q < 0.7


def _write_fasta_fixed(path, bases_2d):
    """>%08d headers, one sequence line per read, written as one numpy array (10 M reads: ~1.1 GB, seconds)"""
    n, length = bases_2d.shape
    rows = np.empty((n, 1 + 8 + 1 + length + 1), dtype=np.uint8)
    rows[:, 0] = ord(">")
    idx = np.arange(n, dtype=np.int64)
    for d in range(8):
        rows[:, 1 + d] = (idx // 10 ** (7 - d)) % 10 + 48
    rows[:, 9] = 10
    rows[:, 10:10 + length] = bases_2d
    rows[:, -1] = 10
    rows.tofile(path)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref did not travel")
def test_c2_whole_vector_equals_the_reference_binary(env, tmp_path):
    """C2 at full size through the executables: `commet_b200/bin/index_and_search` on the two 10 M x 100 bp FASTA files
    against the reference's own index_and_search (oracle/_ref) on the same files -- every bit of the 10 M-read vector.
    The reference is single-threaded (about 3 minutes of indexing + 2 of search here), so it is run as several
    processes over contiguous parts of the query file, each indexing the whole reference file; part sizes are multiples
    of 8 reads, so each part's payload is a byte range of the whole vector."""
    import os
    import subprocess
    from commet_b200 import build
    build.build_tools()
    e, torch = env, env["torch"]
    ref = e["ref"].view(N, L).cpu().numpy()
    qry = e["qry"].view(N, L).cpu().numpy()
    _write_fasta_fixed(tmp_path / "ref.fa", ref)
    (tmp_path / "ref.txt").write_text("ref:ref.fa\n")
    procs = max(1, min(8, (os.cpu_count() or 2) // 2))
    per = ((N + procs - 1) // procs + 7) // 8 * 8
    parts = [(s, min(N, s + per)) for s in range(0, N, per)]
    cmds = []
    for i, (s, t) in enumerate(parts):
        _write_fasta_fixed(tmp_path / f"q{i}.fa", qry[s:t])
        (tmp_path / f"q{i}.txt").write_text(f"q{i}:q{i}.fa\n")
        cmds.append([str(oracle.REF_DIR / "index_and_search"), "-i", "ref.txt", "-s", f"q{i}.txt", "-o", f"ref_out{i}", "-l",
                     f"ref_out{i}", "-k", str(K), "-t", str(T)])
    running = [subprocess.Popen(c, cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    # meanwhile: the drop-in tool on the whole query file
    _write_fasta_fixed(tmp_path / "qry.fa", qry)
    (tmp_path / "qry.txt").write_text("qry:qry.fa\n")
    r = subprocess.run([str(build.BIN / "index_and_search"), "-i", "ref.txt", "-s", "qry.txt", "-o", "gpu_out", "-l", "gpu_out",
                        "-k", str(K), "-t", str(T)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    comment, n, mine = oracle.read_bv_file(tmp_path / "gpu_out" / "qry.fa_in_ref.bv")
    assert n == N and comment == b"qry.fa in ref"
    assert all(p.wait(timeout=1500) == 0 for p in running)
    shared = 0
    for i, (s, t) in enumerate(parts):
        _, n_i, theirs = oracle.read_bv_file(tmp_path / f"ref_out{i}" / f"q{i}.fa_in_ref.bv")
        assert n_i == t - s
        whole = (t - s) // 8
        assert np.array_equal(mine[s // 8:s // 8 + whole], theirs[:whole]), f"part {i}: payload differs from the reference's"
        if (t - s) % 8:                                          # only the last part can end inside a byte
            assert mine[s // 8 + whole] == theirs[whole]
        shared += int(np.unpackbits(theirs, bitorder="little")[:t - s].sum())
    assert 0.4 < shared / N < 0.6
    log = (tmp_path / "gpu_out" / "qry_in_ref.log").read_text()
    assert f"[indexed {N}, searched {N}, shared {shared}]" in log
