"""Pin the C oracle against the reference's own binaries (oracle/_ref, compiled
from /root/reference by oracle/Makefile).  Skipped when oracle/_ref is absent.
CPU only; a few seconds."""
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


def _write_set(rng, tmp, name, reads_per_file, with_bv):
    """Write the files of one set in random formats; returns (fof item list, valid stream, per-file info)."""
    items, stream, info = [], [], []
    for fi, reads in enumerate(reads_per_file):
        kind = int(rng.integers(0, 5))
        path = tmp / f"{name}_{fi}"
        if kind == 0:
            H.write_fasta(path.with_suffix(".fa"), reads); path = path.with_suffix(".fa")
        elif kind == 1:
            H.write_fasta(path.with_suffix(".fa"), reads, width=int(rng.integers(7, 40)),
                          final_newline=bool(rng.integers(0, 2)), blank_every=int(rng.integers(0, 4)))
            path = path.with_suffix(".fa")
        elif kind == 2:
            H.write_fastq(path.with_suffix(".fq"), reads, final_newline=bool(rng.integers(0, 2)))
            path = path.with_suffix(".fq")
        elif kind == 3:
            H.write_fasta(path.with_suffix(".fa.gz"), reads, gz=True); path = path.with_suffix(".fa.gz")
        else:
            H.write_fastq(path.with_suffix(".fq.gz"), reads, gz=True); path = path.with_suffix(".fq.gz")
        assert oracle.parse_reads(path) == reads
        n = len(reads)
        if with_bv:
            valid = (rng.random(n) < 0.7).astype(np.uint8)
            if valid.sum() == 0:
                valid[int(rng.integers(0, n))] = 1
            bvp = tmp / f"{name}_{fi}.in.bv"
            oracle.write_bv_file(bvp, b"input", n, oracle.tags_to_bv(valid))
            items.append(f"{path},{bvp}")
        else:
            valid = np.ones(n, dtype=np.uint8)
            items.append(str(path))
        stream += [r for r, v in zip(reads, valid) if v]
        info.append((path, n, valid))
    return items, stream, info


@pytest.mark.parametrize("seed", range(40))
def test_index_and_search_fuzz(tmp_path, seed):
    _index_and_search_case(tmp_path, seed, None)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 22, 24, 27, 29])
def test_index_and_search_small_and_large_k(tmp_path, k):
    """the k the fuzz above does not draw: k <= 3 (max_kmer = 1e9 / 2^(33-k) truncates to 0: nothing is ever indexed and
    every call of index_reads loses a read, index_reads.h:48-49,60), single-byte filters, and filters of 2 MiB to 256 MiB"""
    _index_and_search_case(tmp_path, 500 + k, k)


def _index_and_search_case(tmp_path, seed, k_forced):
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.integers(8, 21))
    if k_forced is not None:
        k = k_forced
    t = int(rng.integers(0, 4))
    L = int(rng.integers(k, 4 * k))
    dirt = dict(p_N=float(rng.choice([0, 0.02])), p_lower=float(rng.choice([0, 0.3])),
                p_other=float(rng.choice([0, 0.01])))
    n_idx_files = int(rng.integers(1, 3))
    # enough k-mers for several chunks at small k
    maxk = oracle.max_kmer(k)
    n_ref = int(min(400, max(20, 3 * maxk // max(1, (L - k + 1)) + 5)))
    ref_files = [H.make_ref_set(rng, n_ref, max(1, L - 10), L + 10, **dirt) for _ in range(n_idx_files)]
    all_ref = [r for f in ref_files for r in f]
    iitems, istream, _ = _write_set(rng, tmp_path, "idx", ref_files, with_bv=bool(rng.integers(0, 2)))
    (tmp_path / "index.txt").write_text("refset:" + ";".join(iitems) + "\n")

    n_sets = int(rng.integers(1, 3))
    qinfo, lines = [], []
    for s in range(n_sets):
        nf = int(rng.integers(1, 3))
        files = [H.make_query_set(rng, all_ref, int(rng.integers(5, 120)), max(1, L - 10), L + 10, **dirt)
                 for _ in range(nf)]
        items, stream, info = _write_set(rng, tmp_path, f"q{s}", files, with_bv=bool(rng.integers(0, 2)))
        lines.append(f"Q{s}:" + ";".join(items))
        qinfo.append((f"Q{s}", stream, info))
    (tmp_path / "query.txt").write_text("\n".join(lines) + "\n")

    out = tmp_path / "out"
    r = subprocess.run([str(oracle.REF_DIR / "index_and_search"), "-i", str(tmp_path / "index.txt"),
                        "-s", str(tmp_path / "query.txt"), "-o", str(out), "-l", str(out),
                        "-k", str(k), "-t", str(t)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    tags, st = oracle.index_and_search(k, t, H.to_stream(istream), [H.to_stream(q[1]) for q in qinfo])
    for s, (name, stream, info) in enumerate(qinfo):
        pos = 0
        for path, n, valid in info:
            comment, nb, payload = oracle.read_bv_file(out / f"{path.name}_in_refset.bv")
            assert nb == n and comment == f"{path} in refset".encode()
            nv = int(valid.sum())
            exp = np.zeros(n, dtype=np.uint8)
            exp[valid.astype(bool)] = tags[s][pos:pos + nv]
            pos += nv
            assert np.array_equal(payload, oracle.tags_to_bv(exp)), (seed, k, t, path)
        log = (out / f"{name}_in_refset.log").read_text()
        m = re.search(r"\[indexed (\d+), searched (\d+), shared (\d+)\]", log)
        assert [int(x) for x in m.groups()] == [st["indexed"], st["searched"][s], st["shared"][s]]


@pytest.mark.parametrize("seed", range(25))
def test_filter_reads_fuzz(tmp_path, seed):
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(1, 300))
    reads = []
    for _ in range(n):
        L = int(rng.integers(1, 160))
        kind = rng.random()
        if kind < 0.15:
            r = np.full(L, ord(rng.choice(list("ACGTacgtN"))), dtype=np.uint8)
        elif kind < 0.3:
            r = np.tile(np.frombuffer(b"AC", dtype=np.uint8), L)[:L]
        else:
            r = H.dirty(rng, H.random_read(rng, L), p_N=float(rng.choice([0, 0.05])),
                        p_lower=float(rng.choice([0, 0.4])), p_other=float(rng.choice([0, 0.02])))
        reads.append(r.tobytes())
    path = tmp_path / "in.fa"
    if rng.random() < 0.5:
        H.write_fasta(path, reads)
    else:
        path = tmp_path / "in.fq"
        H.write_fastq(path, reads)
    args, kw = [], {}
    if rng.random() < 0.8:
        kw["min_len"] = int(rng.integers(0, 120)); args += ["-l", str(kw["min_len"])]
    if rng.random() < 0.6:
        kw["max_N"] = int(rng.integers(0, 6)); args += ["-n", str(kw["max_N"])]
    if rng.random() < 0.8:
        e = round(float(rng.uniform(0, 2.1)), int(rng.integers(0, 7)))
        kw["min_shannon"] = float(np.float32(float(str(e)))); args += ["-e", str(e)]
    if rng.random() < 0.4:
        kw["max_reads"] = int(rng.integers(0, n + 3)); args += ["-m", str(float(kw["max_reads"]))]
    outp = tmp_path / "o.bv"
    r = subprocess.run([str(oracle.REF_DIR / "filter_reads"), str(path), *args, "-o", str(outp)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _, nb, payload = oracle.read_bv_file(outp)
    bv, cnt = oracle.filter_reads(*H.to_stream(reads), **kw)
    assert nb == n
    assert np.array_equal(payload, bv), (seed, args)
    nums = [int(x) for x in re.findall(r"(\d+) reads removed", r.stdout)]
    sel = int(re.search(r"selected reads = (\d+)", r.stdout).group(1))
    assert nums == [cnt["rm_length"], cnt["rm_N"], cnt["rm_shannon"]] and sel == cnt["selected"]


@pytest.mark.parametrize("seed", range(8))
def test_bvop_fuzz(tmp_path, seed):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([0, 1, 7, 8, 9, 63, 64, 1000, 4099]))
    a = rng.integers(0, 256, size=n // 8 + 1).astype(np.uint8)
    b = rng.integers(0, 256, size=n // 8 + 1).astype(np.uint8)
    # printable payloads are not required: the reader copies raw bytes; but the header scan stops
    # at the first '#', so keep '#' out of the comments only.
    oracle.write_bv_file(tmp_path / "a.bv", b"A", n, a)
    oracle.write_bv_file(tmp_path / "b.bv", b"B", n, b)
    for flag, op in (("-a", oracle.BV_AND), ("-o", oracle.BV_OR), ("-d", oracle.BV_ANDNOT), ("-n", oracle.BV_NOT)):
        cmd = [str(oracle.REF_DIR / "bvop"), str(tmp_path / "a.bv"), flag]
        if flag != "-n":
            cmd.append(str(tmp_path / "b.bv"))
        r = subprocess.run(cmd + ["-p", str(tmp_path / "c.bv"), "-i"], capture_output=True, text=True)
        assert r.returncode == 0
        _, nb, payload = oracle.read_bv_file(tmp_path / "c.bv")
        exp = oracle.bvop(op, a, b)
        assert nb == n and np.array_equal(payload, exp)
        line = r.stdout.split("\n")[-2]
        assert line == f"  {oracle.nb_one(exp, n)} / {n} reads selected"
