"""extract_reads drop-in (commet_b200/bin/extract_reads) against the reference's own binary (oracle/_ref/extract_reads,
compiled from /root/reference/src/extract_reads.cpp) on seeded files: FASTA / multi-line FASTA with blank lines /
FASTQ, plain and gzip, with and without a final newline, random and degenerate vectors.  Host I/O only: no GPU."""
import gzip
import subprocess

import numpy as np
import pytest

from commet_b200 import build
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.skipif(not (oracle.REF_DIR / "extract_reads").exists(), reason="oracle/_ref/extract_reads not built")


@pytest.fixture(scope="module")
def tool():
    build.build_tools()
    return build.BIN / "extract_reads"


def _write(rng, tmp, kind, reads):
    p = tmp / "in"
    if kind == "fa":
        return H.write_fasta(p.with_suffix(".fa"), reads, final_newline=bool(rng.integers(0, 2)))
    if kind == "fa_multi":
        return H.write_fasta(p.with_suffix(".fa"), reads, width=int(rng.integers(5, 40)),
                             final_newline=bool(rng.integers(0, 2)), blank_every=int(rng.integers(0, 4)))
    if kind == "fq":
        return H.write_fastq(p.with_suffix(".fq"), reads, final_newline=bool(rng.integers(0, 2)))
    if kind == "fa_gz":
        return H.write_fasta(p.with_suffix(".fa.gz"), reads, gz=True, width=int(rng.choice([0, 13])) or None)
    if kind == "fa_gz_blank":
        return H.write_fasta(p.with_suffix(".fa.gz"), reads, gz=True, width=11, blank_every=2,
                             final_newline=bool(rng.integers(0, 2)))
    if kind == "fq_gz":
        return H.write_fastq(p.with_suffix(".fq.gz"), reads, gz=True, final_newline=bool(rng.integers(0, 2)))
    raise ValueError(kind)


def _run(exe, args, cwd):
    return subprocess.run([str(exe), *map(str, args)], cwd=cwd, capture_output=True, timeout=60)


KINDS = ["fa", "fa_multi", "fq", "fa_gz", "fa_gz_blank", "fq_gz"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", range(6))
def test_extract_reads_vs_reference_binary(tool, tmp_path, kind, seed):
    rng = np.random.default_rng(7000 + 31 * seed + KINDS.index(kind))
    n = int(rng.integers(1, 200))
    reads = H.make_ref_set(rng, n, 1, 90, p_N=0.02, p_lower=0.1)
    path = _write(rng, tmp_path, kind, reads)
    mode = seed % 3
    if mode == 0:
        sel = (rng.random(n) < 0.5).astype(np.uint8)
    elif mode == 1:
        sel = np.ones(n, dtype=np.uint8)
    else:
        sel = np.zeros(n, dtype=np.uint8)
        sel[int(rng.integers(0, n))] = 1
    bvp = tmp_path / "sel.bv"
    oracle.write_bv_file(bvp, b"selection", n, oracle.tags_to_bv(sel))
    gz = kind.endswith("gz") or "_gz" in kind
    a = _run(tool, [path, bvp, "-o", tmp_path / "mine.out"], tmp_path)
    b = _run(oracle.REF_DIR / "extract_reads", [path, bvp, "-o", tmp_path / "ref.out"], tmp_path)
    assert a.returncode == b.returncode == 0, (a.stderr, b.stderr)
    mine, ref = (tmp_path / "mine.out").read_bytes(), (tmp_path / "ref.out").read_bytes()
    if gz:
        mine, ref = gzip.decompress(mine), gzip.decompress(ref)
    assert mine == ref
    # the selected sequences survive the round trip through our own parser
    if sel.any():
        back = tmp_path / ("back.fq" if "fq" in kind else "back.fa")
        back.write_bytes(mine)
        assert oracle.parse_reads(back) == [r for r, s in zip(reads, sel) if s]
    if not gz:          # stdout mode exists only for plain inputs (extract_reads.cpp:147-150)
        a = _run(tool, [path, bvp], tmp_path)
        b = _run(oracle.REF_DIR / "extract_reads", [path, bvp], tmp_path)
        assert a.returncode == b.returncode == 0 and a.stdout == b.stdout


def test_extract_reads_errors_match_reference(tool, tmp_path):
    reads = [b"ACGTACGT", b"TTTTGGGG", b"ACACACAC"]
    fa = H.write_fasta(tmp_path / "x.fa", reads)
    gzp = H.write_fasta(tmp_path / "x.fa.gz", reads, gz=True)
    bad = tmp_path / "bad.bv"
    oracle.write_bv_file(bad, b"c", 5, oracle.tags_to_bv(np.ones(5, dtype=np.uint8)))
    good = tmp_path / "good.bv"
    oracle.write_bv_file(good, b"c", 3, oracle.tags_to_bv(np.ones(3, dtype=np.uint8)))
    for args in ([fa, bad], [gzp, good], ["-v"], [fa, good, "-z"]):
        a = _run(tool, args, tmp_path)
        b = _run(oracle.REF_DIR / "extract_reads", args, tmp_path)
        assert a.returncode == b.returncode, args
        assert a.stdout == b.stdout, args
        assert a.stderr == b.stderr, args


def test_extract_after_index_and_search_selection(tool, tmp_path):
    """The tool's usual place in the flow: the reads a .bv selects are the ones written out, in file order."""
    rng = np.random.default_rng(5)
    reads = H.make_ref_set(rng, 500, 40, 80)
    fq = H.write_fastq(tmp_path / "s.fq", reads)
    sel = (rng.random(500) < 0.3).astype(np.uint8)
    bvp = tmp_path / "s.bv"
    oracle.write_bv_file(bvp, b"s.fq in other", 500, oracle.tags_to_bv(sel))
    r = _run(tool, [fq, bvp, "-o", tmp_path / "o.fq"], tmp_path)
    assert r.returncode == 0
    assert oracle.parse_reads(tmp_path / "o.fq") == [x for x, s in zip(reads, sel) if s]
