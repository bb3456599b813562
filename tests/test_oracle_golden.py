"""CPU: the C oracle against the committed golden vectors (produced by the unmodified reference,
tests/golden/make_golden.py).  This is what pins the oracle on a box without /root/reference."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from tests.golden import fixtures

GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    w = tmp_path_factory.mktemp("golden")
    fixtures.materialize(w)
    return w


def bv_file_sha(comment: bytes, n: int, payload: np.ndarray) -> str:
    return hashlib.sha256(comment + b"\n#" + str(n).encode() + b"\n" + bytes(payload)).hexdigest()


@pytest.mark.parametrize("k", [20, 22])
def test_chunk_boundary_known_answer(work, k):
    """A.fa in A.fa loses one read per chunk boundary (k=20: 8 reads, k=22: 2 reads)."""
    reads = oracle.parse_reads(work / "ABCDE_bench" / "A.fa")
    stream = oracle.to_stream(reads)
    tags, info = oracle.index_and_search(k, 2, stream, [stream])
    g = GOLDEN["chunk_boundary_A_in_A"][str(k)]
    assert [info["indexed"], info["searched"][0], info["shared"][0]] == g["counters"]
    assert bv_file_sha(b"ABCDE_bench/A.fa in A", len(reads), oracle.tags_to_bv(tags[0])) == g["bv_sha256"]
    untagged = np.flatnonzero(tags[0] == 0).tolist()
    if k == 20:
        assert untagged == [1342, 2685, 4028, 5371, 6714, 8057, 9400, 10743]
    else:
        assert untagged == [5487, 10975]


def test_filter_reads_golden(work):
    """filter_reads -l 100 -e 1.9 -n 0 -m 9000 on A.fa (Commet.py run 'abcde_3sets_k21_filtered')."""
    reads = oracle.parse_reads(work / "ABCDE_bench" / "A.fa")
    bv, cnt = oracle.filter_reads(*oracle.to_stream(reads), min_len=100, max_N=0, min_shannon=float(np.float32(1.9)),
                                  max_reads=9000)
    comment = (b"----------------\nReference file\n  A.fa\nFilter Options\n  min read size     : 100\n"
               b"  max number of N   : 0\n  min shannon index : 1.9\n")
    assert bv_file_sha(comment, len(reads), bv) == GOLDEN["abcde_3sets_k21_filtered"]["bv"]["A.fa.bv"]
    assert cnt["selected"] == 9000


def test_three_pass_refinement_golden(work):
    """The A-in-B / B-in-(A in B) / A-in-(B in (A in B)) rounds of Commet.py:186-240 for sets 1 and 3 of
    ABCDE_bench at k=32, replayed with the oracle; compares the final .bv files."""
    g = GOLDEN["abcde_3sets_k32"]["bv"]
    A = oracle.parse_reads(work / "ABCDE_bench" / "A.fa")
    D = oracle.parse_reads(work / "ABCDE_bench" / "D.fa")
    sA, sD = oracle.to_stream(A), oracle.to_stream(D)
    # round a: D in A (raw)
    (d_in_a,), _ = oracle.index_and_search(32, 2, sA, [sD])
    # round b: index D restricted to d_in_a, search A
    Dsel = [r for r, tg in zip(D, d_in_a) if tg]
    (a_in_d,), _ = oracle.index_and_search(32, 2, oracle.to_stream(Dsel), [sA])
    assert bv_file_sha(b"ABCDE_bench/A.fa in set3", len(A), oracle.tags_to_bv(a_in_d)) == g["A.fa_in_set3.bv"]
    # round c: index A restricted to a_in_d, search D
    Asel = [r for r, tg in zip(A, a_in_d) if tg]
    (d_in_a2,), _ = oracle.index_and_search(32, 2, oracle.to_stream(Asel), [sD])
    assert bv_file_sha(b"ABCDE_bench/D.fa in set1", len(D), oracle.tags_to_bv(d_in_a2)) == g["D.fa_in_set1.bv"]


def test_bvop_golden(work):
    n = 10000
    # NOT of the 2000/10000 vector reports 8008: padding bits are flipped too and the count is clamped
    assert "  8008 / 10000 reads selected" in GOLDEN["bvop"]["not"]["stdout"]
    tags = np.zeros(n, np.uint8); tags[:2000] = 1
    out = oracle.bvop(oracle.BV_NOT, oracle.tags_to_bv(tags))
    assert oracle.nb_one(out, n) == 8008 and out[-1] == 0xFF
