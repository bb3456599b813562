"""GPU: the drop-in executables (commet_b200/bin) against the golden outputs of the unmodified reference
and, when oracle/_ref travelled with the snapshot, live against the reference binaries on fuzzed files."""
import hashlib
import json
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from tests import commet_flow
from tests import helpers as H
from tests.golden import fixtures

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())


@pytest.fixture(scope="module")
def bin_dir():
    from commet_b200 import build
    build.build_all()
    return build.BIN


def sha(p):
    return hashlib.sha256(Path(p).read_bytes()).hexdigest()


def check_flow(work, res, g):
    for n in ("plain", "percentage", "normalized"):
        assert res[n] == g["csv"][n]
        assert hashlib.sha256(res[n].encode()).hexdigest() == g["csv_sha256"][n]
    got = {p.name: sha(p) for p in (work / "output_commet").glob("*.bv")}
    assert got == g["bv"]


@pytest.mark.parametrize("case,config,kw", [
    ("abcde_3sets_k32", "ABCDE_bench/sets_config.txt", dict(k=32)),
    ("abcde_5sets_k32", "ABCDE_bench/five_sets.txt", dict(k=32)),
    ("dissymmetry_k33", "test_dissymmetry/fof.txt", dict(k=33)),
    ("abcde_3sets_k21_filtered", "ABCDE_bench/sets_config.txt", dict(k=21, t=3, l=100, e=1.9, n=0, m=9000)),
])
def test_commet_flow_bit_exact(bin_dir, tmp_path, case, config, kw):
    """Whole Commet.py flow (filter_reads -> N^2-1 index_and_search -> bvop -i matrices): every .bv file and the
    three CSV matrices byte-identical to the reference's."""
    fixtures.materialize(tmp_path)
    res = commet_flow.run(config, bin_dir, tmp_path, **kw)
    check_flow(tmp_path, res, GOLDEN[case])


@pytest.mark.skipif(not (oracle.REF_DIR / "Commet.py").exists(), reason="oracle/_ref/Commet.py did not travel")
@pytest.mark.parametrize("case,config,opts", [
    ("abcde_3sets_k32", "ABCDE_bench/sets_config.txt", ["-k", "32"]),
    ("abcde_5sets_k32", "ABCDE_bench/five_sets.txt", ["-k", "32"]),
    ("dissymmetry_k33", "test_dissymmetry/fof.txt", ["-k", "33"]),
    ("abcde_3sets_k21_filtered", "ABCDE_bench/sets_config.txt", ["-k", "21", "-t", "3", "-l", "100", "-e", "1.9", "-n", "0", "-m", "9000"]),
])
def test_unmodified_commet_py_over_the_drop_in_tools(bin_dir, tmp_path, case, config, opts):
    """The seam itself (Commet.py:447,480-488): the reference's own orchestrator, byte for byte as shipped, run with
    `-b commet_b200/bin` -- the command tests/golden/make_golden.py ran with `-b oracle/_ref` to record the goldens."""
    import sys
    fixtures.materialize(tmp_path)
    r = subprocess.run([sys.executable, str(oracle.REF_DIR / "Commet.py"), config, "-b", str(bin_dir), *opts], cwd=tmp_path,
                       capture_output=True, text=True)
    out = tmp_path / "output_commet"
    assert (out / "matrix_plain.csv").exists(), (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    res = {n: (out / f"matrix_{n}.csv").read_text() for n in ("plain", "percentage", "normalized")}
    check_flow(tmp_path, res, GOLDEN[case])


@pytest.mark.parametrize("k", [20, 22, 33])
def test_chunk_boundary_known_answer(bin_dir, tmp_path, k):
    fixtures.materialize(tmp_path)
    (tmp_path / "a.txt").write_text("A:ABCDE_bench/A.fa\n")
    out = tmp_path / "o"
    r = subprocess.run([str(bin_dir / "index_and_search"), "-i", "a.txt", "-s", "a.txt", "-o", str(out), "-l", str(out),
                        "-k", str(k), "-t", "2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = GOLDEN["chunk_boundary_A_in_A"][str(k)]
    assert sha(out / "A.fa_in_A.bv") == g["bv_sha256"]
    m = re.search(r"indexed (\d+), searched (\d+), shared (\d+)", (out / "A_in_A.log").read_text())
    assert [int(x) for x in m.groups()] == g["counters"]


def test_full_mode(bin_dir, tmp_path):
    fixtures.materialize(tmp_path)
    (tmp_path / "fa.txt").write_text("set1:test_dissymmetry/A.fa\n")
    (tmp_path / "fb.txt").write_text("set2:test_dissymmetry/B.fa\n")
    out = tmp_path / "full"
    r = subprocess.run([str(bin_dir / "index_and_search"), "-i", "fa.txt", "-s", "fb.txt", "-o", str(out), "-l", str(out),
                        "-k", "25", "-f"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert {p.name: sha(p) for p in out.glob("*.bv")} == GOLDEN["full_mode_k25"]


def test_bvop_known_answers(bin_dir, tmp_path):
    fixtures.materialize(tmp_path)
    commet_flow.run("ABCDE_bench/sets_config.txt", bin_dir, tmp_path, k=32)
    for name, args in (("not", ["C.fa_in_set1.bv", "-n"]), ("and", ["A.fa_in_set2.bv", "-a", "A.fa_in_set3.bv"]),
                       ("or", ["A.fa_in_set2.bv", "-o", "A.fa_in_set3.bv"]),
                       ("andnot", ["A.fa_in_set2.bv", "-d", "A.fa_in_set3.bv"])):
        a = ["output_commet/" + x if x.endswith(".bv") else x for x in args]
        r = subprocess.run([str(bin_dir / "bvop"), *a, "-p", f"bvop_{name}.bv", "-i"], cwd=tmp_path, capture_output=True,
                           text=True)
        assert r.returncode == 0
        assert r.stdout == GOLDEN["bvop"][name]["stdout"]
        assert sha(tmp_path / f"bvop_{name}.bv") == GOLDEN["bvop"][name]["file_sha256"]
    # size mismatch exits 1 (boolean_vector.h:420-423)
    r = subprocess.run([str(bin_dir / "bvop"), "output_commet/A.fa_in_set2.bv", "-a", "output_commet/C.fa_in_set1.bv"],
                       cwd=tmp_path, capture_output=True)
    assert r.returncode == 1
    # no -p: header + raw payload on stdout
    r = subprocess.run([str(bin_dir / "bvop"), "output_commet/C.fa_in_set1.bv", "-n"], cwd=tmp_path, capture_output=True)
    assert r.stdout.startswith(b"NOT output_commet/C.fa_in_set1.bv\n\n#10000\n") and len(r.stdout.split(b"#10000\n", 1)[1]) == 1251


# ---- live against the reference binaries -------------------------------------------------------------
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref did not travel")


def _write_set(rng, tmp, name, reads_per_file, with_bv):
    items = []
    for fi, reads in enumerate(reads_per_file):
        kind = int(rng.integers(0, 5))
        base = tmp / f"{name}_{fi}"
        if kind == 0:
            path = H.write_fasta(base.with_suffix(".fa"), reads)
        elif kind == 1:
            path = H.write_fasta(base.with_suffix(".fa"), reads, width=int(rng.integers(7, 40)),
                                 final_newline=bool(rng.integers(0, 2)), blank_every=int(rng.integers(0, 4)))
        elif kind == 2:
            path = H.write_fastq(base.with_suffix(".fq"), reads, final_newline=bool(rng.integers(0, 2)))
        elif kind == 3:
            path = H.write_fasta(base.with_suffix(".fa.gz"), reads, gz=True)
        else:
            path = H.write_fastq(base.with_suffix(".fq.gz"), reads, gz=True)
        if with_bv:
            n = len(reads)
            valid = (rng.random(n) < 0.7).astype(np.uint8)
            if valid.sum() == 0:
                valid[int(rng.integers(0, n))] = 1
            bvp = tmp / f"{name}_{fi}.in.bv"
            oracle.write_bv_file(bvp, b"input", n, oracle.tags_to_bv(valid))
            items.append(f"{path},{bvp}")
        else:
            items.append(str(path))
    return items


@needs_ref
@pytest.mark.parametrize("seed", range(30))
def test_index_and_search_vs_reference_binary(bin_dir, tmp_path, seed):
    rng = np.random.default_rng(3000 + seed)
    k = int(rng.integers(8, 21))
    t = int(rng.integers(0, 4))
    L = int(rng.integers(k, 4 * k))
    dirt = dict(p_N=float(rng.choice([0, 0.02])), p_lower=float(rng.choice([0, 0.3])), p_other=float(rng.choice([0, 0.01])))
    maxk = oracle.max_kmer(k)
    n_ref = int(min(400, max(20, 3 * maxk // max(1, (L - k + 1)) + 5)))
    ref_files = [H.make_ref_set(rng, n_ref, max(1, L - 10), L + 10, **dirt) for _ in range(int(rng.integers(1, 3)))]
    all_ref = [r for f in ref_files for r in f]
    iitems = _write_set(rng, tmp_path, "idx", ref_files, with_bv=bool(rng.integers(0, 2)))
    (tmp_path / "index.txt").write_text(" refset:" + " ; ".join(iitems) + "\n")      # name keeps its spaces, files do not
    lines, names = [], []
    for s in range(int(rng.integers(1, 4))):
        files = [H.make_query_set(rng, all_ref, int(rng.integers(5, 120)), max(1, L - 10), L + 10, **dirt)
                 for _ in range(int(rng.integers(1, 3)))]
        lines.append(f"Q{s}:" + ";".join(_write_set(rng, tmp_path, f"q{s}", files, with_bv=bool(rng.integers(0, 2)))))
        names.append(f"Q{s}")
    (tmp_path / "query.txt").write_text("\n".join(lines) + "\n")
    outs = {}
    full = ["-f"] if seed % 5 == 4 else []
    for who, tool in (("ref", oracle.REF_DIR / "index_and_search"), ("gpu", bin_dir / "index_and_search")):
        out = tmp_path / who
        r = subprocess.run([str(tool), "-i", str(tmp_path / "index.txt"), "-s", str(tmp_path / "query.txt"), "-o", str(out),
                            "-l", str(out), "-k", str(k), "-t", str(t), *full], capture_output=True, text=True)
        assert r.returncode == 0, (who, r.stderr)
        outs[who] = out
    ref_bvs = {p.name: p.read_bytes() for p in outs["ref"].glob("*.bv")}
    gpu_bvs = {p.name: p.read_bytes() for p in outs["gpu"].glob("*.bv")}
    assert ref_bvs.keys() == gpu_bvs.keys() and len(ref_bvs) > 0
    for name in ref_bvs:
        assert ref_bvs[name] == gpu_bvs[name], (seed, k, t, name)
    pat = re.compile(r"\[indexed \d+, searched \d+, shared \d+\](\n[0-9.eE+-]+%)?")
    for lg in outs["ref"].glob("*.log"):
        a = pat.search(lg.read_text()).group(0)
        b = pat.search((outs["gpu"] / lg.name).read_text()).group(0)
        assert a == b, (seed, lg.name)


@pytest.mark.skipif(not (oracle.REF_DIR / "compare_reads").exists(), reason="oracle/_ref/compare_reads did not travel")
@pytest.mark.parametrize("seed", range(6))
def test_compare_reads_vs_reference_binary(bin_dir, tmp_path, seed):
    """src/compare_reads.cpp:237-333: the three passes behind their own argv; stdout (minus the clock lines) and both
    .bv files equal the reference's.  One chunk per index: with a lost read the reference's loop never ends (:248)."""
    rng = np.random.default_rng(8000 + seed)
    k = int(rng.integers(14, 25))
    t = int(rng.integers(0, 4))
    L = int(rng.integers(k, 4 * k))
    dirt = dict(p_N=float(rng.choice([0, 0.02])), p_lower=float(rng.choice([0, 0.3])))
    per_read = max(1, L - k + 1)
    n_max = int(max(5, min(300, oracle.max_kmer(k) // (2 * per_read) - 2)))     # both sets stay below max_kmer
    a_files = [H.make_ref_set(rng, int(rng.integers(3, n_max // 2 + 4)), max(1, L - 10), L + 10, **dirt) for _ in range(int(rng.integers(1, 3)))]
    all_a = [r for f in a_files for r in f]
    b_files = [H.make_query_set(rng, all_a, int(rng.integers(3, n_max // 2 + 4)), max(1, L - 10), L + 10, **dirt)
               for _ in range(int(rng.integers(1, 3)))]
    (tmp_path / "a.txt").write_text("setA:" + ";".join(_write_set(rng, tmp_path, "a", a_files, with_bv=bool(seed % 2))) + "\n")
    (tmp_path / "b.txt").write_text("setB:" + ";".join(_write_set(rng, tmp_path, "b", b_files, with_bv=bool(seed % 3 == 0))) + "\n")
    res = {}
    for who, tool in (("ref", oracle.REF_DIR / "compare_reads"), ("gpu", bin_dir / "compare_reads")):
        out = tmp_path / who
        r = subprocess.run([str(tool), "-i", str(tmp_path / "a.txt"), "-s", str(tmp_path / "b.txt"), "-o", str(out), "-l", str(out),
                            "-k", str(k), "-t", str(t)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, (who, r.stderr)
        text = [ln for ln in r.stdout.split("\n") if not re.match(r"(Index |Search|Total ) time", ln)]
        res[who] = ({p.name: p.read_bytes() for p in out.glob("*.bv")}, text)
    assert res["ref"][0].keys() == res["gpu"][0].keys() and len(res["ref"][0]) >= 2
    assert res["ref"][0] == res["gpu"][0], (seed, k, t)
    assert res["ref"][1] == res["gpu"][1], (seed, k, t)


@needs_ref
def test_unreadable_file_with_a_vector_is_skipped_and_the_run_goes_on(bin_dir, tmp_path):
    """FileManager::addFile(file, bv) (file_manager.h:167-175, the form Commet.py's "file,bv" entries take): one message,
    the file is skipped, the other files of the set are processed"""
    rng = np.random.default_rng(77)
    ref = H.make_ref_set(rng, 60, 50, 70)
    qry = H.make_query_set(rng, ref, 40, 50, 70)
    H.write_fasta(tmp_path / "r.fa", ref)
    H.write_fasta(tmp_path / "q.fa", qry)
    oracle.write_bv_file(tmp_path / "all_r.bv", b"x", len(ref), oracle.tags_to_bv(np.ones(len(ref), dtype=np.uint8)))
    oracle.write_bv_file(tmp_path / "all_q.bv", b"x", len(qry), oracle.tags_to_bv(np.ones(len(qry), dtype=np.uint8)))
    (tmp_path / "i.txt").write_text("I:nofile.fa,all_r.bv;r.fa,all_r.bv\n")
    (tmp_path / "s.txt").write_text("S:q.fa,all_q.bv;gone.fa,all_q.bv\n")
    res = {}
    for who, tool in (("ref", oracle.REF_DIR / "index_and_search"), ("gpu", bin_dir / "index_and_search")):
        r = subprocess.run([str(tool), "-i", "i.txt", "-s", "s.txt", "-o", who, "-l", who, "-k", "16"], cwd=tmp_path,
                           capture_output=True, text=True)
        res[who] = (r.returncode, r.stderr, {p.name: p.read_bytes() for p in (tmp_path / who).glob("*.bv")})
    assert res["ref"] == res["gpu"]
    assert res["ref"][0] == 0 and len(res["ref"][2]) == 1


@needs_ref
def test_negative_max_n_drops_every_read_that_passes_the_length_test(bin_dir, tmp_path):
    """`number_of_N(read) > max_N` with a negative max_N is true for every read (src/filter_reads.cpp:192)"""
    rng = np.random.default_rng(78)
    H.write_fasta(tmp_path / "in.fa", H.make_ref_set(rng, 200, 20, 90, p_N=0.02))
    res = {}
    for who, tool in (("ref", oracle.REF_DIR / "filter_reads"), ("gpu", bin_dir / "filter_reads")):
        r = subprocess.run([str(tool), "in.fa", "-l", "40", "-n", "-3", "-o", f"{who}.bv"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, (who, r.stderr)
        res[who] = ((tmp_path / f"{who}.bv").read_bytes(), [ln for ln in r.stdout.split("\n") if not ln.startswith("Total  time")])
    assert res["ref"] == res["gpu"]
    assert "Number of selected reads = 0" in res["gpu"][1]


@needs_ref
def test_more_than_thirty_search_sets_in_one_invocation(bin_dir, tmp_path):
    """Commet.py puts every other sample into one -s file; the reference takes any number of search sets"""
    rng = np.random.default_rng(79)
    ref = H.make_ref_set(rng, 300, 40, 60)      # several chunks at k=14
    H.write_fasta(tmp_path / "r.fa", ref)
    (tmp_path / "i.txt").write_text("I:r.fa\n")
    lines = []
    for s in range(37):
        H.write_fasta(tmp_path / f"q{s}.fa", H.make_query_set(rng, ref, int(rng.integers(3, 30)), 40, 60))
        lines.append(f"Q{s:02d}:q{s}.fa")
    (tmp_path / "s.txt").write_text("\n".join(lines) + "\n")
    res = {}
    for who, tool in (("ref", oracle.REF_DIR / "index_and_search"), ("gpu", bin_dir / "index_and_search")):
        r = subprocess.run([str(tool), "-i", "i.txt", "-s", "s.txt", "-o", who, "-l", who, "-k", "14", "-t", "1"], cwd=tmp_path,
                           capture_output=True, text=True)
        assert r.returncode == 0, (who, r.stderr)
        res[who] = {p.name: p.read_bytes() for p in (tmp_path / who).glob("*.bv")}
    assert len(res["ref"]) == 37 and res["ref"] == res["gpu"]


@needs_ref
@pytest.mark.parametrize("devices,full", [("0,0", False), ("0,0,0", True), ("0,0,0,0", False)])
def test_index_and_search_tool_over_several_ranks(bin_dir, tmp_path, devices, full):
    """COMMET_B200_DEVICES / COMMET_B200_GPUS: the drop-in tool deals the index set over the GPUs (commet_group_*); files
    and logs are those of the reference binary, also for the three passes of -f"""
    import os
    rng = np.random.default_rng(500 + len(devices))
    ref_files = [H.make_ref_set(rng, 900, 40, 110, p_N=0.01) for _ in range(2)]
    all_ref = [r for f in ref_files for r in f]
    (tmp_path / "index.txt").write_text("R:" + ";".join(_write_set(rng, tmp_path, "idx", ref_files, with_bv=True)) + "\n")
    lines = []
    for s in range(1 if full else 3):
        files = [H.make_query_set(rng, all_ref, int(rng.integers(40, 700)), 40, 110, p_N=0.01) for _ in range(2)]
        lines.append(f"Q{s}:" + ";".join(_write_set(rng, tmp_path, f"q{s}", files, with_bv=bool(s % 2))))
    (tmp_path / "query.txt").write_text("\n".join(lines) + "\n")
    outs = {}
    for who, tool, env in (("ref", oracle.REF_DIR / "index_and_search", {}), ("gpu", bin_dir / "index_and_search",
                                                                           {"COMMET_B200_DEVICES": devices, "COMMET_B200_DIST_BLOCK": "50"})):
        out = tmp_path / who
        r = subprocess.run([str(tool), "-i", str(tmp_path / "index.txt"), "-s", str(tmp_path / "query.txt"), "-o", str(out),
                            "-l", str(out), "-k", "16", "-t", "2", *(["-f"] if full else [])], capture_output=True, text=True,
                           env={**os.environ, **env})
        assert r.returncode == 0, (who, r.stderr)
        outs[who] = out
    a = {p.name: p.read_bytes() for p in outs["ref"].glob("*.bv")}
    b = {p.name: p.read_bytes() for p in outs["gpu"].glob("*.bv")}
    assert a == b and len(a) >= 2
    pat = re.compile(r"\[indexed \d+, searched \d+, shared \d+\](\n[0-9.eE+-]+%)?")
    for lg in outs["ref"].glob("*.log"):
        assert pat.search(lg.read_text()).group(0) == pat.search((outs["gpu"] / lg.name).read_text()).group(0), lg.name


@needs_ref
@pytest.mark.parametrize("seed", range(20))
def test_filter_reads_vs_reference_binary(bin_dir, tmp_path, seed):
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(1, 3000))
    reads = []
    for _ in range(n):
        L = int(rng.integers(1, 160))
        kind = rng.random()
        if kind < 0.15:
            r = np.full(L, ord(rng.choice(list("ACGTacgtN"))), dtype=np.uint8)
        elif kind < 0.3:
            r = np.tile(np.frombuffer(bytes(rng.choice([b"AC", b"AT", b"ACGT", b"AAC"])), dtype=np.uint8), L)[:L]
        else:
            r = H.dirty(rng, H.random_read(rng, L), p_N=float(rng.choice([0, 0.05])), p_lower=float(rng.choice([0, 0.4])),
                        p_other=float(rng.choice([0, 0.02])))
        reads.append(r.tobytes())
    kind = seed % 4
    if kind == 0: path = H.write_fasta(tmp_path / "in.fa", reads)
    elif kind == 1: path = H.write_fastq(tmp_path / "in.fq", reads)
    elif kind == 2: path = H.write_fasta(tmp_path / "in.fa.gz", reads, gz=True, width=31)
    else: path = H.write_fastq(tmp_path / "in.fq.gz", reads, gz=True)
    args = []
    if rng.random() < 0.8: args += ["-l", str(int(rng.integers(0, 120)))]
    if rng.random() < 0.6: args += ["-n", str(int(rng.integers(0, 6)))]
    if rng.random() < 0.8: args += ["-e", str(rng.choice(["1.0", "1.5", "2", "0.5", str(round(float(rng.uniform(0, 2.1)), 4))]))]
    if rng.random() < 0.4: args += ["-m", str(float(int(rng.integers(0, n + 3))))]
    if rng.random() < 0.3: args += ["-c", "my comment"]
    res = {}
    for who, tool in (("ref", oracle.REF_DIR / "filter_reads"), ("gpu", bin_dir / "filter_reads")):
        outp = tmp_path / f"{who}.bv"
        r = subprocess.run([str(tool), str(path), *args, "-o", str(outp)], capture_output=True, text=True)
        assert r.returncode == 0, (who, r.stderr)
        res[who] = (outp.read_bytes(), [ln for ln in r.stdout.split("\n") if not ln.startswith("Total  time")])
    assert res["ref"][0] == res["gpu"][0], (seed, args)
    assert res["ref"][1] == res["gpu"][1], (seed, args)


# ---- commet_nxn: the whole Commet.py run in one process on resident sets ----------------------------------
def _nxn(bin_dir, cwd, config, out="output_commet/", gpus=None, **kw):
    args = [str(bin_dir / "commet_nxn"), config, "-o", out, "-q"]
    for key, val in kw.items():
        args += [f"-{key}", str(val)]
    if gpus:
        args += ["--gpus", str(gpus)]
    r = subprocess.run(args, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o = Path(cwd) / out
    return {n: (o / f"matrix_{n}.csv").read_text() for n in ("plain", "percentage", "normalized")}


@pytest.mark.parametrize("case,config,kw", [
    ("abcde_3sets_k32", "ABCDE_bench/sets_config.txt", dict(k=32)),
    ("abcde_5sets_k32", "ABCDE_bench/five_sets.txt", dict(k=32)),
    ("dissymmetry_k33", "test_dissymmetry/fof.txt", dict(k=33)),
    ("abcde_3sets_k21_filtered", "ABCDE_bench/sets_config.txt", dict(k=21, t=3, l=100, e=1.9, n=0, m=9000)),
])
def test_commet_nxn_bit_exact(bin_dir, tmp_path, case, config, kw):
    """One process, resident sets, device-side matrices: every .bv and the three CSVs byte-identical to the
    unmodified Commet.py + reference binaries (golden)."""
    fixtures.materialize(tmp_path)
    res = _nxn(bin_dir, tmp_path, config, **kw)
    check_flow(tmp_path, res, GOLDEN[case])


def _random_config(rng, tmp, n_sets, L, dirt):
    base = H.make_ref_set(rng, int(rng.integers(60, 400)), max(1, L - 10), L + 10, **dirt)
    lines = []
    for s in range(n_sets):
        files = []
        for _ in range(int(rng.integers(1, 3))):
            files.append(H.make_query_set(rng, base, int(rng.integers(20, 300)), max(1, L - 10), L + 10,
                                          frac_shared=float(rng.uniform(0.2, 0.9)), **dirt))
        items = _write_set(rng, tmp, f"s{s}", files, with_bv=False)
        lines.append(f"set{s} : " + " ; ".join(str(Path(i).relative_to(tmp)) for i in items))
    (tmp / "cfg.txt").write_text("\n".join(lines) + "\n")
    return "cfg.txt"


@pytest.mark.parametrize("seed", range(4))
def test_commet_nxn_vs_tool_flow(bin_dir, tmp_path, seed):
    """Random multi-file sets (FASTA/FASTQ/gz), several chunks per index, filters on: commet_nxn against the
    Commet.py flow driven over the drop-in tools (themselves pinned against the reference binaries)."""
    rng = np.random.default_rng(15000 + seed)
    k = int(rng.integers(10, 19))
    t = int(rng.integers(1, 4))
    L = int(rng.integers(2 * k, 5 * k))
    dirt = dict(p_N=float(rng.choice([0, 0.02])), p_lower=float(rng.choice([0, 0.3])))
    cfg = _random_config(rng, tmp_path, int(rng.integers(2, 4)), L, dirt)     # 3 or 8 tool rounds + filters + bvop -i forks
    kw = dict(k=k, t=t)
    if seed % 2 == 1:
        kw.update(l=int(rng.integers(1, L)), e=round(float(rng.uniform(0.5, 1.95)), 3), n=int(rng.integers(0, 3)))
    if seed % 4 == 2:
        kw.update(m=int(rng.integers(10, 400)))
    exp = commet_flow.run(cfg, bin_dir, tmp_path, out_dir="flow_out/", **kw)
    got = _nxn(bin_dir, tmp_path, cfg, out="nxn_out/", **kw)
    assert got == exp, (seed, kw)
    a = {p.name: p.read_bytes() for p in (tmp_path / "flow_out").glob("*.bv")}
    b = {p.name: p.read_bytes() for p in (tmp_path / "nxn_out").glob("*.bv")}
    assert a.keys() == b.keys() and len(a) > 0
    for name in a:
        # the comments embed the output directory only through the file paths given in the config: identical here
        assert a[name] == b[name], (seed, name)
    pat = re.compile(r"\[indexed \d+, searched \d+, shared \d+\]")
    for lg in (tmp_path / "flow_out").glob("*.log"):
        assert pat.search(lg.read_text()).group(0) == pat.search((tmp_path / "nxn_out" / lg.name).read_text()).group(0), lg.name


def test_commet_nxn_multi_gpu_matches_single(bin_dir, tmp_path):
    """rounds spread over every visible GPU give the files of the single-GPU run"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fixtures.materialize(tmp_path)
    one = _nxn(bin_dir, tmp_path, "ABCDE_bench/five_sets.txt", out="one/", gpus=1, k=32)
    many = _nxn(bin_dir, tmp_path, "ABCDE_bench/five_sets.txt", out="many/", k=32)
    assert one == many == GOLDEN["abcde_5sets_k32"]["csv"]
    a = {p.name: sha(p) for p in (tmp_path / "one").glob("*.bv")}
    b = {p.name: sha(p) for p in (tmp_path / "many").glob("*.bv")}
    assert a == b == GOLDEN["abcde_5sets_k32"]["bv"]


def test_python_float_formatting_of_matrices(bin_dir, tmp_path):
    """CSV numbers are Python 3 str(float): exponent form below 1e-4, shortest round-trip digits"""
    rng = np.random.default_rng(5)
    big = [H.random_read(rng, 40).tobytes() for _ in range(30011)]
    H.write_fasta(tmp_path / "big.fa", big)
    H.write_fasta(tmp_path / "small.fa", [big[7], big[123], H.random_read(rng, 40).tobytes()])
    (tmp_path / "c.txt").write_text("big:big.fa\nsmall:small.fa\n")
    got = _nxn(bin_dir, tmp_path, "c.txt", out="o/", k=20, t=1)
    rows = [ln.split(";") for ln in got["percentage"].strip().split("\n")]
    shared = [ln.split(";") for ln in got["plain"].strip().split("\n")]
    c, n = int(shared[1][2]), int(shared[1][1])
    assert c == 2 and n == 30011
    assert rows[1][2] == str(100 * c / float(n)) and "e-" not in rows[1][2]
    norm = [ln.split(";") for ln in got["normalized"].strip().split("\n")]
    assert norm[1][2] == str(100 * (c + int(shared[2][1])) / float(n + 3))
