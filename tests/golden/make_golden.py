"""Regenerates tests/golden/golden.json and the fixture copies under tests/golden/data/.

Needs /root/reference (the reference checkout) and oracle/_ref (its tools compiled by oracle/Makefile).
Everything recorded here is produced by the UNMODIFIED reference: Commet.py driving its own binaries.
The fixture FASTA files are the reference's own test data (ABCDE_bench/, test_dissymmetry/), stored gzipped;
B.fa == D.fa and C.fa == E.fa byte for byte, and test_dissymmetry/A.fa == ABCDE_bench/A.fa, so three files
plus the two small dissymmetry files are kept.

    python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
BIN = ROOT / "oracle" / "_ref"
sys.path.insert(0, str(ROOT))


def sha(p):
    return hashlib.sha256(Path(p).read_bytes()).hexdigest()


def materialize(work: Path):
    """lay the fixtures out under `work` exactly as the tests do (from tests/golden/data)"""
    from tests.golden import fixtures
    fixtures.materialize(work)


def commet(work, config, *opts):
    out = work / "output_commet"
    if out.exists():
        shutil.rmtree(out)
    subprocess.run([sys.executable, str(REF / "Commet.py"), config, "-b", str(BIN), *opts], cwd=work,
                   capture_output=True, text=True)
    res = {"bv": {p.name: sha(p) for p in sorted(out.glob("*.bv"))},
           "csv": {n: (out / f"matrix_{n}.csv").read_text() for n in ("plain", "percentage", "normalized")}}
    res["csv_sha256"] = {n: hashlib.sha256(v.encode()).hexdigest() for n, v in res["csv"].items()}
    return res


def main():
    data = HERE / "data"
    data.mkdir(exist_ok=True)
    for src, dst in (("ABCDE_bench/A.fa", "A.fa.gz"), ("ABCDE_bench/B.fa", "B.fa.gz"), ("ABCDE_bench/C.fa", "C.fa.gz"),
                     ("test_dissymmetry/B.fa", "dissym_B.fa.gz"), ("test_dissymmetry/C.fa", "dissym_C.fa.gz")):
        with open(REF / src, "rb") as f, open(data / dst, "wb") as g:
            g.write(gzip.compress(f.read(), 9, mtime=0))
    shutil.copy(REF / "ABCDE_bench" / "sets_config.txt", data / "sets_config.txt")
    golden = {}
    with tempfile.TemporaryDirectory() as td:
        work = Path(td)
        materialize(work)
        golden["abcde_3sets_k32"] = commet(work, "ABCDE_bench/sets_config.txt", "-k", "32")
        golden["abcde_5sets_k32"] = commet(work, "ABCDE_bench/five_sets.txt", "-k", "32")
        golden["dissymmetry_k33"] = commet(work, "test_dissymmetry/fof.txt", "-k", "33")
        # filters on, small k so that every index_and_search needs several chunks (max_kmer(21) = 244140)
        golden["abcde_3sets_k21_filtered"] = commet(work, "ABCDE_bench/sets_config.txt", "-k", "21", "-t", "3", "-l", "100",
                                                    "-e", "1.9", "-n", "0", "-m", "9000")
        # chunk-boundary known answers: A.fa in A.fa
        (work / "a.txt").write_text("A:ABCDE_bench/A.fa\n")
        chunk = {}
        for k in (20, 22, 33):
            out = work / f"chunk{k}"
            subprocess.run([str(BIN / "index_and_search"), "-i", "a.txt", "-s", "a.txt", "-o", str(out), "-l", str(out),
                            "-k", str(k), "-t", "2"], cwd=work, capture_output=True, check=True)
            log = (out / "A_in_A.log").read_text()
            chunk[str(k)] = {"bv_sha256": sha(out / "A.fa_in_A.bv"),
                             "counters": [int(x) for x in re.search(r"indexed (\d+), searched (\d+), shared (\d+)", log).groups()]}
        golden["chunk_boundary_A_in_A"] = chunk
        # -f full mode (three passes in one process), dissymmetry sets
        (work / "fa.txt").write_text("set1:test_dissymmetry/A.fa\n")
        (work / "fb.txt").write_text("set2:test_dissymmetry/B.fa\n")
        out = work / "full"
        subprocess.run([str(BIN / "index_and_search"), "-i", "fa.txt", "-s", "fb.txt", "-o", str(out), "-l", str(out), "-k", "25",
                        "-f"], cwd=work, capture_output=True, check=True)
        golden["full_mode_k25"] = {p.name: sha(p) for p in sorted(out.glob("*.bv"))}
        # bvop known answers
        o3 = work / "o3"
        commet3 = commet(work, "ABCDE_bench/sets_config.txt", "-k", "32")
        oc = work / "output_commet"
        bv = {}
        for name, args in (("not", ["C.fa_in_set1.bv", "-n"]), ("and", ["A.fa_in_set2.bv", "-a", "A.fa_in_set3.bv"]),
                           ("or", ["A.fa_in_set2.bv", "-o", "A.fa_in_set3.bv"]),
                           ("andnot", ["A.fa_in_set2.bv", "-d", "A.fa_in_set3.bv"])):
            a = ["output_commet/" + x if x.endswith(".bv") else x for x in args]
            r = subprocess.run([str(BIN / "bvop"), *a, "-p", f"bvop_{name}.bv", "-i"], cwd=work, capture_output=True, text=True)
            bv[name] = {"file_sha256": sha(work / f"bvop_{name}.bv"), "stdout": r.stdout}
        golden["bvop"] = bv
    (HERE / "golden.json").write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
