"""Lays the committed fixture files out the way the reference repo does (relative paths matter: the
.bv comments embed the read-file paths as given on the command line)."""
import gzip
import shutil
from pathlib import Path

DATA = Path(__file__).resolve().parent / "data"


def _gunzip(src: Path, dst: Path):
    dst.parent.mkdir(parents=True, exist_ok=True)
    dst.write_bytes(gzip.decompress(src.read_bytes()))


def materialize(work: Path):
    ab = work / "ABCDE_bench"
    for name, src in (("A", "A"), ("B", "B"), ("C", "C"), ("D", "B"), ("E", "C")):   # D == B, E == C
        _gunzip(DATA / f"{src}.fa.gz", ab / f"{name}.fa")
    shutil.copy(DATA / "sets_config.txt", ab / "sets_config.txt")
    (ab / "five_sets.txt").write_text("".join(f"{n}:ABCDE_bench/{n}.fa\n" for n in "ABCDE"))
    td = work / "test_dissymmetry"
    _gunzip(DATA / "A.fa.gz", td / "A.fa")
    _gunzip(DATA / "dissym_B.fa.gz", td / "B.fa")
    _gunzip(DATA / "dissym_C.fa.gz", td / "C.fa")
    # the reference's fof.txt points at a non-existent testsafac/ directory; same sets, real paths
    (td / "fof.txt").write_text("set1: test_dissymmetry/A.fa\nset2: test_dissymmetry/B.fa\nset3: test_dissymmetry/C.fa\n")
