"""CPU: what the drop-in executables do BEFORE they need the GPU -- usage, version, unknown options, missing option
values, missing mandatory arguments, unreadable inputs -- is the reference's, byte for byte on stdout and stderr and in
the exit code (oracle/_ref = the reference's own sources compiled by oracle/Makefile)."""
import subprocess

import numpy as np
import pytest

from commet_b200 import build
from oracle import oracle

TOOLS = ("index_and_search", "filter_reads", "bvop", "extract_reads", "compare_reads")
pytestmark = pytest.mark.skipif(not all((oracle.REF_DIR / t).exists() for t in TOOLS), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    build.build_tools()
    d = tmp_path_factory.mktemp("cli")
    oracle.write_bv_file(d / "a.bv", b"a", 10, oracle.tags_to_bv(np.ones(10, dtype=np.uint8)))
    oracle.write_bv_file(d / "b.bv", b"b", 12, oracle.tags_to_bv(np.ones(12, dtype=np.uint8)))
    (d / "x.fa").write_text(">r\nACGT\n")
    (d / "sets.txt").write_text("s:x.fa\n")
    (d / "missing.txt").write_text("s:nofile.fa\n")
    (d / "two.txt").write_text("a:x.fa\nb:x.fa\n")
    (d / "empty.txt").write_text("\n\n")
    (d / "afile").write_text("")
    (d / "badbv.txt").write_text("s:x.fa,b.bv\n")
    return d


CASES = [
    ("index_and_search", []), ("index_and_search", ["-h"]), ("index_and_search", ["-v"]), ("index_and_search", ["-z"]),
    ("index_and_search", ["-i"]), ("index_and_search", ["-i", "sets.txt", "-s"]), ("index_and_search", ["-k"]),
    ("index_and_search", ["-t"]), ("index_and_search", ["-i", "sets.txt"]), ("index_and_search", ["-s", "sets.txt"]),
    ("index_and_search", ["-i", "nofile.txt", "-s", "nofile.txt", "-o", "o1", "-l", "o1"]),
    ("index_and_search", ["-k", "31", "-t", "3", "-z"]),
    # FileManager::addFile: an unreadable read file is reported twice and ends the run; a vector of the wrong size too
    # (file_manager.h:119-143, fasta_file.h:104-107) -- on the index side and on the search side
    ("index_and_search", ["-i", "missing.txt", "-s", "sets.txt", "-o", "o2", "-l", "o2"]),
    ("index_and_search", ["-i", "sets.txt", "-s", "missing.txt", "-o", "o3", "-l", "o3"]),
    ("index_and_search", ["-i", "badbv.txt", "-s", "sets.txt", "-o", "o4", "-l", "o4"]),
    ("index_and_search", ["-i", "sets.txt", "-s", "badbv.txt", "-o", "o5", "-l", "o5"]),
    # output / log paths that exist and are not directories (src/index_and_search.cpp:178-191), more than one index
    # set, a repeated -i, an index fof without any set
    ("index_and_search", ["-i", "sets.txt", "-s", "sets.txt", "-o", "afile", "-l", "o6"]),
    ("index_and_search", ["-i", "sets.txt", "-s", "sets.txt", "-o", "o7", "-l", "afile"]),
    ("index_and_search", ["-i", "two.txt", "-s", "missing.txt", "-o", "o8", "-l", "o8"]),
    ("index_and_search", ["-i", "sets.txt", "-i", "two.txt", "-s", "missing.txt", "-o", "o9", "-l", "o9"]),
    ("index_and_search", ["-i", "empty.txt", "-s", "sets.txt", "-o", "o10", "-l", "o10"]),
    # compare_reads (src/compare_reads.cpp:82-183): same argv conventions, its own usage text and messages
    ("compare_reads", []), ("compare_reads", ["-h"]), ("compare_reads", ["-v"]), ("compare_reads", ["-z"]),
    ("compare_reads", ["-i"]), ("compare_reads", ["-i", "sets.txt", "-s"]), ("compare_reads", ["-k", "31", "-t", "3", "-z"]),
    ("compare_reads", ["-i", "sets.txt", "-i", "two.txt", "-s", "sets.txt", "-s", "two.txt", "-o", "afile", "-l", "c1"]),
    ("compare_reads", ["-i", "sets.txt", "-s", "sets.txt", "-o", "c2", "-l", "afile"]),
    ("compare_reads", ["-i", "missing.txt", "-s", "sets.txt", "-o", "c3", "-l", "c3"]),
    ("compare_reads", ["-i", "sets.txt", "-s", "badbv.txt", "-o", "c4", "-l", "c4"]),
    ("filter_reads", []), ("filter_reads", ["-h"]), ("filter_reads", ["-v"]), ("filter_reads", ["-z"]),
    ("filter_reads", ["nofile.fa", "-o", "x.bv"]),
    ("bvop", []), ("bvop", ["-h"]), ("bvop", ["-v"]), ("bvop", ["-z"]), ("bvop", ["nofile.bv", "-i"]), ("bvop", ["a.bv", "-q"]),
    # the host-side checks of a binary operator come before the device is touched: size mismatch, missing second vector,
    # and an invocation with nothing to compute (boolean_vector.h:420-423, src/bvop.cpp:133-160)
    ("bvop", ["a.bv", "-a", "b.bv"]), ("bvop", ["a.bv", "-d", "b.bv", "-p", "out.bv"]), ("bvop", ["a.bv", "-o", "nofile.bv"]),
    ("bvop", ["a.bv"]),
    ("extract_reads", []), ("extract_reads", ["-h"]), ("extract_reads", ["-v"]), ("extract_reads", ["-z"]),
    ("extract_reads", ["x.fa"]), ("extract_reads", ["nofile.fa", "a.bv"]), ("extract_reads", ["x.fa", "nofile.bv"]),
]


@pytest.mark.parametrize("tool,args", CASES, ids=[f"{t}-{'_'.join(a) or 'noargs'}" for t, a in CASES])
def test_cli_behaviour_before_the_gpu_is_needed(work, tool, args):
    mine = subprocess.run([str(build.BIN / tool), *args], cwd=work, capture_output=True, timeout=60)
    ref = subprocess.run([str(oracle.REF_DIR / tool), *args], cwd=work, capture_output=True, timeout=60)
    assert ref.returncode >= 0, "the reference itself crashed on this case: not a comparison"
    assert mine.returncode == ref.returncode
    assert mine.stdout == ref.stdout
    assert mine.stderr == ref.stderr


def test_missing_option_values_are_clean_errors_where_the_reference_crashes(work):
    """`filter_reads x.fa -l` and `bvop a.bv -a` read argv[argc] in the reference (segmentation fault); here exit 1"""
    for tool, args in (("filter_reads", ["x.fa", "-l"]), ("bvop", ["a.bv", "-a"])):
        r = subprocess.run([str(build.BIN / tool), *args], cwd=work, capture_output=True, timeout=60)
        assert r.returncode == 1 and b"needs an argument" in r.stderr
