"""CPU: host-side pieces of commet_nxn that need no GPU -- the Python-3 float formatting of the CSV matrices
(Commet.py:298,313 write str(float)) is compiled out of the tool's source and compared with Python itself."""
import random
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_py_float_matches_python_str(tmp_path):
    src = (ROOT / "commet_b200" / "csrc" / "tools" / "commet_nxn.cpp").read_text()
    fn = src[src.index("static std::string py_float(double x)"):src.index("static void ensure_dir")]
    (tmp_path / "t.cpp").write_text(
        "#include <charconv>\n#include <cmath>\n#include <string>\n#include <cstdio>\n#include <cstdlib>\n" + fn +
        'int main(){double x; while(scanf("%la",&x)==1) puts(py_float(x).c_str());}\n')
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(tmp_path / "t"), str(tmp_path / "t.cpp")], check=True)
    rnd = random.Random(1)
    vals = [0.0, 100.0, 25.0, 100 / 3, 1e-5, 5e-06, 1e16, 1e15, 123456789012345678.0, 0.0001, 0.00012345, 1.5e-7,
            200 / 3, 99.99999999999999, 1e22, 3.0e-310, 100 * 1 / float(20_000_000), 100 * 7 / float(500_000_000)]
    for _ in range(5000):
        c, n = rnd.randint(0, 10 ** rnd.randint(1, 9)), rnd.randint(1, 10 ** rnd.randint(1, 9))
        vals += [100 * c / float(n), 100 * (c + n) / float(n + rnd.randint(1, 10 ** 9))]
    out = subprocess.run([str(tmp_path / "t")], input="\n".join(float.hex(v) for v in vals), capture_output=True,
                         text=True, check=True).stdout.split("\n")
    bad = [(v, o) for v, o in zip(vals, out) if o != str(v)]
    assert not bad, bad[:5]


def test_parallel_fasta_loader_equals_sequential_reader(tmp_path):
    """fast_fasta.hpp (mmap, chunked, two passes) yields the records of readers.hpp's parse_fasta -- the restatement of
    the reference's FastaFile -- for any chunk size: multi-line records, blank lines, CR, '>' inside lines, no final newline."""
    import numpy as np
    from tests import helpers as H
    host = ROOT / "commet_b200" / "csrc" / "host"
    (tmp_path / "h.cpp").write_text(r'''
#include "fast_fasta.hpp"
#include "readers.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace commet_host;
int main(int argc, char **argv) {
    ParsedFile pf;
    if (!parse_reads_file(argv[1], pf, "x")) return 2;
    FastaMap m;
    if (!m.open(argv[1], (size_t)atol(argv[2]))) return 3;
    for (size_t c = 0; c < m.n_chunks(); c++) m.pass<false>(c, nullptr, 0, nullptr, 0);
    m.finish_scan();
    if (m.n_records != pf.nb_reads || m.n_bytes != pf.seq.size()) { printf("counts %lu %lu vs %lu %lu\n", (unsigned long)m.n_records, (unsigned long)m.n_bytes, (unsigned long)pf.nb_reads, (unsigned long)pf.seq.size()); return 1; }
    std::vector<uint8_t> seq(m.n_bytes + 1);
    std::vector<uint64_t> off(m.n_records + 1, 0);
    off[m.n_records] = m.n_bytes;
    uint64_t pos = 0, rec = 0;
    for (size_t c = 0; c < m.n_chunks(); c++) { m.pass<true>(c, seq.data(), pos, off.data(), rec); pos += m.bytes[c]; rec += m.records[c]; }
    seq.resize(m.n_bytes);
    if (seq.size() != pf.seq.size() || memcmp(seq.data(), pf.seq.data(), seq.size()) != 0 || off != pf.off) { puts("content differs"); return 1; }
    printf("ok %lu records %lu chunks\n", (unsigned long)m.n_records, (unsigned long)m.n_chunks());
    return 0;
}''')
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(host), "-o", str(tmp_path / "h"), str(tmp_path / "h.cpp"), "-lz"], check=True)
    rng = np.random.default_rng(3)
    for case in range(12):
        reads = H.make_ref_set(rng, int(rng.integers(1, 400)), 1, int(rng.integers(2, 300)), p_N=0.02, p_lower=0.2)
        if case % 3 == 0:
            reads = [r.replace(b"C", b">", 1) if i % 5 == 0 and len(r) > 3 and not r.startswith(b"C") else r for i, r in enumerate(reads)]
        p = H.write_fasta(tmp_path / f"c{case}.fa", reads, width=[None, 7, 60, 13][case % 4], final_newline=case % 2 == 0,
                          blank_every=[0, 3, 1][case % 3])
        if case % 4 == 3:
            p.write_bytes(p.read_bytes().replace(b"\n", b"\r\n", 5))
        for chunk in (16, 64, 1000, 1 << 20):
            r = subprocess.run([str(tmp_path / "h"), str(p), str(chunk)], capture_output=True, text=True)
            assert r.returncode == 0, (case, chunk, r.stdout)
