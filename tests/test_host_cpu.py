"""CPU: the C++ host side of the drop-in tools without a GPU -- the fof grammar (csrc/host/set_parser.hpp =
include/set_parser.h:46-102), the read-file parsers (csrc/host/readers.hpp = FastaFile/FastqFile and their gzip
twins), the valid-read stream and per-file vectors of a set (csrc/host/read_set.hpp = FileManager,
include/file_manager.h:88-112,117-222,245-252) and the .bv files (csrc/host/bv.hpp = boolean_vector.h:302-414).
A tiny driver built from those headers dumps what the tools would hand to the C-ABI; the expectation comes from the
oracle's restatements of the same formats (pinned against the reference binaries in test_oracle_vs_ref.py)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from tests import helpers as H

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "commet_b200" / "csrc" / "host"

DRIVER = r'''
#include "read_set.hpp"
#include "set_parser.hpp"
#include <cstdio>
using namespace commet_host;
// argv: <fof> <out_dir> <suffix>: dumps every set's valid-read stream, tags every 3rd valid read (and the last one)
// and writes the per-file vectors as FileManager::save_bv does
int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    std::map<std::string, SetSpec> sets = read_sets(argv[1]);
    for (auto &kv : sets) {
        ReadSet rs;
        rs.nickname = kv.first;
        for (size_t i = 0; i < kv.second.files.size(); i++) rs.add_file(kv.second.files[i], kv.second.bvs[i]);
        rs.build_stream(false);
        printf("SET\t%s\t%zu\t%llu\n", kv.first.c_str(), rs.files.size(), (unsigned long long)rs.n_valid());
        for (uint64_t r = 0; r < rs.n_valid(); r++) {
            fwrite(rs.bases.data() + rs.offs[r], 1, rs.offs[r + 1] - rs.offs[r], stdout);
            fputc('\n', stdout);
        }
        for (auto &f : rs.files) printf("FILE\t%s\t%llu\t%zu\n", f.fname.c_str(), (unsigned long long)f.nb_reads, f.valid_pos.size());
        std::vector<uint8_t> tags(rs.n_valid() / 8 + 1, 0);
        for (uint64_t r = 0; r < rs.n_valid(); r++)
            if (r % 3 == 0 || r + 1 == rs.n_valid()) tags[r / 8] |= (uint8_t)(1u << (r % 8));
        rs.scatter_tags(tags);
        rs.save_bv(argv[2], std::string(argv[3]) + kv.first);
    }
    return 0;
}
'''


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostdrv")
    (d / "drv.cpp").write_text(DRIVER)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-I", str(HOST), "-o", str(d / "drv"), str(d / "drv.cpp"), "-lz"],
                   check=True)
    return d / "drv"


def _write_file(rng, path, reads, kind):
    if kind == 0:
        return H.write_fasta(path.with_suffix(".fa"), reads, final_newline=bool(rng.integers(0, 2)))
    if kind == 1:
        return H.write_fasta(path.with_suffix(".fa"), reads, width=int(rng.integers(5, 40)), final_newline=bool(rng.integers(0, 2)),
                             blank_every=int(rng.integers(0, 4)))
    if kind == 2:
        return H.write_fastq(path.with_suffix(".fq"), reads, final_newline=bool(rng.integers(0, 2)))
    if kind == 3:
        return H.write_fasta(path.with_suffix(".fa.gz"), reads, gz=True, width=int(rng.choice([0, 17])) or None)
    return H.write_fastq(path.with_suffix(".fq.gz"), reads, gz=True)


@pytest.mark.parametrize("seed", range(12))
def test_sets_streams_and_vectors(driver, tmp_path, seed):
    rng = np.random.default_rng(4000 + seed)
    n_sets = int(rng.integers(1, 4))
    lines, expect = [], {}
    for s in range(n_sets):
        name = ["zeta", "Alpha", "m 1", "b"][s] if seed % 2 else f"S{9 - s}"
        items, stream, files = [], [], []
        for fi in range(int(rng.integers(1, 4))):
            reads = H.make_ref_set(rng, int(rng.integers(1, 80)), 1, 70, p_N=0.03, p_lower=0.2)
            p = _write_file(rng, tmp_path / f"s{s}_f{fi}", reads, int(rng.integers(0, 5)))
            n = len(reads)
            if rng.integers(0, 2):
                valid = (rng.random(n) < 0.6).astype(np.uint8)
                bvp = tmp_path / f"s{s}_f{fi}.in.bv"
                oracle.write_bv_file(bvp, b"some input\nvector", n, oracle.tags_to_bv(valid))
                items.append(f" {p} , {bvp} " if seed % 3 == 0 else f"{p},{bvp}")
            else:
                valid = np.ones(n, dtype=np.uint8)
                items.append(f"  {p}" if seed % 3 == 0 else str(p))
            stream += [r for r, v in zip(reads, valid) if v]
            files.append((p, n, valid))
        lines.append((name + ":" if not (seed % 4 == 3 and s == 0) else "") + ";".join(items))
        expect[name if not (seed % 4 == 3 and s == 0) else "SET1"] = (stream, files)
    fof = tmp_path / "sets.txt"
    fof.write_text("\n".join(lines) + ("\n\n" if seed % 2 else ""))
    # the oracle's restatement of the grammar agrees on names, order, files and vectors
    parsed = oracle.parse_fof(fof)
    assert [p[0] for p in parsed] == sorted(expect)
    out = tmp_path / "bv"
    out.mkdir()
    r = subprocess.run([str(driver), str(fof), str(out), "in_"], capture_output=True, timeout=60)
    assert r.returncode == 0, r.stderr
    rows = r.stdout.split(b"\n")
    i = 0
    for name in sorted(expect):                    # std::map order = sorted names (set_parser.h:46)
        stream, files = expect[name]
        head = rows[i].split(b"\t")
        assert head[0] == b"SET" and head[1].decode() == name and int(head[2]) == len(files) and int(head[3]) == len(stream)
        assert rows[i + 1:i + 1 + len(stream)] == stream
        i += 1 + len(stream)
        # tags of every 3rd valid read (and the last) land on the reads' RECORD positions in their files
        tag = np.zeros(len(stream), dtype=np.uint8)
        tag[::3] = 1
        if len(stream):
            tag[-1] = 1
        s0 = 0
        for (p, n, valid) in files:
            frow = rows[i].split(b"\t")
            i += 1
            nv = int(valid.sum())
            assert frow[0] == b"FILE" and frow[1].decode() == str(p) and int(frow[2]) == n and int(frow[3]) == nv
            exp_bits = np.zeros(n, dtype=np.uint8)
            exp_bits[np.nonzero(valid)[0]] = tag[s0:s0 + nv]
            s0 += nv
            comment, nbits, payload = oracle.read_bv_file(out / f"{p.name}_in_in_{name}.bv")      # file_manager.h:247
            assert comment == f"{p} in in_{name}".encode() and nbits == n                       # :248
            assert np.array_equal(payload, oracle.tags_to_bv(exp_bits))


def test_vector_size_mismatch_and_unknown_format_follow_the_reference(driver, tmp_path):
    reads = [b"ACGT", b"GGCC"]
    fa = H.write_fasta(tmp_path / "a.fa", reads)
    bad = tmp_path / "bad.bv"
    oracle.write_bv_file(bad, b"x", 3, oracle.tags_to_bv(np.ones(3, dtype=np.uint8)))
    (tmp_path / "f1.txt").write_text(f"s:{fa},{bad}\n")
    (tmp_path / "o").mkdir()
    r = subprocess.run([str(driver), str(tmp_path / "f1.txt"), str(tmp_path / "o"), "x"], capture_output=True)
    assert r.returncode == 1 and b"boolean vector size are not equal -> quit" in r.stderr       # fasta_file.h:104-107
    junk = tmp_path / "junk.txt"
    junk.write_bytes(b"not a read file\n")
    (tmp_path / "f2.txt").write_text(f"s:{junk};{fa}\n")
    r = subprocess.run([str(driver), str(tmp_path / "f2.txt"), str(tmp_path / "o"), "x"], capture_output=True)
    # an unusable file is ignored with a message, the set goes on with the others (file_manager.h:150-156)
    assert r.returncode == 0 and b"-> ignore" in r.stderr
    assert r.stdout.split(b"\n")[0] == b"SET\ts\t1\t2"


def test_parallel_fasta_path_of_the_tools_equals_the_sequential_one(driver, tmp_path):
    """parse_reads_file switches to the mmap + chunked loader for large plain FASTA files: a 3-chunk file (> 32 MiB,
    multi-line records) must give the same stream through both paths (COMMET_B200_FAST_FASTA_MIN forces either)."""
    import hashlib
    import os
    rng = np.random.default_rng(77)
    n, L, width = 300_000, 120, 60
    arr = H.ACGT[rng.integers(0, 4, size=(n, L))]
    arr[rng.random(n) < 0.01, 33] = ord("N")
    with open(tmp_path / "big.fa", "wb") as f:
        lines = np.empty((n, 2, width + 1), dtype=np.uint8)
        lines[:, :, :width] = arr.reshape(n, 2, width)
        lines[:, :, width] = ord("\n")
        for s in range(0, n, 50_000):
            e = min(n, s + 50_000)
            f.write(b"".join(b">read%d\n" % i + lines[i].tobytes() for i in range(s, e)))
    assert (tmp_path / "big.fa").stat().st_size > (32 << 20)
    (tmp_path / "big.txt").write_text(f"big:{tmp_path / 'big.fa'}\n")
    outs = []
    for min_bytes in ("0", str(1 << 40)):
        o = tmp_path / f"o{len(outs)}"
        o.mkdir()
        r = subprocess.run([str(driver), str(tmp_path / "big.txt"), str(o), "x"], capture_output=True, timeout=120,
                           env={**os.environ, "COMMET_B200_FAST_FASTA_MIN": min_bytes})
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert hashlib.sha256(outs[0]).digest() == hashlib.sha256(outs[1]).digest()
    rows = outs[0].split(b"\n")
    assert rows[0] == b"SET\tbig\t1\t%d" % n
    assert rows[1] == arr[0].tobytes() and rows[n] == arr[n - 1].tobytes()


def test_encode_arithmetic_of_the_staging_kernels_on_every_byte(tmp_path):
    """encode_word / encode32 (commet_b200/csrc/kernels/staging.cuh: what k_encode and k_stage_filter compute per 32 bases) compiled
    for the host out of the kernel source: H, L and the validity bit of every byte value in every lane of a word, with
    arbitrary neighbours, against the per-character definition (hash_key.h:65-91 A=00 C=01 G=10 T=11; alphabet.h:44-58)."""
    import subprocess
    src = (ROOT / "commet_b200" / "csrc" / "kernels" / "staging.cuh").read_text()
    fn = src[src.index("__device__ __forceinline__ void encode_word("):src.index("__global__ void __launch_bounds__(256)\nk_encode(")]
    (tmp_path / "e.cpp").write_text(r'''
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#define __device__
#define __forceinline__ inline
struct uint4 { uint32_t x, y, z, w; };
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return (hi << s) | (lo >> (32 - s)); }
''' + fn + r'''
static bool is_in(uint8_t c) { return c && strchr("ACGTacgt", c) != nullptr; }
int main() {
    std::mt19937_64 rng(1);
    unsigned long long checked = 0;
    for (int round = 0; round < 2048; round++) {
        uint8_t b[32];
        for (int i = 0; i < 32; i++) {
            uint64_t r = rng();
            b[i] = round % 3 == 0 ? (uint8_t)r : (round % 3 == 1 ? "ACGTacgtNn-*"[r % 12] : "ACGT"[r % 4]);
        }
        for (int lane = 0; lane < 32; lane++)
            for (int v = 0; v < 256; v += (round < 8 ? 1 : 37)) {           // every value in every lane for the first rounds
                uint8_t s[32];
                memcpy(s, b, 32);
                s[lane] = (uint8_t)(v + (round < 8 ? 0 : round) & 255);
                uint4 q0, q1;
                memcpy(&q0, s, 16);
                memcpy(&q1, s + 16, 16);
                uint32_t H, L, V;
                encode32(q0, q1, H, L, V);
                for (int i = 0; i < 32; i++) {
                    const uint8_t c = s[i];
                    const uint32_t h = (c >> 2) & 1u, l = ((c >> 1) ^ (c >> 2)) & 1u, ok = is_in(c);
                    if (((H >> i) & 1u) != h || ((L >> i) & 1u) != l || ((V >> i) & 1u) != ok) {
                        printf("byte %d of the word = 0x%02x: H %u/%u L %u/%u V %u/%u\n", i, c, (H >> i) & 1u, h, (L >> i) & 1u, l, (V >> i) & 1u, ok);
                        return 1;
                    }
                    checked++;
                }
            }
    }
    printf("ok %llu\n", checked);
    return 0;
}''')
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(tmp_path / "e"), str(tmp_path / "e.cpp")], check=True)
    out = subprocess.run([str(tmp_path / "e")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout + out.stderr


def test_key_arithmetic_of_the_kernels_against_the_oracle_filter(tmp_path):
    """make_keys / key_word / key_bit (commet_b200/csrc/kernels/common.cuh: the four keys of a k-mer as windows of the bit-planes and
    their place in the filter) compiled for the host out of the kernel source: the filter they build from a read equals the
    oracle's (hash_key.h:65-91, bloom_filter.h:112-131), and the reverse keys of the reverse complement are the forward
    keys of the read, mirrored (hash_key.h:99-125)."""
    import subprocess
    import numpy as np
    from oracle import oracle
    src = (ROOT / "commet_b200" / "csrc" / "kernels" / "common.cuh").read_text()
    fn = src[src.index("struct Keys { uint64_t a, b, c, d; };"):src.rindex("}  // namespace commet")]
    (tmp_path / "k.cpp").write_text(r'''
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#define __device__
#define __forceinline__ inline
static inline uint64_t __brevll(uint64_t x) { uint64_t r = 0; for (int i = 0; i < 64; i++) r |= ((x >> i) & 1ull) << (63 - i); return r; }
''' + fn + r'''
static int code(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; default: return 3; } }
// plane windows of the k-mer starting at p: bit i = base p + i
static void planes(const std::string &s, size_t p, int k, uint64_t &hv, uint64_t &lv)
{
    hv = lv = 0;
    for (int i = 0; i < k; i++) { int c = code(s[p + i]); hv |= (uint64_t)(c >> 1) << i; lv |= (uint64_t)(c & 1) << i; }
}
int main(int argc, char **argv)
{
    const int k = atoi(argv[1]);
    const std::string s = argv[2];
    std::string rc(s.rbegin(), s.rend());
    for (char &c : rc) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    const uint64_t mask = k == 64 ? ~0ull : (1ull << k) - 1;
    std::vector<uint32_t> filt((((size_t)1 << (k - 1)) + 3) / 4, 0);
    const size_t n = s.size() - k + 1;
    for (size_t p = 0; p < n; p++) {
        uint64_t hv, lv;
        planes(s, p, k, hv, lv);
        Keys f = make_keys(hv, lv, k, mask, false);
        const uint64_t key[4] = {f.a, f.b, f.c, f.d};
        for (int j = 0; j < 4; j++) filt[key_word(key[j])] |= key_bit(key[j], j);
        planes(rc, n - 1 - p, k, hv, lv);
        Keys r = make_keys(hv, lv, k, mask, true);
        if (r.a != f.a || r.b != f.b || r.c != f.c || r.d != f.d) { printf("reverse keys differ at %zu\n", p); return 1; }
        if (k <= 32 && (key_word((uint32_t)f.a) != key_word(f.a) || key_bit((uint32_t)f.d, 3) != key_bit(f.d, 3))) { puts("32-bit overloads differ"); return 1; }
    }
    const uint8_t *b = reinterpret_cast<const uint8_t *>(filt.data());
    for (size_t i = 0; i < ((size_t)1 << (k - 1)); i++) printf("%02x", b[i]);
    puts("");
    return 0;
}''')
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(tmp_path / "k"), str(tmp_path / "k.cpp")], check=True)
    rng = np.random.default_rng(3)
    for k in (1, 2, 3, 4, 5, 8, 11, 13, 16, 17):
        for _ in range(3):
            seq = "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(k, k + 60))))
            out = subprocess.run([str(tmp_path / "k"), str(k), seq], capture_output=True, text=True)
            assert out.returncode == 0, out.stdout
            got = np.frombuffer(bytes.fromhex(out.stdout.strip()), dtype=np.uint8)
            want = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)
            oracle.index_chunk(want, k, np.frombuffer(seq.encode(), dtype=np.uint8), np.array([0, len(seq)], dtype=np.uint64), 0, 1 << 62)
            assert np.array_equal(got, want), (k, seq)
