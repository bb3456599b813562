"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/commet_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (commet_b200/) never does.

Besides the C restatement this file holds small pure-Python restatements of
the reference's *host formats* (FASTA/FASTQ record semantics of
include/fasta_file.h:143-185 and include/fastq_file.h:120-190, the .bv file of
include/boolean_vector.h:302-414, the fof grammar of include/set_parser.h:46-102)
used to turn files into the (bases, offsets) streams the C oracle consumes.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", str(HERE)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.commet_oracle_filter_bytes.restype = C.c_uint64
        _lib.commet_oracle_filter_bytes.argtypes = [C.c_int]
        _lib.commet_oracle_max_kmer.restype = C.c_uint64
        _lib.commet_oracle_max_kmer.argtypes = [C.c_int]
        _lib.commet_oracle_keys.restype = None
        _lib.commet_oracle_keys.argtypes = [_u8p, C.c_uint64, C.c_int, C.c_int, _u64p, _i64p]
        _lib.commet_oracle_index_chunk.restype = C.c_uint64
        _lib.commet_oracle_index_chunk.argtypes = [_u8p, C.c_int, _u8p, _u64p, C.c_uint64, C.c_uint64,
                                                   C.c_uint64, _u64p, _u64p]
        _lib.commet_oracle_search.restype = C.c_uint64
        _lib.commet_oracle_search.argtypes = [_u8p, C.c_int, C.c_int, _u8p, _u64p, C.c_uint64, _u8p,
                                              _u64p, _u64p, _u64p]
        _lib.commet_oracle_index_and_search.restype = C.c_int
        _lib.commet_oracle_index_and_search.argtypes = [
            C.c_int, C.c_int, C.c_uint64, _u8p, _u64p, C.c_uint64, C.c_int,
            C.POINTER(_u8p), C.POINTER(_u64p), _u64p, C.POINTER(_u8p), _u64p, _u64p, _u64p]
        _lib.commet_oracle_shannon.restype = C.c_float
        _lib.commet_oracle_shannon.argtypes = [_u8p, C.c_uint64]
        _lib.commet_oracle_filter_reads.restype = None
        _lib.commet_oracle_filter_reads.argtypes = [_u8p, _u64p, C.c_uint64, C.c_int64, C.c_int64,
                                                    C.c_float, C.c_int64, _u8p, _u64p]
        _lib.commet_oracle_bvop.restype = None
        _lib.commet_oracle_bvop.argtypes = [C.c_int, _u8p, _u8p, _u8p, C.c_uint64]
        _lib.commet_oracle_nb_one.restype = C.c_uint64
        _lib.commet_oracle_nb_one.argtypes = [_u8p, C.c_uint64]
        _lib.commet_oracle_bv_init_true.restype = None
        _lib.commet_oracle_bv_init_true.argtypes = [_u8p, C.c_uint64]
    return _lib


def _p8(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def _p64(a: np.ndarray):
    return a.ctypes.data_as(_u64p)


def _stream(bases, offs):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    return bases, offs


# --------------------------------------------------------------------------
# C oracle front-ends
# --------------------------------------------------------------------------
def filter_bytes(k: int) -> int:
    return int(lib().commet_oracle_filter_bytes(k))


def max_kmer(k: int) -> int:
    return int(lib().commet_oracle_max_kmer(k))


def keys(seq: bytes, k: int, reverse: bool = False):
    """(keys[len,4] u64, size[len] i64) after each char; HashKey add/rv_add."""
    s = np.frombuffer(seq, dtype=np.uint8).copy()
    out = np.zeros((max(len(seq), 1), 4), dtype=np.uint64)
    size = np.zeros(max(len(seq), 1), dtype=np.int64)
    lib().commet_oracle_keys(_p8(s if s.size else np.zeros(1, np.uint8)), len(seq), k, int(reverse),
                             _p64(out), size.ctypes.data_as(_i64p))
    return out[:len(seq)], size[:len(seq)]


def index_chunk(filt: np.ndarray, k: int, bases, offs, start: int, maxk: int):
    """Index one chunk into `filt` (in place). Returns (next_start, n_indexed_reads, n_kmers)."""
    bases, offs = _stream(bases, offs)
    n = len(offs) - 1
    ni = C.c_uint64(0)
    nk = C.c_uint64(0)
    nxt = lib().commet_oracle_index_chunk(_p8(filt), k, _p8(bases), _p64(offs), n, start, maxk,
                                          C.byref(ni), C.byref(nk))
    return int(nxt), int(ni.value), int(nk.value)


def search(filt: np.ndarray, k: int, t: int, bases, offs, tags: np.ndarray):
    """search_reads on one filter. tags (u8 per read) updated in place.
    Returns dict(found, searched, tests, lookups)."""
    bases, offs = _stream(bases, offs)
    n = len(offs) - 1
    ns, nt, nl = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    found = lib().commet_oracle_search(_p8(filt), k, t, _p8(bases), _p64(offs), n, _p8(tags),
                                       C.byref(ns), C.byref(nt), C.byref(nl))
    return dict(found=int(found), searched=int(ns.value), tests=int(nt.value), lookups=int(nl.value))


def index_and_search(k: int, t: int, index_stream, query_streams, maxk: int | None = None):
    """Full chunk loop. index_stream=(bases, offs); query_streams=[(bases, offs), ...].
    Returns (tags list of u8 arrays, info dict)."""
    if maxk is None:
        maxk = max_kmer(k)
    ib, io = _stream(*index_stream)
    qs = [_stream(b, o) for b, o in query_streams]
    ns = len(qs)
    tags = [np.zeros(max(len(o) - 1, 1), dtype=np.uint8) for _, o in qs]
    qb_arr = (_u8p * ns)(*[_p8(b) for b, _ in qs])
    qo_arr = (_u64p * ns)(*[_p64(o) for _, o in qs])
    tg_arr = (_u8p * ns)(*[_p8(tg) for tg in tags])
    nq = np.array([len(o) - 1 for _, o in qs], dtype=np.uint64)
    searched = np.zeros(ns, dtype=np.uint64)
    shared = np.zeros(ns, dtype=np.uint64)
    stats = np.zeros(5, dtype=np.uint64)
    rc = lib().commet_oracle_index_and_search(k, t, maxk, _p8(ib), _p64(io), len(io) - 1, ns,
                                              qb_arr, qo_arr, _p64(nq), tg_arr,
                                              _p64(searched), _p64(shared), _p64(stats))
    if rc != 0:
        raise MemoryError("oracle filter allocation failed")
    tags = [tg[:len(o) - 1] for tg, (_, o) in zip(tags, qs)]
    info = dict(chunks=int(stats[0]), indexed=int(stats[1]), tests=int(stats[2]), lookups=int(stats[3]),
                kmers=int(stats[4]), searched=[int(x) for x in searched], shared=[int(x) for x in shared])
    return tags, info


def shannon(seq: bytes) -> float:
    s = np.frombuffer(seq, dtype=np.uint8).copy()
    return float(lib().commet_oracle_shannon(_p8(s), len(seq)))


def filter_reads(bases, offs, min_len=0, max_N=-1, min_shannon=0.0, max_reads=-1):
    """Returns (bv payload u8[n//8+1], counters dict)."""
    bases, offs = _stream(bases, offs)
    n = len(offs) - 1
    bv = np.zeros(n // 8 + 1, dtype=np.uint8)
    cnt = np.zeros(4, dtype=np.uint64)
    lib().commet_oracle_filter_reads(_p8(bases), _p64(offs), n, min_len, max_N,
                                     C.c_float(min_shannon), max_reads, _p8(bv), _p64(cnt))
    return bv, dict(rm_length=int(cnt[0]), rm_N=int(cnt[1]), rm_shannon=int(cnt[2]), selected=int(cnt[3]))


BV_AND, BV_OR, BV_ANDNOT, BV_NOT = 0, 1, 2, 3


def bvop(op: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty_like(a)
    lib().commet_oracle_bvop(op, _p8(a), _p8(b), _p8(out), a.size)
    return out


def nb_one(bv: np.ndarray, n_bits: int) -> int:
    bv = np.ascontiguousarray(bv, dtype=np.uint8)
    return int(lib().commet_oracle_nb_one(_p8(bv), n_bits))


def bv_init_true(n: int) -> np.ndarray:
    bv = np.empty(n // 8 + 1, dtype=np.uint8)
    lib().commet_oracle_bv_init_true(_p8(bv), n)
    return bv


def tags_to_bv(tags: np.ndarray) -> np.ndarray:
    """byte-per-read tags -> .bv payload (n/8+1 bytes, LSB-first)."""
    n = len(tags)
    out = np.zeros(n // 8 + 1, dtype=np.uint8)
    packed = np.packbits(tags.astype(np.uint8), bitorder="little")
    out[:len(packed)] = packed
    return out


def bv_to_tags(bv: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(bv, dtype=np.uint8), bitorder="little")[:n]


# --------------------------------------------------------------------------
# Host formats (pure Python, small inputs)
# --------------------------------------------------------------------------
def read_bv_file(path) -> tuple[bytes, int, np.ndarray]:
    """include/boolean_vector.h:347-414 -> (comment, n_bits, payload)."""
    raw = Path(path).read_bytes()
    h = raw.index(b"#")
    comment = raw[:h][:-1]            # drop the '\n' before '#'
    e = raw.index(b"\n", h)
    n = int(raw[h + 1:e])
    nb = n // 8 + 1
    payload = np.frombuffer(raw[e + 1:e + 1 + nb], dtype=np.uint8).copy()
    return comment, n, payload


def write_bv_file(path, comment: bytes, n: int, payload: np.ndarray) -> None:
    """include/boolean_vector.h:302-346: comment + "\\n#" + n + "\\n" + payload."""
    with open(path, "wb") as f:
        f.write(comment + b"\n#" + str(n).encode() + b"\n" + bytes(payload[:n // 8 + 1]))
    os.chmod(path, 0o600)


def _open_text(path):
    with open(path, "rb") as f:
        first = f.read(1)
    if first in (b">", b"@"):
        return open(path, "rb")
    return gzip.open(path, "rb")


def parse_reads(path) -> list[bytes]:
    """All records of a FASTA/FASTQ(.gz) file, reference semantics:
    FASTA: sequence = concatenation of the non-empty lines up to the next '>'
    line (fasta_file.h:166-175).  FASTQ: 4 non-empty lines per record, the
    sequence is the line after the '@' line (fastq_file.h:132-180)."""
    with _open_text(path) as f:
        data = f.read()
    lines = data.split(b"\n")
    if data[:1] == b">":
        reads, cur = [], None
        for ln in lines:
            if ln[:1] == b">":
                if cur is not None:
                    reads.append(b"".join(cur))
                cur = []
            elif ln and cur is not None:
                cur.append(ln)
        if cur is not None:
            reads.append(b"".join(cur))
        return reads
    ne = [ln for ln in lines if ln]
    return [ne[4 * i + 1] for i in range(len(ne) // 4)]


def to_stream(reads: list[bytes]):
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        offs[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy()
    return bases, offs


def parse_fof(path) -> list[tuple[str, list[str], list[str]]]:
    """include/set_parser.h:46-102 -> [(name, files, bvs)] sorted by name
    (std::map order).  Name is NOT trimmed; files and bvs are space-trimmed."""
    sets = {}
    nb = 0
    for line in Path(path).read_text().split("\n"):
        if not line:
            continue
        nb += 1
        if ":" in line:
            name, line = line.split(":", 1)
        else:
            name = f"SET{nb}"
        files, bvs = [], []
        for item in line.split(";"):
            item = item.strip(" ")
            bv = ""
            if "," in item:
                item, bv = item.split(",", 1)
                item, bv = item.strip(" "), bv.strip(" ")
            files.append(item)
            bvs.append(bv)
        sets[name] = (files, bvs)
    return [(n, sets[n][0], sets[n][1]) for n in sorted(sets)]


def have_ref() -> bool:
    return all((REF_DIR / t).exists() for t in ("index_and_search", "filter_reads", "bvop"))
