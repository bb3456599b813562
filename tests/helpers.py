"""Shared generators for the parity tests (seeded, numpy only)."""
from __future__ import annotations

import gzip
from pathlib import Path

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.full(256, ord("N"), dtype=np.uint8)   # NUL bytes would truncate gzgets-read lines
for a, b in zip(b"ACGTacgtN", b"TGCAtgcaN"):
    COMP[a] = b


def random_read(rng, length: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=length)]


def revcomp(r: np.ndarray) -> np.ndarray:
    return COMP[r[::-1]]


def mutate(rng, r: np.ndarray, rate: float) -> np.ndarray:
    r = r.copy()
    m = rng.random(len(r)) < rate
    r[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
    return r


def dirty(rng, r: np.ndarray, p_N=0.0, p_lower=0.0, p_other=0.0) -> np.ndarray:
    """Sprinkle N, lowercase and arbitrary bytes (0x30..0x7f: bytes >= 0x80 index the reference LUT out of bounds = UB; never newline/'>')."""
    r = r.copy()
    if p_lower:
        m = rng.random(len(r)) < p_lower
        r[m] |= 0x20
    if p_N:
        m = rng.random(len(r)) < p_N
        r[m] = ord("N")
    if p_other:
        m = rng.random(len(r)) < p_other
        junk = rng.integers(0x30, 0x80, size=int(m.sum())).astype(np.uint8)
        junk[junk == ord(">")] = ord("n")
        junk[junk == ord("@")] = ord("x")
        r[m] = junk
    return r


def make_ref_set(rng, n: int, len_lo: int, len_hi: int | None = None, **dirt) -> list[bytes]:
    len_hi = len_lo if len_hi is None else len_hi
    out = []
    for _ in range(n):
        L = int(rng.integers(len_lo, len_hi + 1))
        out.append(dirty(rng, random_read(rng, L), **dirt).tobytes())
    return out


def make_query_set(rng, ref: list[bytes], n: int, len_lo: int, len_hi: int | None = None,
                   frac_shared=0.5, sub_rate=0.01, **dirt) -> list[bytes]:
    """frac_shared copies of reference reads (half reverse-complemented,
    sub_rate substitutions), the rest fresh random reads."""
    len_hi = len_lo if len_hi is None else len_hi
    out = []
    for _ in range(n):
        if ref and rng.random() < frac_shared:
            r = np.frombuffer(ref[int(rng.integers(0, len(ref)))], dtype=np.uint8)
            if rng.random() < 0.5:
                r = revcomp(r)
            r = mutate(rng, r, sub_rate)
        else:
            r = random_read(rng, int(rng.integers(len_lo, len_hi + 1)))
        out.append(dirty(rng, r, **dirt).tobytes())
    return out


def to_stream(reads: list[bytes]):
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        offs[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy()
    return bases, offs


def _open_w(path, gz):
    return gzip.open(path, "wb") if gz else open(path, "wb")


def write_fasta(path, reads, width: int | None = None, gz=False, final_newline=True, blank_every=0):
    with _open_w(path, gz) as f:
        for i, r in enumerate(reads):
            f.write(b">r%d some comment\n" % i)
            if width:
                chunks = [r[j:j + width] for j in range(0, len(r), width)]
            else:
                chunks = [r]
            body = b"\n".join(chunks)
            last = i == len(reads) - 1
            f.write(body + (b"" if (last and not final_newline) else b"\n"))
            if blank_every and i % blank_every == 0 and not last:
                f.write(b"\n")
    return Path(path)


def write_fastq(path, reads, gz=False, final_newline=True):
    with _open_w(path, gz) as f:
        for i, r in enumerate(reads):
            last = i == len(reads) - 1
            f.write(b"@r%d\n" % i + r + b"\n+\n" + b"I" * len(r) + (b"" if (last and not final_newline) else b"\n"))
    return Path(path)
