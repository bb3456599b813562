"""BASELINE.json configs[4] (C5): filter_reads (Shannon / N / min-length) and bvop AND / popcount as a
bandwidth-bound sweep with device-resident inputs.  GPU box only:

    python scripts/sweep_c5.py [--reads 16000000] [--batches 4] [--bits 1000000000] > gpurun_out/c5.json

filter_reads: SURVEY 8(d) mix -- lengths uniform in 50..150, 5 % low-complexity reads (homopolymer / dinucleotide),
5 % of reads with 1..10 N, options -l 66 -n 2 -e 1.5.  Reads are generated on the device per batch (seeded) and
streamed through three paths: selection only (k_stage_filter<false>: the ASCII bases read once, bits out), staging and
selection fused (k_stage_filter<true>: bases read once, bit-planes AND bits out), and the two-kernel path (k_encode,
then k_filter over the planes); a 1e9-read run is `--batches 63`.  Each path is timed as a CALL (with its counter
read-backs and synchronisations) -- the kernels alone are in the ncu launch list of the same command.
Algorithmic bytes: 1 B/base in (ASCII) for stage+filter; bvop AND = 3 B per payload byte, popcount = 1 B.
One batch is checked against the CPU oracle on a 200k-read prefix (bit-exact) before timing.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import commet_b200  # noqa: E402


def make_batch(n, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lens = torch.randint(50, 151, (n,), generator=g, device=dev)
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens, 0)
    total = int(offs[-1])
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = acgt[torch.randint(0, 4, (total,), generator=g, device=dev)]
    kind = torch.rand(n, generator=g, device=dev)
    rep = lambda v: torch.repeat_interleave(v, lens)            # per-read value -> per-base value
    # 5 % low complexity: half homopolymer, half dinucleotide repeat
    b0 = rep(acgt[torch.randint(0, 4, (n,), generator=g, device=dev)])
    b1 = rep(acgt[torch.randint(0, 4, (n,), generator=g, device=dev)])
    odd = ((torch.arange(total, device=dev) - rep(offs[:-1])) & 1).bool()
    bases = torch.where(rep(kind < 0.025), b0, bases)
    bases = torch.where(rep((kind >= 0.025) & (kind < 0.05)), torch.where(odd, b1, b0), bases)
    del b0, b1, odd
    # 5 % of reads carry 1..10 N at random positions
    n_count = torch.randint(1, 11, (n,), generator=g, device=dev)
    p_n = rep(((kind >= 0.05) & (kind < 0.10)).float() * n_count.float() / lens.float())
    bases = torch.where(torch.rand(total, generator=g, device=dev) < p_n, torch.full_like(bases, ord("N")), bases)
    del p_n
    return bases.contiguous(), offs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=16_000_000, help="reads per batch")
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--bits", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = commet_b200.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"hbm_peak_GBps": peak}

    # ---- filter_reads ------------------------------------------------------------------------------------
    opts = dict(min_len=66, max_N=2, min_shannon=1.5)
    t_stage = t_filter = t_fused = t_both = 0.0
    n_bases = n_reads = selected = 0
    for b in range(args.batches):
        bases, offs = make_batch(args.reads, 7000 + b, dev)
        n, nb = args.reads, int(offs[-1])
        if bases.numel() % 16:                               # the fused kernels read whole 16-byte vectors
            bases = torch.cat([bases, torch.zeros(16 - bases.numel() % 16, dtype=torch.uint8, device=dev)])
        d_bv = torch.zeros((n // 8 + 1 + 3) // 4, dtype=torch.int32, device=dev)
        d_bv2 = torch.zeros_like(d_bv)
        d_bv3 = torch.zeros_like(d_bv)
        torch.cuda.synchronize()
        if b == 0:      # parity on a prefix, through the same entry points
            from oracle import oracle
            m = 200_000
            hb, ho = bases[:int(offs[m])].cpu().numpy(), offs[:m + 1].cpu().numpy().astype(np.uint64)
            e_bv, e_cnt = oracle.filter_reads(hb, ho, **opts)
            g_bv, g_cnt = ctx.filter_reads(hb, ho, **opts)
            assert np.array_equal(e_bv, g_bv) and e_cnt == g_cnt, "filter_reads differs from the oracle"
            out["filter_parity_prefix"] = {"reads": m, **g_cnt}
            rs = ctx.stage_device(bases.data_ptr(), offs.data_ptr(), n, nb)     # warm-up
            ctx.filter_reads_staged(rs, d_bv.data_ptr(), **opts)
            rs.free()
            ctx.filter_reads_device(bases.data_ptr(), offs.data_ptr(), n, d_bv2.data_ptr(), **opts)
            rs, _ = ctx.stage_device_filtered(bases.data_ptr(), offs.data_ptr(), n, nb, d_bv3.data_ptr(), **opts)
            rs.free()
            ctx.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        with torch.cuda.stream(ext):
            ev[0].record()
        rs = ctx.stage_device(bases.data_ptr(), offs.data_ptr(), n, nb)
        with torch.cuda.stream(ext):
            ev[1].record()
        cnt = ctx.filter_reads_staged(rs, d_bv.data_ptr(), **opts)
        with torch.cuda.stream(ext):
            ev[2].record()
        cnt_f = ctx.filter_reads_device(bases.data_ptr(), offs.data_ptr(), n, d_bv2.data_ptr(), **opts)
        rs.free()                                              # its planes go back to the context's cache: the fused pass takes them
        with torch.cuda.stream(ext):
            ev[3].record()
        rs3, cnt_b = ctx.stage_device_filtered(bases.data_ptr(), offs.data_ptr(), n, nb, d_bv3.data_ptr(), **opts)
        with torch.cuda.stream(ext):
            ev[4].record()
        ctx.sync()
        torch.cuda.synchronize()
        assert cnt_f == cnt and torch.equal(d_bv, d_bv2), "selection-only pass and staged selection differ"
        assert cnt_b == cnt and torch.equal(d_bv, d_bv3), "fused staging+selection and staged selection differ"
        t_stage += ev[0].elapsed_time(ev[1]) * 1e-3
        t_filter += ev[1].elapsed_time(ev[2]) * 1e-3
        t_fused += ev[2].elapsed_time(ev[3]) * 1e-3
        t_both += ev[3].elapsed_time(ev[4]) * 1e-3
        n_bases += nb
        n_reads += n
        selected += cnt["selected"]
        rs3.free()
        del bases, offs, d_bv, d_bv2, d_bv3
    tot = t_stage + t_filter
    out["filter_reads_selection_only"] = {
        "reads": n_reads, "bases": n_bases, "selected": selected, "options": "-l 66 -n 2 -e 1.5", "ms": t_fused * 1e3,
        "reads_per_s": n_reads / t_fused, "algorithmic_GBps": n_bases / t_fused / 1e9, "frac_of_hbm_peak": n_bases / t_fused / 1e9 / peak,
        "note": "commet_filter_reads_dev: k_stage_filter<false> reads the ASCII once (1 B/base + 8 B/read of offsets) and writes 1 bit + 1 "
                "class byte per read; the call includes the undecided-read round trip, the cutoff kernel and the counter read-back"}
    out["filter_reads_fused_with_staging"] = {
        "reads": n_reads, "bases": n_bases, "selected": selected, "options": "-l 66 -n 2 -e 1.5", "ms": t_both * 1e3,
        "reads_per_s": n_reads / t_both, "algorithmic_GBps": n_bases / t_both / 1e9, "frac_of_hbm_peak": n_bases / t_both / 1e9 / peak,
        "traffic_GBps": 1.5 * n_bases / t_both / 1e9, "traffic_frac_of_hbm_peak": 1.5 * n_bases / t_both / 1e9 / peak,
        "note": "commet_reads_from_device_filtered: k_stage_filter<true> reads the ASCII once and writes the bit-planes (0.5 B/base) AND "
                "the selection bits: a set that is filtered and indexed is read from HBM once; traffic = 1.5 B/base"}
    out["filter_reads_two_kernels"] = {
        "reads": n_reads, "bases": n_bases, "selected": selected, "options": "-l 66 -n 2 -e 1.5",
        "stage_ms": t_stage * 1e3, "filter_ms": t_filter * 1e3,
        "reads_per_s": n_reads / tot, "algorithmic_GBps": n_bases / tot / 1e9, "frac_of_hbm_peak": n_bases / tot / 1e9 / peak,
        "stage_only_GBps": n_bases / t_stage / 1e9, "filter_only_reads_per_s": n_reads / t_filter,
        "note": "device-resident ASCII in, 1 bit/read out; stage = k_encode (1 B/base read + 0.5 B/base written), filter = k_filter "
                "over the planes (0.5 B/base) + cutoff; times are CUDA events on the context's stream"}

    # ---- bvop / popcount ---------------------------------------------------------------------------------
    nbytes = args.bits // 8 + 1
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    pad = (nbytes + 15) // 16 * 16
    a = torch.randint(0, 256, (pad,), dtype=torch.uint8, generator=g, device=dev)
    b = torch.randint(0, 256, (pad,), dtype=torch.uint8, generator=g, device=dev)
    o = torch.empty_like(a)
    torch.cuda.synchronize()
    res = {}
    for name, op in (("and", commet_b200.BV_AND), ("or", commet_b200.BV_OR), ("andnot", commet_b200.BV_ANDNOT), ("not", commet_b200.BV_NOT)):
        ctx.bvop_device(op, a.data_ptr(), b.data_ptr(), o.data_ptr(), nbytes)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        for _ in range(args.reps):
            ctx.bvop_device(op, a.data_ptr(), b.data_ptr(), o.data_ptr(), nbytes)
        with torch.cuda.stream(ext):
            e1.record()
        ctx.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        traffic = nbytes * (2 if name == "not" else 3)
        res[name] = {"ms": ms, "algorithmic_GBps": traffic / ms / 1e6, "frac_of_hbm_peak": traffic / ms / 1e6 / peak}
    exp = (a[:nbytes] & b[:nbytes])
    ctx.bvop_device(commet_b200.BV_AND, a.data_ptr(), b.data_ptr(), o.data_ptr(), nbytes)
    ctx.sync()
    assert torch.equal(o[:nbytes], exp), "bvop AND differs from torch"
    ones = ctx.nb_one_device(o.data_ptr(), args.bits)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=dev)
    assert ones == min(int(lut[o[:nbytes].long()].sum()), args.bits), "popcount differs from torch"
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ctx.nb_one_device(o.data_ptr(), args.bits)
    ms = (time.perf_counter() - t0) / args.reps * 1e3
    res["popcount"] = {"ms": ms, "algorithmic_GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                       "note": "includes the 8-byte D2H of the count and a stream sync per call"}
    nvec = 16
    t0 = time.perf_counter()
    for _ in range(args.reps):
        many = ctx.nb_one_device_batch([o.data_ptr()] * nvec, [args.bits] * nvec)
    ms = (time.perf_counter() - t0) / args.reps / nvec * 1e3
    assert many == [ones] * nvec
    res["popcount_batched"] = {"ms_per_vector": ms, "vectors_per_call": nvec, "algorithmic_GBps": nbytes / ms / 1e6,
                               "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                               "note": "commet_bv_popcount_batch_dev: one kernel per vector, one read-back and one sync per call"}
    out["bvop"] = {"bits": args.bits, "payload_bytes": nbytes, **res,
                   "note": "one 1e9-bit vector = 125 MB: a binary op touches 375 MB (> 126 MB L2)"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
