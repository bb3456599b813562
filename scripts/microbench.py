"""Measured ceilings for the roofline: random 32-byte-sector loads / RED.OR over a DRAM-resident (4 GiB)
and an L2-resident (64 MiB) buffer. Run on the GPU box: python scripts/microbench.py > gpurun_out/ceilings.json"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import commet_b200


def main():
    ctx = commet_b200.Context(0)
    out = {}
    for name, nbytes in (("dram_4GiB", 1 << 32), ("l2_64MiB", 1 << 26), ("dram_2GiB", 1 << 31), ("l2_16MiB", 1 << 24)):
        for atomic in (False, True):
            best = 0.0
            for _ in range(3):
                best = max(best, ctx.random_sector_rate(nbytes, 1 << 30, atomic))
            out[f"{name}_{'redor' if atomic else 'load'}_Gsectors_s"] = round(best, 2)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
