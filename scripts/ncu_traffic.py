"""profiles/ncu_traffic.json (what bench.py's roofline.traffic reads) from the committed ncu --set full summary.

    python scripts/ncu_traffic.py profiles/r01_ncu_full_summary.csv profiles/ncu_traffic.json
"""
import csv
import json
import re
import sys

NOTES = {
    "k_search": "filter probes use ld.global.L2::64B: one 64-byte DRAM fill per missed probe instead of the default 128-byte fill",
}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    h = rows[0]
    col = {name.split(" [")[0]: i for i, name in enumerate(h)}
    unit = {name.split(" [")[0]: (name.split("[")[1].rstrip("]") if "[" in name else "") for name in h}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    kernels = {}
    for r in rows[1:]:
        m = re.search(r"(k_[a-z_0-9]+)", r[col["Kernel Name"]])
        if not m:
            continue
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[unit["dram__bytes_read.sum"]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[unit["dram__bytes_write.sum"]]
        e = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
             "ms": round(float(r[col["gpu__time_duration.sum"]]), 3), "kernel": r[col["Kernel Name"]].split("(")[0]}
        if m.group(1) in NOTES:
            e["note"] = NOTES[m.group(1)]
        kernels[m.group(1)] = e
    out = {"source": f"{src} (ncu --set full --clock-control none --import-source on, one launch each, "
                     "bench.py --steps 1 --warmup 0 --no-cpu --no-extra)",
           "workload": {"reads_per_set": 10000000, "read_len": 100, "k": 33, "t": 2}, "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
