#!/usr/bin/env python
"""profiles/traffic.json: measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant DP kernel,
read from `ncu --set full` reports brought back from the GPU box. bench.py reports it as roofline.traffic for the matching workload.

    python tools/ncu_traffic.py KEY=path/to/report.ncu-rep [KEY=...]     KEY = "<config>:<reads>:<haplotypes>:<band>:<dp|flank|ref>"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = head.index(name)
        total += float(vals[i].replace(",", "")) * UNIT[units[i]]
    k = head.index("Kernel Name")
    return total, vals[k]


def main():
    table = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            table = json.load(f)
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        b, kernel = dram_bytes(rep)
        table[key] = b
        table.setdefault("_sources", {})[key] = {"report": os.path.relpath(rep, ROOT), "kernel": kernel.split("(")[0]}
        print(key, "%.1f MB" % (b / 1e6), kernel.split("(")[0])
    with open(OUT, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
