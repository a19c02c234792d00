#!/usr/bin/env python
"""Record the DRAM traffic of captured kernel launches in profiles/ncu_traffic.json, the file bench.py's `roofline.traffic`
reads (key = "<kernel name as the engine reports it>@<voices>").

    python tools/ncu_traffic.py <raw.csv[.gz] from `ncu -i rep --page raw --csv`> "<key>" "<source label>" [launch index]

dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum of that one launch."""
import csv
import gzip
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    path, key, source = sys.argv[1:4]
    idx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hdr, units, row = rows[0], rows[1], rows[2 + idx]
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        total += float(row[i].replace(",", "")) * UNIT[units[i]]
    out = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(out)) if os.path.exists(out) else {}
    def metric(name):
        return float(row[hdr.index(name)].replace(",", "")) if name in hdr else None
    table[key] = {"dram_bytes": int(total), "kernel": row[hdr.index("Kernel Name")], "time": row[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
                  "issue_active_pct": metric("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                  "warp_instructions": metric("smsp__inst_executed.sum"),
                  "warps_active_pct": metric("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "registers_per_thread": metric("launch__registers_per_thread"),
                  "source": source}
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)
    print(key, int(total), table[key]["kernel"])


if __name__ == "__main__":
    main()
