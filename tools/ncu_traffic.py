#!/usr/bin/env python3
"""DRAM traffic of k_core from an `ncu --set full` report -> profiles/traffic.json (read by bench.py's roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep <images in the profiled launch>
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    rep, nimg = sys.argv[1], int(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    r = [x for x in rows[2:] if "k_core" in x[h.index("Kernel Name")]][0]
    rd = to_bytes(r[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_read.sum")])
    wr = to_bytes(r[h.index("dram__bytes_write.sum")], u[h.index("dram__bytes_write.sum")])
    d = {"k_core_dram_bytes_per_image": (rd + wr) / nimg, "dram_read_bytes": rd, "dram_write_bytes": wr, "images_in_launch": nimg,
         "kernel_us": float(r[h.index("gpu__time_duration.sum")].replace(",", "")), "algorithmic_bytes_per_image": 664656,
         "report": os.path.basename(rep), "inst_executed": float(r[h.index("smsp__inst_executed.sum")].replace(",", "")),
         "issue_active_pct": float(r[h.index("smsp__issue_active.avg.pct_of_peak_sustained_active")])}
    d["traffic_over_algorithmic"] = d["k_core_dram_bytes_per_image"] / 664656
    json.dump(d, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
