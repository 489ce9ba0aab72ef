#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep
(needs -lineinfo at compile time and --import-source on at capture time).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_make_mask [top_n] [launch_skip]   (one launch is read)
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k",
                          f"regex:{kern}", "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
    csv.field_size_limit(10 ** 9)
    cur, hdr, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[0].isdigit():
            try:
                ie = int(r[hdr.index("Instructions Executed")])
                sm = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            agg.append((ie, sm, cur, int(r[0]), r[1].strip()[:100]))
    tot = sum(a[0] for a in agg) or 1
    ts = sum(a[1] for a in agg) or 1
    print(f"kernel {kern}: {tot} warp instructions, {ts} stall samples")
    print(" inst%  smpl%  file:line  source")
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{a[0] / tot * 100:5.1f}% {a[1] / ts * 100:5.1f}%  {a[2]}:{a[3]}  {a[4]}")


if __name__ == "__main__":
    main()
