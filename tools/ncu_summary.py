#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --set full` report -> JSON (+ a markdown table): duration, warp instructions, issue
utilisation, pipe utilisation, shared-memory bank conflicts, DRAM traffic, occupancy, top stall reasons.  Writes
profiles/<name>.json and profiles/<name>.md, and refreshes profiles/traffic.json (read by bench.py) from the k_core row.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep <images in the profiled launches> <name>
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0,
        "nsecond": 1e-6, "second": 1e3}
M = {
    "ms": "gpu__time_duration.sum",
    "warp_inst": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "pipe_fp64_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "pipe_tensor_pct": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_pct_of_peak": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
}


def main():
    rep, nimg, name = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    col = {c: i for i, c in enumerate(h)}
    res = []
    for r in rows[2:]:
        kn = r[col["Kernel Name"]]
        short = kn.split("(")[0].split("::")[-1].strip()
        d = {"kernel": short, "images_in_launch": nimg}
        for k, m in M.items():
            if m not in col or r[col[m]] in ("", "n/a"):
                d[k] = None
                continue
            v = float(r[col[m]].replace(",", ""))
            d[k] = v * UNIT.get(u[col[m]], 1.0)
        stalls = {}
        for c, i in col.items():
            if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio"):
                try:
                    stalls[c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        d["top_stalls_per_issue"] = {k: round(v, 2) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:5] if k != "selected"}
        if d.get("warp_inst"):
            d["warp_inst_per_image"] = d["warp_inst"] / nimg
        if d.get("dram_read_bytes") is not None and d.get("dram_write_bytes") is not None:
            d["dram_bytes_per_image"] = (d["dram_read_bytes"] + d["dram_write_bytes"]) / nimg
        res.append(d)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump({"report": os.path.basename(rep), "capture": "ncu --set full --clock-control none (per-launch times are cold-cache and serialised)",
               "kernels": res}, open(os.path.join(ROOT, "profiles", name + ".json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", name + ".md"), "w") as f:
        f.write(f"# {name}: ncu --set full, {nimg} images per launch ({os.path.basename(rep)})\n\n")
        f.write("| kernel | ms | warp-inst / image | issue % | warps % | regs | ALU % | FMA % | LSU % | XU % | FP64 % | tensor % | smem conflicts / wavefronts | DRAM B / image | top stalls (warps per issue) |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for d in res:
            def g(k, fmt="{:.1f}"):
                return "-" if d.get(k) is None else fmt.format(d[k])
            conf = "-" if not d.get("smem_wavefronts") else f"{d['smem_bank_conflicts'] / d['smem_wavefronts'] * 100:.1f} %"
            st = ", ".join(f"{k} {v}" for k, v in d["top_stalls_per_issue"].items())
            f.write(f"| {d['kernel']} | {g('ms', '{:.3f}')} | {g('warp_inst_per_image', '{:,.0f}')} | {g('issue_active_pct')} | {g('warps_active_pct')} | "
                    f"{g('regs', '{:.0f}')} | {g('pipe_alu_pct')} | {g('pipe_fma_pct')} | {g('pipe_lsu_pct')} | {g('pipe_xu_pct')} | {g('pipe_fp64_pct')} | "
                    f"{g('pipe_tensor_pct')} | {conf} | {g('dram_bytes_per_image', '{:,.0f}')} | {st} |\n")
    core = [d for d in res if d["kernel"].startswith("k_core")]
    if core:
        d = core[0]
        t = {"k_core_dram_bytes_per_image": d["dram_bytes_per_image"], "dram_read_bytes": d["dram_read_bytes"], "dram_write_bytes": d["dram_write_bytes"],
             "images_in_launch": nimg, "kernel_ms": d["ms"], "algorithmic_bytes_per_image": 664656, "report": os.path.basename(rep),
             "summary": f"profiles/{name}.json", "inst_executed": d["warp_inst"], "issue_active_pct": d["issue_active_pct"],
             "traffic_over_algorithmic": d["dram_bytes_per_image"] / 664656}
        json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(open(os.path.join(ROOT, "profiles", name + ".md")).read())


if __name__ == "__main__":
    main()
