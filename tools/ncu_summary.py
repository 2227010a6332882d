#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep, `ncu --set full`) into a small JSON under profiles/ (the reports themselves are scratch).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json "description" [key=value ...]"""
import csv, io, json, subprocess, sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def main():
    rep, out, desc = sys.argv[1], sys.argv[2], sys.argv[3]
    extra = dict(a.split("=", 1) for a in sys.argv[4:])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches, total_dram, total_ms, names = [], 0.0, 0.0, []
    for vals in rows[2:]:  # one row per profiled launch
        if len(vals) < len(hdr):
            continue
        m = {}
        for i, h in enumerate(hdr):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                m[h] = [vals[i], units[i]]
        kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""

        def num(k):
            v, u = m[k]
            f = float(v.replace(",", ""))
            return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "s": 1e3}.get(u, 1)
        total_dram += num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        total_ms += num("gpu__time_duration.sum")
        names.append(kname)
        launches.append({"kernel": kname, "metrics": m})
    summary = {"kernel": " + ".join(names), "description": desc, "source_report": rep + " (scratch, not committed)",
               "dram_bytes_per_launch": total_dram, "duration_ms_under_ncu": total_ms, "launches": launches}
    summary.update(extra)
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps({k: summary[k] for k in ("kernel", "dram_bytes_per_launch")}))


if __name__ == "__main__":
    main()
