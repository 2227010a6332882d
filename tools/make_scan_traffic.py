#!/usr/bin/env python3
"""profiles/scan_traffic.json (what bench.py's roofline object reads) from an `ncu --set full` report that holds the scan
launches of ONE discover_device call of bench.py (k_bin_scan + k_pair_scan[2]).
usage: python tools/make_scan_traffic.py gpurun_out/x.ncu-rep profiles/rN_binscan_ncu_summary.json [guides] [targets] [k]"""
import csv, io, json, subprocess, sys

rep, summary_path = sys.argv[1], sys.argv[2]
guides = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
targets = int(sys.argv[4]) if len(sys.argv) > 4 else 299989641
k = int(sys.argv[5]) if len(sys.argv) > 5 else 4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def num(vals, key):
    i = hdr.index(key)
    return float(vals[i].replace(",", "")) * SCALE.get(units[i], 1)


per, tot_ms, tot_dram, w_issue, w_alu = [], 0.0, 0.0, 0.0, 0.0
for vals in rows[2:]:
    if len(vals) < len(hdr):
        continue
    name = vals[hdr.index("Kernel Name")]
    if "k_bin_scan" not in name and "k_pair_scan" not in name:
        continue
    ms = num(vals, "gpu__time_duration.sum")
    dram = num(vals, "dram__bytes_read.sum") + num(vals, "dram__bytes_write.sum")
    issue = num(vals, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    alu = num(vals, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")
    stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(vals[i]), 3)
              for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(vals[i]) >= 0.3}
    per.append({"kernel": name.replace("void ", "").replace("ff::", "").replace("(int)", ""), "ms_under_ncu": ms, "dram_bytes": dram,
                "dram_pct_of_peak": num(vals, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "issue_active_pct": issue,
                "alu_pipe_pct": alu, "warp_instructions": num(vals, "smsp__inst_executed.sum"),
                "threads_per_instruction": num(vals, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                "l2_hit_pct": num(vals, "lts__t_sector_hit_rate.pct"), "registers": num(vals, "launch__registers_per_thread"),
                "warps_active_pct": num(vals, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "shared_wavefronts": num(vals, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                "stalls_per_issue_ge_0.3": stalls})
    tot_ms += ms; tot_dram += dram; w_issue += issue * ms; w_alu += alu * ms
out = {"kernel": "k_bin_scan+k_pair_scan", "targets": targets, "max_mismatch": k, "guides_in_profiled_launch": guides,
       "dram_bytes_per_profiled_launch": tot_dram, "bound": "issue", "issue_frac": w_issue / tot_ms / 100.0, "alu_pipe_frac": w_alu / tot_ms / 100.0,
       "per_launch": per,
       "source": "%s (ncu --set full --clock-control none; the scan launches of one %d-guide discover_device call of bench.py, summed)" % (summary_path, guides)}
json.dump(out, open("profiles/scan_traffic.json", "w"), indent=1)
json.dump(dict(out, source_report=rep + " (scratch, not committed)"), open(summary_path, "w"), indent=1)
print(json.dumps({"dram": tot_dram, "ms": tot_ms, "issue": out["issue_frac"], "alu": out["alu_pipe_frac"], "kernels": [p["kernel"] for p in per]}))
