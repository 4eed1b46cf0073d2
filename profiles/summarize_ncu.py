#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into two small tracked files:

    python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/rN_name

-> profiles/rN_name_metrics.csv   selected raw metrics of the first captured launch
-> profiles/rN_name_hotlines.csv  warp-stall samples and executed instructions per CUDA source line
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size",
    "launch__block_size", "smsp__average_warp_latency_per_inst_issued.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out + "_metrics.csv", "w") as f:
        f.write("metric,unit,value\n")
        f.write(f"kernel,,{vals[hdr.index('Kernel Name')]}\n")
        for i, h in enumerate(hdr):
            if h in KEEP or h.startswith("smsp__pcsamp_warps_issue_stalled"):
                f.write(f"{h},{units[i]},{vals[i]}\n")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[2] == "-" and r[0].isdigit():
            try:
                s, n = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
            except ValueError:
                continue
            if s or n:
                k = (cur, int(r[0]), r[1].strip())
                a = agg.get(k, (0, 0))
                agg[k] = (a[0] + s, a[1] + n)
    ts, tn = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
    with open(out + "_hotlines.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["file", "line", "stall_samples_pct", "warp_inst_pct", "source"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
            w.writerow([k[0], k[1], f"{100 * v[0] / ts:.2f}", f"{100 * v[1] / tn:.2f}", k[2][:140]])
    print("wrote", out + "_metrics.csv", out + "_hotlines.csv")


if __name__ == "__main__":
    main()
