#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into two small tracked files:

    python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/rN_name [--traffic profiles/traffic.json]

-> profiles/rN_name_metrics.csv   selected raw metrics of the first captured launch
-> profiles/rN_name_hotlines.csv  warp-stall samples and executed instructions per CUDA source line
-> (--traffic) the per-launch figures bench.py quotes in its `roofline` object, stamped with the hash of the kernel sources they
   were captured on (fc-planner_b200/build.py: kernel_build_id); bench.py prints traffic_stale when the sources have moved on
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
    if "--traffic" in sys.argv:
        import importlib.util
        import json
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location("fcvm_build", os.path.join(root, "fc-planner_b200", "build.py"))
        bm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bm)
        rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
        hdr, units, vals = rows[0], rows[1], rows[2]

        def val(name, default=None):
            if name not in hdr:
                return default
            v = vals[hdr.index(name)].replace(",", "")
            try:
                x = float(v)
            except ValueError:
                return default
            u = units[hdr.index(name)]
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1, "usecond": 1e-3, "msecond": 1, "second": 1e3, "nsecond": 1e-6}.get(u, 1)
            return x * scale
        stalls = sorted(((h[len("smsp__pcsamp_warps_issue_stalled_"):], float(vals[i].replace(",", "") or 0)) for i, h in enumerate(hdr)
                         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")), key=lambda kv: -kv[1])
        tot = sum(v for _, v in stalls) or 1
        t = {"kernel": vals[hdr.index("Kernel Name")], "kernel_build": bm.kernel_build_id(), "source": os.path.basename(out) + "_metrics.csv (ncu --set full --clock-control none)",
             "kernel_ms_under_ncu": val("gpu__time_duration.sum"),
             "dram_bytes_read_per_launch": val("dram__bytes_read.sum"), "dram_bytes_write_per_launch": val("dram__bytes_write.sum"),
             "dram_bytes_per_launch": (val("dram__bytes_read.sum", 0) + val("dram__bytes_write.sum", 0)),
             "l2_bytes_per_launch": val("lts__t_bytes.sum"), "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct"),
             "l1_sectors_global_ld_per_launch": val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), "l1_hit_rate_pct": val("l1tex__t_sector_hit_rate.pct"),
             "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps_per_scheduler": val("smsp__warps_active.avg.per_cycle_active"),
             "eligible_warps_per_scheduler": val("smsp__warps_eligible.avg.per_cycle_active"), "warp_instructions": val("smsp__inst_executed.sum"),
             "active_lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
             "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), "registers_per_thread": val("launch__registers_per_thread"),
             "occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
             "top_stalls_pct": {k: round(100 * v / tot, 1) for k, v in stalls[:6]}}
        path = sys.argv[sys.argv.index("--traffic") + 1]
        json.dump(t, open(path, "w"), indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
