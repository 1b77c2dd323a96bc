"""Prints the handful of ncu --set full metrics the roofline discussion uses, per captured kernel launch."""
import csv, subprocess, sys, io, os
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
# kernel -> the timeline group bench.py reports it under (b200phys_get_timeline)
GROUP = {"solveKernel": "solve", "npGjkKernel": "np_gjk", "npEpaSmallKernel": "np_epa", "npEpaSmKernel": "np_epa", "npClassifyKernel": "np_classify_sort", "bpFilterKernel": "bp_filter", "depRankKernel": "solver_order", "npEpaWarpKernel": "np_epa_redo", "npEpaKernel": "np_epa_redo",
         "bpRowsDirectKernel": "bp_tiles", "bpTileKernel": "bp_tiles", "bpEmitKernel": "bp_emit", "initContactRowsKernel": "solver_rows",
         "npCollideKernel": "np_closed_form_sat", "npCompactKernel": "np_compact"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(path, out_json):
    """DRAM bytes (read + write) per launch of every captured kernel, summed per timeline group -> the file bench.py reads."""
    import json, re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]; units = rows[1]
    per_kernel = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = re.sub(r"^void ", "", d["Kernel Name"]); name = re.sub(r"[<(].*", "", name)
        b = sum(float(d[k].replace(",", "")) * SCALE[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per_kernel.setdefault(name, []).append(b)
    groups = {}
    for name, v in per_kernel.items():
        g = GROUP.get(name)
        if g:
            groups[g] = groups.get(g, 0.0) + sum(v)          # the capture holds ONE step: several launches of a group add up
    groups["_source"] = "ncu --set full --clock-control none, one step of the bench.py headline workload; report summarised in " + os.path.basename(out_json).replace("_traffic.json", "") + "*"
    groups["_per_kernel"] = {k: sum(v) for k, v in per_kernel.items()}
    json.dump(groups, open(out_json, "w"), indent=1)


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]; units = rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:60], d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d and d[k] != "":
                print("   %-82s %s %s" % (k, d[k], u.get(k, "")))
if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
        traffic(sys.argv[1], sys.argv[3])
    else:
        main(sys.argv[1])
