#!/usr/bin/env python3
"""Key figures of one kernel from an .ncu-rep (needs `ncu` on PATH):  python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "sm__icc_request_hit_rate.pct", "sm__cycles_elapsed.max"]


def main():
    raw = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        print("==", d.get("Kernel Name", ("?",))[0][:100])
        for k in KEYS:
            if k in d:
                print("  %-62s %s %s" % (k, d[k][0], d[k][1]))
        st = sorted(((float(v[0]), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")),
                    reverse=True)
        print("  stalls per issue:", ", ".join("%s %.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], v) for v, h in st[:8]))


if __name__ == "__main__":
    main()
