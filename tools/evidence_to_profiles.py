#!/usr/bin/env python3
"""Turns gpurun_out/evidence/ (tools/collect_evidence.sh) into the tracked files under profiles/: per-kernel ncu
summaries + hot source lines, the launch list, the bench lines, and profiles/traffic.json keyed the way bench.py looks
it up ("<kernel name as irlosc_last_kernel prints it>|<workload>|B<batch>").

    python tools/evidence_to_profiles.py r02 --on-box     # on the GPU box: summarise the .ncu-rep files next to them
                                                          # (summaries + traffic.json) and delete the reports (64 MiB cap)
    python tools/evidence_to_profiles.py r02              # here: copy the small files into profiles/
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "evidence")
PROF = os.path.join(ROOT, "profiles")


def raw_metrics(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    return {h: rows[2][i] for i, h in enumerate(rows[0])}


def printed_name(cxx):
    """`osc_step_lane<3, 1, 224, 1>` -> the string the library reports (irlosc_last_kernel)."""
    m = re.search(r"osc_step_lane<\(?(?:int\))?(\d+), \(?(?:bool\))?(\d+|true|false), \(?(?:int\))?(\d+), \(?(?:bool\))?(\d+|true|false)>", cxx)
    if m:
        kd, hb, nt, st = m.groups()
        hb = hb in ("1", "true")
        st = st in ("1", "true")
        return "osc_step_lane<kd%s%s,t%s,%s>" % (kd, ",base" if hb else "", nt, "tma" if st else "ldg")
    m = re.search(r"osc_step_pair<\(?(?:int\))?(\d+), \(?(?:bool\))?(\d+|true|false), \(?(?:int\))?(\d+)>", cxx)
    if m:
        kd, hb, nt = m.groups()
        return "osc_step_pair<kd%s%s,t%s>" % (kd, ",base" if hb in ("1", "true") else "", nt)
    return cxx


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    on_box = "--on-box" in sys.argv
    global PROF
    if on_box:
        PROF = EV
    traffic_path = os.path.join(PROF, "traffic.json")
    traffic = {}
    for f in sorted(os.listdir(EV)):
        src = os.path.join(EV, f)
        if not f.startswith(tag):
            continue
        if f.endswith(".ncu-rep"):
            base = f[:-len(".ncu-rep")]
            summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), src], capture_output=True, text=True).stdout
            lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src, "30"], capture_output=True, text=True).stdout
            with open(os.path.join(PROF, base + "_ncu_summary.txt"), "w") as fh:
                fh.write("# ncu --set full --clock-control none, one launch (tools/collect_evidence.sh); never a bench value\n")
                fh.write(summ + "\n" + lines)
            m = raw_metrics(src)
            wl = re.search(r"_(gain_test|admit_test|worst_case)$", base)
            if wl and ("osc_step_lane" in m.get("Kernel Name", "") or "osc_step_pair" in m.get("Kernel Name", "")):
                rd = float(m["dram__bytes_read.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get("Mbyte", 1e6)
                # units of the raw page: read them from the second header row instead of guessing
                raw = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
                rows = list(csv.reader(io.StringIO(raw)))
                unit = {h: rows[1][i] for i, h in enumerate(rows[0])}
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(m["dram__bytes_read.sum"]) * scale[unit["dram__bytes_read.sum"]]
                wr = float(m["dram__bytes_write.sum"]) * scale[unit["dram__bytes_write.sum"]]
                grid, block = int(m["launch__grid_size"]), int(m["launch__block_size"])
                key = "%s|%s|B%d" % (printed_name(m["Kernel Name"]), wl.group(1), 65536)
                traffic[key] = {"dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                                "gpu_time_us_under_ncu": float(m["gpu__time_duration.sum"]), "grid": grid, "block": block,
                                "source": "profiles/%s_ncu_summary.txt (ncu --set full, one launch of this instantiation)" % base}
            if on_box and "tiles_gain_test" not in f:
                os.remove(src)                       # keep one full report, the rest as summaries
        elif on_box:
            continue
        elif f.endswith((".json", ".csv", ".log", ".txt")) and not f.endswith("_run.log"):
            if f.endswith("_launches.csv"):
                keep = []
                for ln in open(src, errors="replace"):
                    if ln.startswith('"ID"') or "osc_step" in ln or "pack_tiles" in ln or "calc_error" in ln:
                        keep.append(ln)
                with open(os.path.join(PROF, f), "w") as fh:
                    fh.write("".join(keep))
            elif f.endswith(".log") and "ncu_" in f:
                continue
            elif f == "traffic.json":
                continue
            else:
                shutil.copyfile(src, os.path.join(PROF, f))
    if not on_box and os.path.isfile(os.path.join(EV, "traffic.json")):
        traffic = json.load(open(os.path.join(EV, "traffic.json")))
    if traffic:
        with open(traffic_path, "w") as fh:
            json.dump(traffic, fh, indent=1, sort_keys=True)
        print("traffic.json:", json.dumps(traffic, indent=1)[:1200])


if __name__ == "__main__":
    main()
