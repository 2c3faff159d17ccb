#!/usr/bin/env python3
"""Hot source lines of one kernel from an .ncu-rep captured with --import-source on (needs -lineinfo):
    python tools/ncu_lines.py file.ncu-rep [top N]
Aggregates the `cuda,sass` source page per (file, line): warp-stall samples and warp instructions executed."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                                  stderr=subprocess.DEVNULL).decode()
    fname, hdr = None, None
    samples, inst, text = collections.Counter(), collections.Counter(), {}
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "File Path":
            fname = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            hdr = row
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or row[0] in ("Function Name",) or not row[0].strip().isdigit():
            continue
        key = (fname, int(row[0]))
        try:
            samples[key] += int(row[i_s])
            inst[key] += int(row[i_i])
        except (ValueError, IndexError):
            continue
        text[key] = row[1].strip()[:90]
    ts, ti = sum(samples.values()), sum(inst.values())
    print("total samples %d, warp instructions %d" % (ts, ti))
    per_file_s, per_file_i = collections.Counter(), collections.Counter()
    for (f, _l), v in samples.items():
        per_file_s[f] += v
    for (f, _l), v in inst.items():
        per_file_i[f] += v
    for f, v in per_file_s.most_common():
        print("  %-24s samples %5.1f %%   instructions %5.1f %%" % (f, 100.0 * v / max(ts, 1), 100.0 * per_file_i[f] / max(ti, 1)))
    print("-- top lines by stall samples")
    for key, v in samples.most_common(top):
        print("%5.1f %%  inst %5.1f %%  %s:%d  %s" % (100.0 * v / max(ts, 1), 100.0 * inst[key] / max(ti, 1), key[0], key[1], text.get(key, "")))


if __name__ == "__main__":
    main()
