#!/usr/bin/env python3
"""Static SASS opcode statistics per kernel of an object / shared library (no GPU needed):
    python tools/sass_stats.py build/irlosc_lane.o [name-filter]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cnt, ops, name = collections.Counter(), collections.defaultdict(collections.Counter), None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and name:
            cnt[name] += 1
            ops[name][m.group(2).split(".")[0]] += 1
    keys = ["DFMA", "DMUL", "DADD", "MUFU", "LDG", "LDS", "STS", "LDL", "STL", "LDC", "IMAD", "MOV", "BRA", "SHFL", "DSETP", "FSEL", "SEL"]
    for n in sorted(cnt):
        if flt and flt not in n:
            continue
        short = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
        print("%-60s total %6d  " % (short[-60:], cnt[n]) + " ".join("%s %d" % (k, ops[n][k]) for k in keys if ops[n][k]))


if __name__ == "__main__":
    main()
