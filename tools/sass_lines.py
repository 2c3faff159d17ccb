#!/usr/bin/env python3
"""Attribute SASS instructions of one kernel to source lines (needs -lineinfo).

    cuobjdump -xelf all build/irlosc_fused.o && nvdisasm --print-line-info x.cubin > k.sass
    python tools/sass_lines.py k.sass <substring of the mangled kernel name> [opcode prefixes ...]

Prints the instruction total and the source lines with the most instructions whose opcode starts
with one of the prefixes (default: STL LDL = local-memory traffic, i.e. spills / dynamic arrays)."""
import collections
import re
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    ops = tuple(sys.argv[3:]) or ("STL", "LDL")
    txt = open(path).read()
    m = re.search(r"^\.text\.[^\n]*%s[^\n]*:\n" % re.escape(key), txt, re.M)
    if not m:
        sys.exit("kernel not found")
    end = txt.find("\n.text.", m.end())
    body = txt[m.end(): end if end > 0 else len(txt)]
    line, cnt, total, mix = None, collections.Counter(), 0, collections.Counter()
    for l in body.split("\n"):
        f = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if f:
            line = "%s:%s" % (f.group(1).split("/")[-1], f.group(2))
            continue
        i = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if i:
            total += 1
            op = i.group(1).split(".")[0]
            mix[op] += 1
            if op.startswith(ops):
                cnt[(line, op)] += 1
    print("instructions:", total, " matching:", sum(cnt.values()))
    print("mix:", ", ".join("%s %d" % kv for kv in mix.most_common(14)))
    for k, v in cnt.most_common(40):
        print("%4d  %-6s %s" % (v, k[1], k[0]))


if __name__ == "__main__":
    main()
