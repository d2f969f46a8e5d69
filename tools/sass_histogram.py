#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass), for profiles/: shows which
kernels carry tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP) and
mbarrier (SYNCS) instructions.   python tools/sass_histogram.py > profiles/sass_opcodes_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bpreveal_b200", "libbpreveal_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP",
       "UTMAPF", "SYNCS", "ACQBULK", "HMMA", "FFMA", "LDS", "STS", "LDG", "STG", "BAR", "REDUX"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = list(kernels)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    for n, d in zip(names, out.splitlines()):
        demangle[n] = d
    total = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels")
    print("# kernel | instructions | " + " ".join(KEY))
    for n, c in kernels.items():
        total.update(c)
        short = re.sub(r"\(bp::.*", "", demangle[n]).replace("bp::", "")
        print(f"{short} | {sum(c.values())} | " + " ".join(f"{k}={c[k]}" for k in KEY if c[k]))
    print("# TOTAL | " + " ".join(f"{k}={total[k]}" for k in KEY if total[k]))


if __name__ == "__main__":
    sys.exit(main())
