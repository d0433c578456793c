#!/usr/bin/env python
"""For every kernel of libsfgpu.so whose name matches a pattern: the backward branches (loops) of its SASS with their
length, FADD / FMNMX3 counts and the local-memory (spill) instructions inside -- to check that spills stay out of
the unrolled macro-step loop.  Usage: python tools/sass_loops.py [regex] [lib]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "sf_dtw")
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "sigfish_b200", "libsfgpu.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    names = list(funcs)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for name, d in zip(names, dem):
        if not pat.search(d):
            continue
        ins = funcs[name]
        print(d)
        for a, t in ins:
            m = re.search(r"BRA(\.U)?\s+(!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
            if not m:
                continue
            tgt = int(m.group(3), 16)
            if tgt >= a:
                continue
            body = [x for y, x in ins if tgt <= y <= a]
            c = collections.Counter(x.split()[1].split(".")[0] if x.startswith("@") else x.split()[0].split(".")[0] for x in body)
            if c["FMNMX3"] == 0:
                continue
            print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, FADD {c['FADD']}, FMNMX3 {c['FMNMX3']}, SHFL {c['SHFL']}, "
                  f"LDS {c['LDS']}, STS {c['STS']}, LDL {c['LDL']}, STL {c['STL']}")


if __name__ == "__main__":
    main()
