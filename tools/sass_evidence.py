#!/usr/bin/env python
"""SASS evidence of the shipped library (no GPU needed): per kernel the counts of the tcgen05 / TMA / TMEM mnemonics that
/opt/skills/guides/B200_PROFILING.md names.  usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "page_segmentation_b200", "libpcseg_b200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "HMMA", "FFMA", "DFMA", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in KEYS:
                per[cur][op] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("SASS evidence: cuobjdump -sass page_segmentation_b200/libpcseg_b200.so (sm_100a), instruction counts")
    print("total over all kernels: " + "  ".join(f"{k} {total[k]}" for k in KEYS if total[k]))
    print()
    print(f"{'kernel':<78}" + "".join(f"{k:>9}" for k in KEYS[:9]))
    rows = []
    for (name, c), dn in zip(per.items(), demangled):
        if not (c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["LDTM"]):
            continue
        short = re.sub(r"\(anonymous namespace\)::", "", dn)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(.*$", "", short)
        rows.append((short, c))
    for short, c in sorted(rows):
        print(f"{short[:77]:<78}" + "".join(f"{c[k]:>9}" for k in KEYS[:9]))
    print()
    print(f"{len(rows)} kernels issue tcgen05 / TMA instructions; {len(per)} kernels in the library.")


if __name__ == "__main__":
    sys.exit(main())
