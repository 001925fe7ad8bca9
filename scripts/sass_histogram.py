#!/usr/bin/env python
"""Opcode histogram per kernel of libssd3d_b200.so (cuobjdump -sass): which kernels carry the Blackwell-native
instructions (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, USETMAXREG =
setmaxnreg) and which still use the legacy tensor path (HMMA = mma.sync).
usage: sass_histogram.py [lib.so] > profiles/rNN_sass_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mslesions3d_b200", "libssd3d_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "USETMAXREG", "HMMA", "FFMA2",
       "SYNCS", "REDUX", "LDGSTS", "MUFU")
kernels = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1).split(".")[0]] += 1
        cur["_total"] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print("# %s -- %d kernels (cuobjdump -sass, sm_100a)" % (os.path.basename(lib), len(kernels)))
print("%-70s %7s  %s" % ("kernel", "instr", "  ".join(KEY)))
for (name, c), dn in zip(kernels.items(), demangled):
    short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "")).replace("ssd3d::", "")[:70]
    print("%-70s %7d  %s" % (short, c["_total"], "  ".join("%*d" % (len(k), c.get(k, 0)) for k in KEY)))
