"""Opcode histogram per kernel of libmxdet_sm100.so (cuobjdump -sass): which Blackwell / Hopper+ paths each kernel uses.
   python profiles/sass_summary.py > profiles/sass_summary.txt
UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier, UCGABAR = cluster
barrier, REDUX/CREDUX = warp reductions, RED/ATOMG = global atomics, ATOMS = shared atomics, ACQBULK/UTMAPF etc. = async
proxy helpers.  No UTC*MMA / LDTM: nothing on this path is a contraction (north_star)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mxdetection_b200", "libmxdet_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
KEY = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "UCGABAR", "REDUX", "CREDUX", "RED", "ATOMG", "ATOMS", "LDS", "STS", "LDG",
       "STG", "SHFL", "FFMA", "HFMA2", "BAR", "ACQBULK", "LDGSTS"]
print("libmxdet_sm100.so (sm_100a): SASS instruction counts per kernel (static)\n")
print("%-62s %6s  %s" % ("kernel", "instr", "  ".join(k for k in KEY)))
for k, c in hist.items():
    tot = sum(c.values())
    print("%-62s %6d  %s" % (k[:62], tot, "  ".join("%*d" % (len(n), c.get(n, 0)) for n in KEY)))
alls = collections.Counter()
for c in hist.values():
    alls.update(c)
print("\nwhole library: " + ", ".join("%s %d" % (k, alls[k]) for k in KEY if alls[k]))
print("tensor-core / TMEM opcodes (UTC*MMA, LDTM, STTM, HMMA, HGMMA): %d" % sum(v for k, v in alls.items() if k.startswith("UTC") or k in ("LDTM", "STTM", "HMMA", "HGMMA", "QGMMA", "IGMMA")))
