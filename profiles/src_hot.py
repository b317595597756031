"""Per-instruction view of an ncu report (needs -lineinfo + --import-source on):
   python profiles/src_hot.py <report.ncu-rep> [top]
Prints instruction totals per SASS opcode and the address ranges that execute the most warp instructions."""
import csv
import subprocess
import sys
import collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernels may follow each other: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr = b["rows"][0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = b["rows"][1:]
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    samples = sum(int(r[ix["# Samples"]]) for r in data)
    print("== %s\n   warp instructions %d, samples %d" % (b["name"][:100], tot, samples))
    ops = collections.Counter()
    osmp = collections.Counter()
    for r in data:
        op = r[ix["Source"]].split()[0]
        if op.startswith("@"):
            op = r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] += int(r[ix["Instructions Executed"]])
        osmp[op] += int(r[ix["# Samples"]])
    print("   by opcode (Minstr, %samples): " + ", ".join("%s %.1f (%.0f%%)" % (k, v / 1e6, 100.0 * osmp[k] / max(samples, 1)) for k, v in ops.most_common(18)))
    # contiguous regions with similar execution counts
    regs = []
    for i, r in enumerate(data):
        n = int(r[ix["Instructions Executed"]])
        if regs and abs(n - regs[-1][2]) <= 0.02 * max(n, regs[-1][2], 1):
            regs[-1][1] = i; regs[-1][3] += n; regs[-1][4] += int(r[ix["# Samples"]])
        else:
            regs.append([i, i, n, n, int(r[ix["# Samples"]])])
    regs.sort(key=lambda x: -x[3])
    for a, z, n, s, sm in regs[:top]:
        print("   instr %5d..%5d (%4d)  exec/instr %9d  total %6.2f M (%4.1f%%)  samples %4.1f%%   %s" %
              (a, z, z - a + 1, n, s / 1e6, 100.0 * s / tot, 100.0 * sm / max(samples, 1), data[a][ix["Source"]].strip()[:50]))
