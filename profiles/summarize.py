"""Turns gpurun_out/*.ncu-rep + launches_*.csv into the small text summaries committed under profiles/.
Usage: python profiles/summarize.py <tag>        (reads gpurun_out/prof_<tag>.ncu-rep, launches_<tag>.csv)"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def main(tag):
    out = []
    rep = "gpurun_out/prof_%s.ncu-rep" % tag if __import__("os").path.exists("gpurun_out/prof_%s.ncu-rep" % tag) else "gpurun_out/prof_roialign_%s.ncu-rep" % tag
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) >= 3:
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            out.append("== ncu --set full: " + r[hdr.index("Kernel Name")][:110])
            for k in KEYS:
                if k in hdr:
                    out.append("  %-68s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
            out.append("  warp stall reasons (warps per issue-active cycle, > 0.3):")
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                    try:
                        if float(r[i]) > 0.3:
                            out.append("    %-40s %s" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), r[i]))
                    except ValueError:
                        pass
    try:
        lines = [l for l in open("gpurun_out/launches_%s.csv" % tag) if not l.startswith("==")]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in csv.DictReader(lines):
            agg[r["Kernel Name"][:90]][0] += 1
            agg[r["Kernel Name"][:90]][1] += float(r["Metric Value"].replace(",", ""))
        tot = sum(v[1] for v in agg.values())
        out.append("== launch list (ncu --metrics gpu__time_duration.sum, cold-cache serialised: compare SHARES)")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.append("  %5.1f%%  n=%3d  avg=%9.1f us  %s" % (100 * t / tot, n, t / 1e3 / n, k))
    except FileNotFoundError:
        pass
    open("profiles/ncu_%s.txt" % tag, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
