"""Summarise an `ncu --set full` report (.ncu-rep) per kernel as a markdown table for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rNN_x.md   (needs ncu on PATH)"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time us", 1.0),
    ("launch__grid_size", "CTAs", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("dram__bytes_read.sum", "dram rd", None),
    ("dram__bytes_write.sum", "dram wr", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)", 1.0),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %", 1.0),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier", 1.0),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1.0),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb", 1.0),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3}


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("sol::", "")
        a = agg.setdefault(name, {"n": 0})
        a["n"] += 1
        for m, _, _ in METRICS:
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", "")) * UNIT_SCALE.get(units[col[m]], 1.0)
                a[m] = a.get(m, 0.0) + v
    print("# %s\n" % title)
    print("Source: `ncu --set full --clock-control none --import-source on` (cold caches and serialised launches: compare shares,")
    print("not absolutes).  Values are means over the captured launches of each kernel.\n")
    print("| kernel | n | " + " | ".join(lbl for _, lbl, _ in METRICS) + " |")
    print("|---|---|" + "---|" * len(METRICS))
    for name, a in agg.items():
        cells = []
        for m, _, _ in METRICS:
            if m not in a:
                cells.append("-")
                continue
            v = a[m] / a["n"]
            cells.append("%.3g MB" % (v / 1e6) if "bytes" in m else "%.4g" % v)
        print("| `%s` | %d | " % (name, a["n"]) + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
