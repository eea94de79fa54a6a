"""Summarise an `ncu --set full --csv --page raw --log-file X.csv` capture per kernel as a markdown table (for profiles/).
Usage: python scripts/ncu_csv_summary.py gpurun_out/all_raw.csv "title" > profiles/rNN_x.md"""
import collections
import csv
import sys

M = [("gpu__time_duration.sum", "time us"), ("launch__grid_size", "CTAs"), ("launch__registers_per_thread", "regs"),
     ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %")]
US = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3}


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("sol::", "")
        a = agg.setdefault(name, {"n": 0})
        a["n"] += 1
        for m, _ in M:
            if m in col and r[col[m]] not in ("", "n/a"):
                a[m] = a.get(m, 0.0) + float(r[col[m]].replace(",", "")) * US.get(units[col[m]], 1.0)
    print("# %s\n" % title)
    print("Source: `ncu --set full --clock-control none` (cold caches, serialised launches).  Means over the captured launches of each")
    print("kernel; `GB/s` = (dram read + write) / duration.\n")
    print("| kernel | n | " + " | ".join(l for _, l in M) + " | DRAM GB/s |")
    print("|---|---|" + "---|" * (len(M) + 1))
    for name, a in agg.items():
        cells = []
        for m, lbl in M:
            if m not in a:
                cells.append("-"); continue
            v = a[m] / a["n"]
            cells.append("%.3f" % (v / 1e6) if "MB" in lbl else "%.4g" % v)
        t = a.get("gpu__time_duration.sum", 0) / a["n"]
        gbs = (a.get("dram__bytes_read.sum", 0) + a.get("dram__bytes_write.sum", 0)) / a["n"] / (t * 1e-6) / 1e9 if t else 0
        print("| `%s` | %d | " % (name, a["n"]) + " | ".join(cells) + " | %.0f |" % gbs)


if __name__ == "__main__":
    main()
