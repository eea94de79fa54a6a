#!/usr/bin/env python
"""`ncu --set full` report (.ncu-rep) -> one JSON per kernel under <outdir>/ncu_<kernel>.json, stamped with the content hash
of the kernel sources it was captured from (bench.py only reports `roofline.traffic` when the hash matches its own build).
Also writes the SASS opcode histogram of every kernel (cuobjdump) — the evidence for tcgen05 / TMA / TMEM use.

Usage (on the GPU box, after the capture): python scripts/ncu_to_json.py gpurun_out/x/prof.ncu-rep gpurun_out/x "how it was captured"
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (csrc_sha256)

METRICS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "launch__grid_size": "ctas",
    "launch__registers_per_thread": "registers",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
OPS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "LDS", "STS", "LDG",
       "STG", "RED", "ATOM", "SHFL", "REDUX", "BAR")


def base_name(full):
    full = full.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = full.split("(")[0].replace("void ", "").strip()
    n = n.split("<")[0]
    return n.split("::")[-1]


def sass_histograms():
    lib = os.path.join(ROOT, "solver_in_the_loop_b200", "libsol_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    hist = collections.defaultdict(collections.Counter)
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = base_name(d)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            op = m.group(1)
            hist[cur]["_total"] += 1
            for o in OPS:
                if op.startswith(o):
                    hist[cur][o] += 1
                    break
    return hist


def main():
    rep, outdir = sys.argv[1], sys.argv[2]
    how = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr):
            continue
        a = agg.setdefault(base_name(r[col["Kernel Name"]]), collections.defaultdict(float))
        a["n"] += 1
        for m, key in METRICS.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                a[key] += float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
    sha = bench.csrc_sha256()
    hist = sass_histograms()
    os.makedirs(outdir, exist_ok=True)
    summary = {}
    for name, a in agg.items():
        n = a.pop("n")
        d = {k: v / n for k, v in a.items()}
        rec = {"kernel": name, "csrc_sha256": sha, "how": how, "launches_captured": int(n),
               "dram_bytes_per_launch": d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0), "metrics": d,
               "sass_opcodes": dict(hist.get(name, {}))}
        summary[name] = rec
        json.dump(rec, open(os.path.join(outdir, "ncu_%s.json" % name), "w"), indent=1)
    # the two launches of the direct projection as one entry (bench.py's pressure-solve roofline object)
    if "k_direct_solve" in summary and "k_direct_apply" in summary:
        s, p = summary["k_direct_solve"], summary["k_direct_apply"]
        rec = {"kernel": "k_direct_solve+k_direct_apply", "csrc_sha256": sha, "how": how,
               "dram_bytes_per_launch": s["dram_bytes_per_launch"] + p["dram_bytes_per_launch"],
               "metrics": {"duration_us": s["metrics"].get("duration_us", 0) + p["metrics"].get("duration_us", 0)}}
        json.dump(rec, open(os.path.join(outdir, "ncu_k_direct_solve+k_direct_apply.json"), "w"), indent=1)
    print("| kernel | n | us | DRAM rd+wr MB | tensor pipe % | occupancy % | issue % | tcgen05/TMA SASS |")
    print("|---|---|---|---|---|---|---|---|")
    for name, rec in summary.items():
        m = rec["metrics"]; h = rec["sass_opcodes"]
        tc = " ".join("%s:%d" % (k, h[k]) for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM") if h.get(k))
        print("| `%s` | %d | %.2f | %.3f | %.1f | %.1f | %.1f | %s |" % (name, rec["launches_captured"], m.get("duration_us", 0), rec["dram_bytes_per_launch"] / 1e6,
                                                                      m.get("tensor_pipe_active_pct", 0), m.get("occupancy_pct", 0), m.get("issue_pct", 0), tc))


if __name__ == "__main__":
    main()
