#!/usr/bin/env python
"""Summaries of the ncu evidence brought back in gpurun_out/ (see scripts/gpu_ncu_s5.sh), written under profiles/.

  python scripts/summarize_ncu.py launches <ncu --csv launch list> <out.csv> "<header comment>"
      per kernel: launches, total us, share of the listed launches, DRAM MB read / written, DRAM GB/s
  python scripts/summarize_ncu.py full <out.csv> "<header comment>" <a.ncu-rep> [<b.ncu-rep> ...]
      one row per captured launch with the metrics DESIGN.md quotes (needs `ncu` on PATH to read the reports)
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__cycles_active.avg",
]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ub::", "")


def launches(path, out, header):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = collections.OrderedDict()
    for r in rd:
        k = short(r["Kernel Name"])
        d = per.setdefault(k, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
        d["ids"].add(r["ID"])
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ns"] += v * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        elif r["Metric Name"] == "dram__bytes_read.sum":
            d["rd"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
        elif r["Metric Name"] == "dram__bytes_write.sum":
            d["wr"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
    tot = sum(d["ns"] for d in per.values())
    with open(out, "w") as f:
        f.write(f"# {header}\n")
        f.write("kernel,launches,total_us,share,dram_read_MB,dram_write_MB,dram_GBps\n")
        for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
            gbps = (d["rd"] + d["wr"]) / d["ns"] if d["ns"] else 0.0
            f.write(f"\"{k}\",{len(d['ids'])},{d['ns'] / 1e3:.1f},{d['ns'] / tot:.4f},{d['rd'] / 1e6:.1f},"
                    f"{d['wr'] / 1e6:.1f},{gbps:.0f}\n")
        f.write(f"\"TOTAL\",{sum(len(d['ids']) for d in per.values())},{tot / 1e3:.1f},1.0000,"
                f"{sum(d['rd'] for d in per.values()) / 1e6:.1f},{sum(d['wr'] for d in per.values()) / 1e6:.1f},\n")


def full(out, header, reps):
    with open(out, "w") as f:
        f.write(f"# {header}\n")
        wrote_head = False
        for rep in reps:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            lines = [l for l in txt.splitlines() if l.startswith('"')]
            rd = csv.reader(io.StringIO("\n".join(lines)))
            head = next(rd)
            units = next(rd)
            idx = {m: head.index(m) for m in FULL_METRICS if m in head}
            if not wrote_head:
                cols = ["Kernel Name", "Grid Size", "Block Size"] + [f"{m} [{units[idx[m]]}]" for m in idx]
                f.write(",".join(cols) + "\n")
                wrote_head = True
            for r in rd:
                vals = [f"\"{short(r[head.index('Kernel Name')])}\"", f"\"{r[head.index('Grid Size')]}\"",
                        f"\"{r[head.index('Block Size')]}\""] + [r[idx[m]].replace(",", "") for m in idx]
                f.write(",".join(vals) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4:])
