#!/usr/bin/env python
"""Summaries of ncu artefacts for profiles/ (runs on the GPU-less build box).

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/launches_r1_summary.csv
  python tools/ncu_summary.py full gpurun_out/conv_full_r1.ncu-rep profiles/conv_full_r1_summary.csv
"""
import collections
import csv
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ik].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], f"{a[1] / 1e3:.1f}", f"{100 * a[1] / tot:.1f}", f"{a[1] / a[0] / 1e3:.2f}"])
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled")
             and h.endswith("per_issue_active.ratio")]
    with open(dst, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow([hdr[i] for i in idx] + ["top_stalls_per_issue"])
        w.writerow([rows[1][i] for i in idx] + [""])
        for r in rows[2:]:
            st = sorted(((float(r[i] or 0), hdr[i].replace("smsp__average_warps_issue_stalled_", "")
                          .replace("_per_issue_active.ratio", "")) for i in stall), reverse=True)[:4]
            w.writerow([r[i] for i in idx] + [" ".join(f"{n}={v:.2f}" for v, n in st)])
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
