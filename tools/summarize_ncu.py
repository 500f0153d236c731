#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.txt>     per-kernel launch counts / total time / share
  python tools/summarize_ncu.py full <report.ncu-rep> <out.txt>       key `--set full` metrics of every captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|unnamed>::|q3::|void ", "", name)
    return re.sub(r"\(.*", "", name).strip("<> ")


def launches(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {sum(v[0] for v in agg.values())} launches, "
                f"{tot / 1e6:.3f} ms summed (cold-cache, serialised: shares matter, not absolutes)\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:9.2f} {100 * v[1] / tot:6.1f}%\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {len(rows) - 2} captured launches of {path.split('/')[-1]}\n")
        for n, r in enumerate(rows[2:]):
            f.write(f"[{n}] {short(r[ki])}\n")
            for k, i in cols:
                f.write(f"    {k:82s} {r[i]} {units[i]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
