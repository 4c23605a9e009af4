#!/usr/bin/env python
"""Turns ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py rep  <file.ncu-rep> <out.txt>    # selected raw metrics + the details page of every launch
  python tools/ncu_summary.py list <launches.csv>  <out.txt>   # per-kernel totals / shares of a gpu__time_duration launch list
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {path}", ""]
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"  {m:75s} {r[i]:>16s} {units[i]}")
        lines.append("")
    det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
    lines += ["# details page", det]
    open(out, "w").write("\n".join(lines))


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"# per-kernel totals of {path} (gpu__time_duration.sum; cold-cache, serialised: compare SHARES)",
             f"# total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches", ""]
    for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{t / 1e6:10.3f} ms {c:6d} launches {100 * t / tot:5.1f}%  {name[:110]}")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
