#!/usr/bin/env python3
"""Summarise ncu outputs into profiles/ (text, committed).
  tools/profile_summary.py launches <launches.csv> <out.txt> "<command line>"
  tools/profile_summary.py kernel <file.ncu-rep> <out.txt> "<title>" [src.csv]"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(path, out, cmd):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].replace("<unnamed>::", "")[:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    is_torch = lambda k: "native::" in k or "kernelPointwiseApply" in k or k.startswith("void at::") or "elementwise_kernel" in k or "cub::" in k
    ours = {k: v for k, v in agg.items() if not is_torch(k)}
    tot = sum(v[1] for v in ours.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  {cmd}\n")
        f.write("# libqppfusion kernels; per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / v[0] / 1e3:10.1f} us/launch  share {100 * v[1] / tot:5.1f}%  launches={v[0]:3d}  {k}\n")
        t_n = sum(v[0] for k, v in agg.items() if is_torch(k))
        t_t = sum(v[1] for k, v in agg.items() if is_torch(k))
        f.write(f"# PyTorch kernels in the same capture (synthetic data generation, dtype casts, parity spot check; outside "
                f"the timed regions): {t_n} launches, {t_t / 1e6:.2f} ms in total\n")


def kernel(rep, out, title, src=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on  ({rep.split('/')[-1]})\n")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                f.write(f"{h:95s} {v} {u}\n")
        if src:
            r = subprocess.run([sys.executable, __file__.replace("profile_summary.py", "ncu_lines.py"), src, "25"],
                               capture_output=True, text=True).stdout
            f.write("\n# per source line (share of warp-instructions executed / of stall samples)\n" + r)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        kernel(*sys.argv[2:])
