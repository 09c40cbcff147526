#!/usr/bin/env python3
"""Headline metrics + per-source-line-range instruction shares of one kernel in an .ncu-rep.
usage: tools/ncu_phase.py file.ncu-rep source.cu queries [lo-hi:name ...]"""
import csv
import subprocess
import sys

rep, src_name, nq = sys.argv[1], sys.argv[2], float(sys.argv[3])
ranges = []
for a in sys.argv[4:]:
    rng, name = a.split(":")
    lo, hi = rng.split("-")
    ranges.append((name, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__warps_eligible.avg.per_cycle_active", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg, cur, h2 = [], "", None
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h2 = r
        ii, isamp = h2.index("Instructions Executed"), h2.index("# Samples")
        continue
    if h2 and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), int(r[ii]), int(r[isamp]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(a[2] for a in agg) or 1
ts = sum(a[3] for a in agg) or 1
print(f"total warp-instructions {tot} = {tot / nq:.0f} per query; samples {ts}")
for name, lo, hi in ranges:
    i = sum(a[2] for a in agg if a[0] == src_name and lo <= a[1] <= hi)
    s_ = sum(a[3] for a in agg if a[0] == src_name and lo <= a[1] <= hi)
    print(f"  {name:28s} {lo:4d}-{hi:<4d} {100 * i / tot:5.1f}% inst ({i / nq:8.0f}/query) {100 * s_ / ts:5.1f}% samples")
other = {}
for a in agg:
    if a[0] != src_name:
        o = other.setdefault(a[0], [0, 0])
        o[0] += a[2]
        o[1] += a[3]
for k, v in other.items():
    print(f"  {k:28s}           {100 * v[0] / tot:5.1f}% inst ({v[0] / nq:8.0f}/query) {100 * v[1] / ts:5.1f}% samples")
print("--- top lines by samples")
for a in sorted(agg, key=lambda x: -x[3])[:14]:
    print(f"  {100 * a[3] / ts:5.1f}% smp {100 * a[2] / tot:5.1f}% inst  {a[0]}:{a[1]:<4d} {a[4]}")
