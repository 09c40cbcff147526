#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line.
usage: tools/ncu_lines.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
cur_file = ""
agg = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit():
        try:
            agg.append((int(r[i_inst]), int(r[i_samp]), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot_i = sum(a[0] for a in agg) or 1
tot_s = sum(a[1] for a in agg) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("--- by instructions executed")
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100*a[0]/tot_i:5.1f}% inst {100*a[1]/tot_s:5.1f}% smp  {a[2]}:{a[3]:<4d} {a[4]}")
print("--- by stall samples")
for a in sorted(agg, key=lambda x: -x[1])[:top]:
    print(f"{100*a[1]/tot_s:5.1f}% smp {100*a[0]/tot_i:5.1f}% inst  {a[2]}:{a[3]:<4d} {a[4]}")
