#!/usr/bin/env python3
"""Executed SASS opcodes of one kernel attributed to source lines: joins the per-instruction counters of an ncu report
(--page source --print-source sass) with nvdisasm's line info of the same cubin, by instruction offset.
usage: tools/sass_by_line.py report.ncu-rep object.o 'mangled-name-substring' source.cu queries"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, pat, src_path, nq = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." in l and pat in l][0]
end = next((i for i in range(start + 1, len(dis)) if dis[i].strip().startswith(".section")), len(dis))
cur, off2 = None, {}
for ln in dis[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off2[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][0], 16)
per, ops, smp, tot, ts = collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    try:
        off, n, s = int(r[0], 16) - base, int(r[ii]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    if off not in off2:
        continue
    fl, txt = off2[off]
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt)
    op = "IMAD.MOV" if m and m.group(2).startswith("IMAD.MOV") else (m.group(2).split(".")[0] if m else "?")
    per[fl][op] += n
    per[fl]["_all"] += n
    per[fl]["_smp"] += s
    ops[op] += n
    smp[op] += s
    tot += n
    ts += s
print(f"{tot / nq:.0f} warp-instructions per query (under the profiler), {ts} stall samples")
print("opcode mix: " + ", ".join(f"{o} {100 * n / tot:.1f}% ({100 * smp[o] / ts:.1f}% smp)" for o, n in ops.most_common(14)))
ctrl = ("ISETP", "BRA", "BSSY", "BSYNC", "SEL", "IMAD.MOV", "PLOP3")
src = open(src_path).read().split("\n")
print("source lines by control / select / move instructions (ISETP BRA BSSY BSYNC SEL IMAD.MOV PLOP3), per query:")
for fl, c in sorted(per.items(), key=lambda kv: -sum(kv[1][o] for o in ctrl))[:16]:
    cs = sum(c[o] for o in ctrl)
    text = src[fl[1] - 1].strip()[:64] if fl and fl[0] == os.path.basename(src_path) else str(fl)
    print(f"  line {fl[1] if fl else 0:4d}: {cs / nq:6.0f} of {c['_all'] / nq:6.0f} instr ({100 * c['_all'] / tot:4.1f}% of all, {100 * c['_smp'] / ts:4.1f}% smp) | {text}")
