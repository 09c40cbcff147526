#!/usr/bin/env python3
"""Host-side probe (no GPU): throughput of the native run-file parser / writer (csrc/qf_io.cu: qf_runs_parse,
qf_runfile_write) next to the reference's way of doing the same work -- pd.read_csv(sep=r"\\s+") as in
src/fusion.py:67-72 and the per-line "%s Q0 %s %d %.6f %s" formatting loop of write_runfile (src/fusion.py:480-492).
    python tools/io_probe.py [n_queries=600] [depth=1000]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qpp-fusion-rag_b200"))
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

from qppfusion import native_io  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 600
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
R = 5
rng = np.random.default_rng(3)
tmp = tempfile.mkdtemp(prefix="qf_io_")
paths = []
for r in range(R):
    p = os.path.join(tmp, f"run{r}.res")
    with open(p, "w") as f:
        for q in range(nq):
            docs = rng.choice(2_681_468, size=depth, replace=False)
            sc = np.sort(rng.random(depth) * 30)[::-1]
            f.write("".join(f"test{q} Q0 doc{d} {i + 1} {s:.6f} run{r}\n" for i, (d, s) in enumerate(zip(docs, sc))))
    paths.append(p)
n_lines = nq * depth * R
n_bytes = sum(os.path.getsize(p) for p in paths)
cores = os.cpu_count()

def timed(fn, reps=3):
    """(first call, best of reps) in seconds"""
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t)
    return ts[0], min(ts), r


read_pd = lambda: [pd.read_csv(p, sep=r"\s+", names=["qid", "iter", "docno", "rank", "score", "runid"],
                               dtype={"qid": str, "docno": str}) for p in paths]
c_pd, t_pd, dfs = timed(read_pd, 2)
c_nat, t_nat, parsed = timed(lambda: native_io.parse_runs(paths))
assert sum(len(c["qid"]) for c in parsed.cols) == n_lines and parsed.all_micro
print(f"parse  {n_lines} lines / {n_bytes / 1e6:.0f} MB ({R} files): pandas read_csv(sep=r'\\s+') {t_pd:.2f} s = "
      f"{n_bytes / t_pd / 1e6:.0f} MB/s (1 thread)   qf_runs_parse {t_nat:.2f} s = {n_bytes / t_nat / 1e6:.0f} MB/s "
      f"({cores} threads; first call {c_nat:.2f} s)   x{t_pd / t_nat:.1f}")

# write: one fused run (union-sized lists) from CSR arrays
packed, qid_order = native_io.pack_parsed(parsed)
Q = len(qid_order)
lens = np.full(Q, 2 * depth, dtype=np.int64)
offsets = np.concatenate([[0], np.cumsum(lens)])
ids = rng.integers(0, len(parsed.doc_off) - 1, size=int(offsets[-1]), dtype=np.int32)
scores = np.sort(rng.random(int(offsets[-1])))[::-1].copy()
out_nat = os.path.join(tmp, "fused_native.res")
c_wn, t_wn, _ = timed(lambda: native_io.write_runfile_csr(out_nat, parsed, qid_order, offsets, ids, scores, "wcombsum-rsd"))
docs = parsed.doc_strings()
qstr = parsed.qid_strings()
out_py = os.path.join(tmp, "fused_python.res")
sample_q = max(1, Q // 10)
t = time.perf_counter()
with open(out_py, "w") as f:                       # the reference's loop (src/fusion.py:486-491), on a tenth of the queries
    for i in range(sample_q):
        qid = qstr[qid_order[i]]
        for rank, j in enumerate(range(offsets[i], offsets[i + 1]), start=1):
            f.write(f"{qid} Q0 {docs[ids[j]]} {rank} {scores[j]:.6f} wcombsum-rsd\n")
t_wp = (time.perf_counter() - t) * Q / sample_q
wb = os.path.getsize(out_nat)
print(f"write  {int(offsets[-1])} lines / {wb / 1e6:.0f} MB: python per-line loop {t_wp:.2f} s (extrapolated from {sample_q} queries) "
      f"= {wb / t_wp / 1e6:.0f} MB/s   qf_runfile_write {t_wn:.2f} s = {wb / t_wn / 1e6:.0f} MB/s ({cores} threads; first call "
      f"{c_wn:.2f} s)   x{t_wp / t_wn:.1f}")
for p in paths + [out_nat, out_py]:
    os.remove(p)
os.rmdir(tmp)
