#!/usr/bin/env python3
"""CPU-only probe of the host side of run_fusion (no GPU needed): parse -> pack -> qpp load -> weights matrix.
usage: python tools/host_path_probe.py [n_queries]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from qppfusion import fusion, native_io as nio  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = tempfile.mkdtemp(prefix="qf_hostprobe_")
t = time.perf_counter()
runs_dir, qpp_dir = bench.write_synth_run_files(d, nq, seed=3)
print(f"generated {nq} queries in {time.perf_counter() - t:.2f}s -> {runs_dir}")
files, suffix = fusion._run_files(runs_dir, True)
names = [f.replace(suffix, "") for f in files]
sz = sum(os.path.getsize(os.path.join(runs_dir, f)) for f in files)
t = time.perf_counter(); parsed = nio.parse_runs([os.path.join(runs_dir, f) for f in files], names); t_parse = time.perf_counter() - t
t = time.perf_counter(); packed, qorder = nio.pack_parsed(parsed, "weighted"); t_pack = time.perf_counter() - t
t = time.perf_counter(); qpp = fusion.load_qpp_scores(qpp_dir); t_qpp = time.perf_counter() - t
t = time.perf_counter(); w = fusion._weights_matrix(packed, qpp, 5); t_w = time.perf_counter() - t
t = time.perf_counter(); ds = parsed.doc_strings(); t_ds = time.perf_counter() - t
print(f"{sz / 1e6:.0f} MB text: parse {t_parse * 1e3:.0f} ms ({sz / t_parse / 1e6:.0f} MB/s), pack {t_pack * 1e3:.0f} ms, "
      f"load_qpp {t_qpp * 1e3:.0f} ms, weights {t_w * 1e3:.0f} ms, doc_strings {t_ds * 1e3:.0f} ms; "
      f"total {(t_parse + t_pack + t_qpp + t_w) * 1e3:.0f} ms = {nq / (t_parse + t_pack + t_qpp + t_w):.0f} q/s host-side ceiling")
