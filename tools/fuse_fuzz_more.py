#!/usr/bin/env python3
"""GPU: 40 more seeds of tests/test_fuse_fuzz_gpu.py (320 random cases through all six fusion kernel paths, qf_fuse_deep
included) -- a one-off stress beyond the 6 seeds the test suite runs.  python tools/fuse_fuzz_more.py"""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
import test_fuse_fuzz_gpu as t
n = 0
for seed in range(6, 46):
    t.test_kernel_paths_agree_and_match_the_oracle(seed)
    n += 1
print("extra fuzz seeds ok:", n)
