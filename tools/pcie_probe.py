#!/usr/bin/env python3
"""GPU-box probe: pinned H2D / D2H bandwidth alone and with both directions active, at the e2e step's sizes and with
the smaller D2H a micro-unit result format would give.  (Answers: is e2e limited by H2D alone, or by the duplex sum?)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
for up, down, label in ((138_204_272, 1 << 20, "H2D alone"), (1 << 20, 86_769_408, "D2H alone"),
                        (138_204_272, 86_769_408, "e2e step (ids+f64 scores down)"),
                        (138_204_272, 57_900_000, "ids + int32 micro scores down"),
                        (138_204_272, 28_950_000, "ids only down")):
    r = bench.pcie_probe(dev, up, down)
    print(f"{label:36s}: {r['ms_per_step_copies_only']:.3f} ms  H2D {r['h2d_gbs']:.1f} GB/s  D2H {r['d2h_gbs']:.1f} GB/s "
          f"-> {3452 / r['ms_per_step_copies_only'] / 1e3:.3f} M q/s ceiling", flush=True)
