#!/usr/bin/env python3
"""GPU probe: the own tcgen05 GEMM (qf_linear_bf16) against torch.nn.functional.linear (cuBLASLt) on the four Linear
shapes of a BERT-base layer at the packed NQ token count."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qpp-fusion-rag_b200"))
import torch  # noqa: E402

from qppfusion import _native as N  # noqa: E402

dev = torch.device("cuda:0")
M = int(os.environ.get("PROBE_M", 48640))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, Nn, K, act in (("qkv", 2304, 768, 0), ("out", 768, 768, 0), ("ffn1+gelu", 3072, 768, 1), ("ffn2", 768, 3072, 0)):
    X = torch.randn((M, K), device=dev).to(torch.bfloat16)
    W = (torch.randn((Nn, K), device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn((Nn,), device=dev).to(torch.bfloat16)
    Y = torch.empty((M, Nn), dtype=torch.bfloat16, device=dev)

    def ours():
        N.check(N.lib().qf_linear_bf16(N.ptr(W), N.ptr(X), N.ptr(b), N.ptr(Y), M, Nn, K, act, N.stream_ptr(dev)))

    def lib():
        y = torch.nn.functional.linear(X, W, b)
        return torch.nn.functional.gelu(y) if act else y

    t_o, t_l = timed(ours), timed(lib)
    fl = 2.0 * M * Nn * K
    ref = lib().float()
    ours()
    torch.cuda.synchronize()
    err = float((Y.float() - ref).abs().max())
    print(f"{name:10s} M={M} N={Nn} K={K}: ours {t_o * 1e3:7.1f} us = {fl / t_o / 1e9:7.0f} TFLOP/s | torch/cuBLASLt"
          f"{' + gelu pass' if act else ''} {t_l * 1e3:7.1f} us = {fl / t_l / 1e9:7.0f} TFLOP/s | max |diff| {err:.4f}", flush=True)
