"""GPU parity of the own tensor-core GEMM (csrc/qf_gemm.cu, the encoder's Linear layers) against a torch fp32 reference
of the same op: Y = act(X . W^T + bias), bf16 operands, fp32 accumulation, bf16 output.  Tolerance: one bf16 rounding of
the result (relative 2^-8) plus fp32 summation-order noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def linear(W, X, b, act):
    from qppfusion import _native as N
    M, K = X.shape
    Nn = W.shape[0]
    Y = torch.empty((M, Nn), dtype=torch.bfloat16, device=X.device)
    with torch.cuda.device(X.device):
        N.check(N.lib().qf_linear_bf16(N.ptr(W), N.ptr(X), N.ptr(b), N.ptr(Y), M, Nn, K, act, N.stream_ptr(X.device)))
    return Y


@pytest.mark.parametrize("M,N,K,act", [(128, 128, 64, 0), (128, 256, 64, 0), (384, 768, 768, 0), (1024, 2304, 768, 0),
                                        (512, 768, 3072, 0), (256, 3072, 768, 1), (640, 384, 128, 1),
                                        (128 * 150, 256, 192, 0)])
def test_linear_matches_fp32_reference(M, N, K, act):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    X = (torch.randn((M, K), generator=g, device=dev) * 0.7).to(torch.bfloat16)
    W = (torch.randn((N, K), generator=g, device=dev) * (1.0 / K ** 0.5)).to(torch.bfloat16)
    b = (torch.randn((N,), generator=g, device=dev) * 0.3).to(torch.bfloat16)
    Y = linear(W, X, b, act)
    torch.cuda.synchronize()
    ref = X.float() @ W.float().T + b.float()
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = (Y.float() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2e-3
    assert bool((err <= tol).all()), f"max err {float(err.max())} at {int(err.argmax())}"
    # every tile was written (no stale rows): compare the row / column sums too
    assert torch.allclose(Y.float().sum(dim=0), ref.to(torch.bfloat16).float().sum(dim=0), rtol=2e-2, atol=0.5)


def test_repeated_calls_and_alignment_errors():
    from qppfusion import _native as N
    dev = torch.device("cuda:0")
    X = torch.randn((256, 128), device=dev).to(torch.bfloat16)
    W = torch.randn((256, 128), device=dev).to(torch.bfloat16)
    b = torch.zeros((256,), device=dev, dtype=torch.bfloat16)
    a = linear(W, X, b, 0)
    for _ in range(5):
        assert torch.equal(linear(W, X, b, 0), a)              # deterministic
    with pytest.raises(N.NativeError):
        linear(W[:100], X, b[:100], 0)                         # N % 128 != 0
