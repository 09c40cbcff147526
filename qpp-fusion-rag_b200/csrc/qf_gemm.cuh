// Own tensor-core GEMM for the encoder's Linear layers (qf_gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace qf {
// Y[M,N] = act(X[M,K] . W[N,K]^T + bias[N]); act: 0 = identity, 1 = erf-GELU.  Envelope: M % 128 == 0, N % 128 == 0,
// K % 64 == 0, 16-byte aligned operands.
bool gemm_supported(long long M, int N, int K);
int gemm_bias_act(const __nv_bfloat16* W, const __nv_bfloat16* X, const __nv_bfloat16* bias, __nv_bfloat16* Y, long long M,
                  int N, int K, int act, cudaStream_t st);
}  // namespace qf
