/*
 * Plain-C host of the C ABI (no Python, no torch): what a cgo / JNI / N-API binding does underneath.
 * Fuses two tiny queries with W-CombSUM -- the loop at reference src/fusion.py:266-278 -- and prints the ranked lists.
 *
 *   gcc -std=c99 -Iinclude examples/c_host_fuse.c -o c_host_fuse \
 *       -Lqpp-fusion-rag_b200/lib -lqppfusion -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/qpp-fusion-rag_b200/lib -Wl,-rpath,/usr/local/cuda/lib64
 *
 * The CUDA runtime is only used here to own the device buffers (cudaMalloc / cudaMemcpy); its four entry points are
 * declared by hand so that the file compiles with a C compiler and no CUDA headers.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qppfusion.h"

extern int cudaMalloc(void** p, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind); /* 1 = H2D, 2 = D2H */
extern int cudaFree(void* p);
extern int cudaDeviceSynchronize(void);

#define Q 2
#define R 2
#define K 4

static void* to_device(const void* host, size_t n) {
  void* d = NULL;
  if (cudaMalloc(&d, n) != 0 || cudaMemcpy(d, host, n, 1) != 0) {
    fprintf(stderr, "cudaMalloc/cudaMemcpy failed\n");
    exit(2);
  }
  return d;
}

int main(void) {
  /* doc ids, scores in micro-units (score * 1e6, the 6-decimal values of a .res file), list lengths, QPP weights */
  const int32_t ids[Q][R][K] = {{{7, 3, 9, 1}, {3, 5, 7, 0}}, {{2, 4, 0, 0}, {4, 2, 8, 6}}};
  const int32_t micro[Q][R][K] = {{{900000, 800000, 500000, 100000}, {700000, 600000, 200000, 0}},
                                  {{1000000, 250000, 0, 0}, {750000, 750000, 500000, 125000}}};
  const int32_t lens[Q][R] = {{4, 3}, {2, 4}};
  const double w[Q][R] = {{0.5, 0.25}, {1.0, 0.5}};
  int32_t out_ids[Q][R * K], out_lens[Q];
  double out_scores[Q][R * K];

  if (qf_device_check(0) != 0) {
    fprintf(stderr, "%s\n", qf_last_error());
    return 3;
  }
  void* d_ids = to_device(ids, sizeof ids);
  void* d_sc = to_device(micro, sizeof micro);
  void* d_lens = to_device(lens, sizeof lens);
  void* d_w = to_device(w, sizeof w);
  void *d_oi = NULL, *d_os = NULL, *d_ol = NULL;
  cudaMalloc(&d_oi, sizeof out_ids);
  cudaMalloc(&d_os, sizeof out_scores);
  cudaMalloc(&d_ol, sizeof out_lens);

  const int rc = qf_fuse((const int32_t*)d_ids, d_sc, QF_SCORE_MICRO, NULL, (const int32_t*)d_lens, (const double*)d_w, Q, R,
                         K, QF_METHOD_SUM, QF_NORM_NONE, 60.0, 0u, (int32_t*)d_oi, (double*)d_os, NULL, NULL,
                         (int32_t*)d_ol, NULL, NULL /* default stream */);
  if (rc != 0) {
    fprintf(stderr, "qf_fuse failed (%d): %s\n", rc, qf_last_error());
    return 1;
  }
  cudaDeviceSynchronize();
  cudaMemcpy(out_ids, d_oi, sizeof out_ids, 2);
  cudaMemcpy(out_scores, d_os, sizeof out_scores, 2);
  cudaMemcpy(out_lens, d_ol, sizeof out_lens, 2);
  for (int q = 0; q < Q; ++q) {
    printf("q%d:", q);
    for (int j = 0; j < out_lens[q]; ++j) printf(" %d:%.17g", out_ids[q][j], out_scores[q][j]);
    printf("\n");
  }
  cudaFree(d_ids); cudaFree(d_sc); cudaFree(d_lens); cudaFree(d_w); cudaFree(d_oi); cudaFree(d_os); cudaFree(d_ol);
  return 0;
}
