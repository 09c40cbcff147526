/*
 * TEST INFRASTRUCTURE ONLY -- C restatement (oracle) of the reference's weighted fusion and QPP bridge
 * arithmetic, for full-size parity checks and as the multi-threaded CPU baseline of bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Restates (citations relative to /root/reference):
 *   weighted_combsum / weighted_combmnz / weighted_rrf   src/fusion.py:237-380
 *   write_runfile's ranking (stable, score descending)   src/fusion.py:488
 *   QPPBridge.computeQPPScores, 13 estimators            src/qpp/java/qpp/QPPBridge.java:134-453
 * It is validated against oracle/fusion_oracle.py and oracle/qpp_oracle.py (tests/test_oracle_c.py), which in
 * turn are pinned on the real reference's golden vectors -- see those files for the parity status.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -o oracle/_build/libqf_oracle.so oracle/qf_oracle.c -lm
 * (-ffp-contract=off: the reference's float64 arithmetic has one rounding per operator, no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QF_SUM 0
#define QF_MNZ 1
#define QF_RRF 2

/* ------------------------------------------------------------------------------------------ fusion */
typedef struct {
  int32_t key;
  int32_t idx; /* index into the union arrays, -1 = empty */
} slot_t;

typedef struct {
  double score;
  int32_t pos;
} rank_t;

static int cmp_rank(const void* a, const void* b) {
  const rank_t* x = (const rank_t*)a;
  const rank_t* y = (const rank_t*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->pos > y->pos) - (x->pos < y->pos); /* stable: ties keep first-seen order */
}

/* One query.  lists are given in visiting order (the caller sorts W-RRF rows by rank beforehand, as
 * pandas sort_values("rank") does).  micro != NULL: scores are int32 micro-units, value = i / 1e6.
 * Outputs in ranked order; out_pos = first-seen index.  Returns the union size. */
static int fuse_one(const int32_t* ids, const int32_t* micro, const double* f64, const int32_t* ranks, const int32_t* lens,
                    const double* w, int R, int K, int method, double rrf_k, int32_t* out_ids, double* out_scores,
                    int32_t* out_counts, int32_t* out_pos, slot_t* table, int H, double* acc, int32_t* cnt, int32_t* doc,
                    rank_t* rk) {
  int n_u = 0;
  for (int i = 0; i < H; ++i) table[i].idx = -1;
  for (int r = 0; r < R; ++r) {
    const int len = lens[r];
    for (int i = 0; i < len; ++i) {
      const int32_t d = ids[(size_t)r * K + i];
      uint32_t h = ((uint32_t)d * 2654435761u) & (uint32_t)(H - 1);
      while (table[h].idx >= 0 && table[h].key != d) h = (h + 1) & (uint32_t)(H - 1);
      int u = table[h].idx;
      if (u < 0) { /* defaultdict(float): first-seen insertion, start value 0.0 */
        u = n_u++;
        table[h].key = d;
        table[h].idx = u;
        acc[u] = 0.0;
        cnt[u] = 0;
        doc[u] = d;
      }
      double v;
      if (method == QF_RRF) {
        const double rank = ranks ? (double)ranks[(size_t)r * K + i] : (double)(i + 1);
        v = w[r] / (rrf_k + rank);
      } else {
        const double s = micro ? (double)micro[(size_t)r * K + i] / 1.0e6 : f64[(size_t)r * K + i];
        v = w[r] * s;
      }
      acc[u] = acc[u] + v;
      cnt[u] += 1;
    }
  }
  for (int u = 0; u < n_u; ++u) {
    double s = acc[u];
    if (method == QF_MNZ) s = s * (double)cnt[u];
    rk[u].score = s + 0.0;
    rk[u].pos = u;
  }
  qsort(rk, (size_t)n_u, sizeof(rank_t), cmp_rank);
  for (int j = 0; j < n_u; ++j) {
    const int u = rk[j].pos;
    out_ids[j] = doc[u];
    out_scores[j] = rk[j].score;
    if (out_counts) out_counts[j] = cnt[u];
    if (out_pos) out_pos[j] = u;
  }
  return n_u;
}

/* Q queries, layout identical to qf_fuse ([Q][R][K] inputs, [Q][R*K] outputs).  Returns 0. */
int qfo_fuse_weighted(const int32_t* ids, const int32_t* micro, const double* f64, const int32_t* ranks, const int32_t* lens,
                      const double* weights, int Q, int R, int K, int method, double rrf_k, int32_t* out_ids,
                      double* out_scores, int32_t* out_counts, int32_t* out_pos, int32_t* out_lens, int n_threads) {
  const int n = R * K;
  int H = 16;
  while (H < 2 * n) H <<= 1;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
  {
    slot_t* table = (slot_t*)malloc(sizeof(slot_t) * (size_t)H);
    double* acc = (double*)malloc(sizeof(double) * (size_t)n);
    int32_t* cnt = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* doc = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    rank_t* rk = (rank_t*)malloc(sizeof(rank_t) * (size_t)n);
#pragma omp for schedule(dynamic, 8)
    for (int q = 0; q < Q; ++q) {
      const size_t o = (size_t)q * R * K;
      out_lens[q] = fuse_one(ids + o, micro ? micro + o : NULL, f64 ? f64 + o : NULL, ranks ? ranks + o : NULL,
                             lens + (size_t)q * R, weights + (size_t)q * R, R, K, method, rrf_k, out_ids + (size_t)q * n,
                             out_scores + (size_t)q * n, out_counts ? out_counts + (size_t)q * n : NULL,
                             out_pos ? out_pos + (size_t)q * n : NULL, table, H, acc, cnt, doc, rk);
    }
    free(table); free(acc); free(cnt); free(doc); free(rk);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ QPP bridge */
static double java_average(const double* x, int n, double or_else) {
  /* DoubleStream.average(): Kahan-compensated sum (Collectors.sumWithCompensation) / count */
  if (n == 0) return or_else;
  double s = 0.0, c = 0.0, simple = 0.0;
  for (int i = 0; i < n; ++i) {
    const double tmp = x[i] - c;
    const double velvel = s + tmp;
    c = (velvel - s) - tmp;
    s = velvel;
    simple += x[i];
  }
  double tot = s - c;
  if (isnan(tot) && isinf(simple)) tot = simple;
  return tot / (double)n;
}

static double pop_var(const double* x, int n) {
  const double mean = java_average(x, n, 0.0);
  double tmp[256];
  for (int i = 0; i < n; ++i) tmp[i] = (x[i] - mean) * (x[i] - mean);
  return java_average(tmp, n, 0.0);
}

/* scores: the (already float32-rounded) list prefix as doubles; feats = {n_terms, n_unique, max_len, sum_len};
 * out[13] in config.qpp.methods order. k_cut <= 256. */
void qfo_bridge_scores(const double* scores, int n, const int32_t* feats, int k_cut, double* out) {
  for (int c = 0; c < 13; ++c) out[c] = 0.0;
  if (n <= 0) return;
  const int k = n < k_cut ? n : k_cut;
  const double* x = scores;
  const double mean = java_average(x, k, 0.0);
  const int nT = feats[0] > 1 ? feats[0] : 1;
  {
    double a = 0.0;
    for (int i = 0; i < k; ++i) { const double d = x[i] - mean; a += d * d; }
    out[0] = (a / k) * 1.0;                                               /* nqc */
  }
  {
    double mu = mean <= 0 ? 1.0 : mean, a = 0.0;
    for (int i = 0; i < k; ++i) if (x[i] > 0) a += x[i] * fabs(log(x[i] / mu));
    out[1] = a / k;                                                       /* smv */
  }
  {
    const double base = 1.0 / fmax(0.01, mean);
    double a = 0.0;
    for (int i = 0; i < k; ++i) a += (x[i] - base);
    out[2] = a / ((double)nT * k);                                        /* wig */
  }
  if (k >= 2) {
    double best = 0.0, cum = 0.0;
    for (int i = 2; i <= k; ++i) {
      const double v = pop_var(x, i);
      const double sd = sqrt(v);
      if (sd > best) best = sd;
      double a = 0.0;                     /* computeNQC(subset): plain-sum variance about the Kahan mean */
      const double m = java_average(x, i, 0.0);
      for (int j = 0; j < i; ++j) { const double d = x[j] - m; a += d * d; }
      cum += a / i;
    }
    out[3] = best / sqrt((double)(nT > 1 ? nT : 1));                      /* SigmaMax */
    out[9] = cum / (k - 1);                                               /* cumnqc */
  }
  if (k >= 2) {
    const double thr = x[0] * 0.5;
    double f[256];
    int c = 0;
    for (int i = 0; i < k; ++i) if (x[i] >= thr) f[c++] = x[i];
    out[4] = c >= 2 ? sqrt(pop_var(f, c)) : 0.0;                          /* SigmaX */
  }
  if (k >= 3) {
    const double sd = sqrt(pop_var(x, k));
    if (!(sd < 1e-10)) {
      double a = 0.0;
      for (int i = 0; i < k; ++i) a += pow((x[i] - mean) / sd, 3);
      out[5] = a / k;                                                     /* RSD (skewness) */
    }
  }
  {
    double u = 0.0, ws = 0.0;
    const int kk = k < 20 ? k : 20;
    for (int i = 0; i < kk; ++i) {
      const double w = 1.0 / (log((double)(i + 2)) / log(2.0));
      u += x[i] * w;
      ws += w;
    }
    out[6] = u / ws;                                                      /* UEF */
  }
  out[7] = log(1.0 + feats[1]) + log(1.0 + feats[2]) * 0.5;               /* MaxIDF proxy */
  {
    const double avg_len = feats[0] > 0 ? (double)feats[3] / feats[0] : 3.0;
    const double div = feats[0] > 0 ? (double)feats[1] / feats[0] : NAN;
    out[8] = log(1.0 + avg_len) * div;                                    /* avgidf proxy */
  }
  if (mean > 0) {
    double a = 0.0;
    for (int i = 0; i < k; ++i) {
      if (x[i] <= 0) continue;
      const double f2 = ((x[i] - mean) * (x[i] - mean)) / x[i];
      a += pow(pow(1.0, 0.33) * pow(f2, 0.33), 0.33);
    }
    out[10] = (a / k) * 1.0;                                              /* snqc */
  }
  if (k >= 2) {
    double mx = x[0], mn = x[0];
    for (int i = 1; i < k; ++i) { if (x[i] > mx) mx = x[i]; if (x[i] < mn) mn = x[i]; }
    const double range = mx - mn;
    out[11] = out[12] = range > 0 ? log(1.0 + 1.0 / range) : 0.0;         /* dense-qpp(-m) */
  }
}

/* Batched: micro-unit lists [Q][R][K] -> out [Q][R][13]; inputs rounded decimal -> float32 -> double. */
int qfo_qpp_bridge_batch(const int32_t* micro, const int32_t* lens, const int32_t* feats, int Q, int R, int K, int top_k,
                         int k_cut, double* out, int n_threads) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 0 ? n_threads : 1)
  for (int li = 0; li < Q * R; ++li) {
    double x[256];
    int n = lens[li];
    if (n > top_k) n = top_k;
    const int k = n < k_cut ? n : k_cut;
    for (int i = 0; i < k; ++i) x[i] = (double)(float)((double)micro[(size_t)li * K + i] / 1.0e6);
    qfo_bridge_scores(x, n < k ? n : k, feats + (size_t)(li / R) * 4, k_cut, out + (size_t)li * 13);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ arithmetic proofs */
/* The CUDA kernels rebuild the double of a 6-decimal score from its integer micro-units without a division:
 *     q0 = a * 1e-6;  r = fma(-1e6, q0, a);  q = fma(r, 1e-6, q0)
 * (Markstein's correction step).  This exhaustive check counts the integers a in [lo, hi) for which q differs
 * from the correctly rounded a / 1e6 (the value strtod("x.xxxxxx") returns).  Expected result: 0. */
long long qfo_check_micro_division(long long lo, long long hi, int n_threads) {
  long long bad = 0;
#pragma omp parallel for reduction(+ : bad) num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
  for (long long a = lo; a < hi; ++a) {
    const double x = (double)a;
    const double q0 = x * 1.0e-6;
    const double r = fma(-1.0e6, q0, x);
    const double q = fma(r, 1.0e-6, q0);
    if (q != x / 1.0e6) ++bad;
  }
  return bad;
}
