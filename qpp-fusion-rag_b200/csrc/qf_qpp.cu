// qf_qpp_stats / qf_qpp_colnorm / qf_qpp_select_weight
//
// Batched score-distribution QPP predictors: one warp per (query, retriever) list, float64 arithmetic,
// coalesced loads of the list prefix, warp-shuffle reductions.  Restates
//   qpp.QPPBridge.computeQPPScores            reference src/qpp/java/qpp/QPPBridge.java:134-453
//   the research predictors behind the names  NQC/SMV/WIG/SigmaX/NQCCalibrated/RSD Specificity.java
//   OverlapStats.computeRBO                   src/qpp/java/correlation/OverlapStats.java:77-99
//   _normalize_qpp + "%.6f" round trip        src/qpp.py:329-374
//   get_qpp_weight                            src/fusion.py:116-158
#include "qf_common.cuh"

namespace {

constexpr int MAXK = 128;        // k_cut envelope (reference uses 50)
constexpr int MAXLIST = 1024;    // list length envelope for the bootstrap columns
constexpr int WARPS = 4;
constexpr int NT = WARPS * 32;

struct QppParams {
  const void* scores;
  const int32_t* lens;
  const int32_t* text_feats;
  const double* query_idf;
  const int32_t* perms;
  int n_samples, perm_len;
  int Q, R, K, top_k, k_cut, wig_k;
  double* out;
  int32_t* out_status;
};

// decimal micro-units -> nearest float32 of the DECIMAL value (Float.parseFloat semantics): the
// double quotient is correctly rounded; if it sits exactly on a float32 rounding boundary the sign of
// the exact residual decides.
__device__ __forceinline__ float micro_to_f32(int v) {
  const double d = qf::micro_to_f64(v);       // == v / 1e6 correctly rounded (exhaustively verified), no division
  const float fu = __double2float_ru(d), fd = __double2float_rd(d);
  if (fu == fd) return fu;
  const double mid = 0.5 * ((double)fu + (double)fd);
  if (d > mid) return fu;
  if (d < mid) return fd;
  const double res = fma(d, 1.0e6, -(double)v);   // sign(d - v/1e6)
  if (res > 0) return fd;
  if (res < 0) return fu;
  return __double2float_rn(d);
}

template <int ST> struct Raw;
template <> struct Raw<QF_SCORE_F64> { using type = double; static __device__ float f32(double v) { return __double2float_rn(v); } };
template <> struct Raw<QF_SCORE_F32> { using type = float;  static __device__ float f32(float v) { return v; } };
template <> struct Raw<QF_SCORE_MICRO> { using type = int;  static __device__ float f32(int v) { return micro_to_f32(v); } };

// population variance of x[idx(j)], j < cnt, by the two-pass formula; all lanes get the result
template <typename F>
__device__ __forceinline__ double warp_var(F get, int cnt, int lane, double* mean_out) {
  double s = 0.0;
  for (int j = lane; j < cnt; j += 32) s += get(j);
  s = qf::warp_sum(s);
  const double mean = s / (double)cnt;
  double v = 0.0;
  for (int j = lane; j < cnt; j += 32) {
    const double d = get(j) - mean;
    v += d * d;
  }
  v = qf::warp_sum(v);
  if (mean_out) *mean_out = mean;
  return v / (double)cnt;
}

// RESEARCH = false: the 13 QPPBridge columns only (the pipeline's step 03); no bootstrap buffers, so more warps per SM.
template <int ST, bool RESEARCH>
__global__ void __launch_bounds__(NT) qpp_kernel(const QppParams p) {
  using raw_t = typename Raw<ST>::type;
  constexpr int ML = RESEARCH ? MAXLIST : 4;
  __shared__ double s_x[WARPS][MAXK];          // first k scores (float32-rounded, widened)
  __shared__ float s_all[WARPS][ML];           // whole list as float32 (bootstrap)
  __shared__ unsigned short s_samp[WARPS][RESEARCH ? MAXK : 4];
  __shared__ double s_S[ML + 1];               // RBO suffix sums S[t] = sum_{d = t+1 .. Kd} p^(d-1) / d
  __shared__ double s_uefw[20];                // UEF rank weights 1 / log2(i + 2)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool research = RESEARCH;
  const double p09 = (double)0.9f;              // RBO p is a Java float
  // RBO(original, sample; depth n, p) = (1 - p) * sum_{d=1..n} overlap(d)/d * p^(d-1) with
  // overlap(d) = #{j < m : max(j, b_j) < d}  (OverlapStats.java:77-99).  Exchanging the sums, every sample entry j
  // contributes sum_{d = t_j+1 .. n} p^(d-1)/d with t_j = max(j, b_j): a difference of two suffix sums of a series
  // that depends on nothing but p -- S[t_j] - S[n].  The table is built once per CTA; a sample then costs m look-ups
  // instead of a scan over all n depths (same value up to summation order, ~1e-15 relative).
  if (research) {
    const int Kd = p.K < MAXLIST ? p.K : MAXLIST;
    const double lp = log(p09);
    for (int d = tid; d < Kd; d += NT) s_S[d] = exp((double)d * lp) / (double)(d + 1);     // term of depth d + 1
    if (tid == 0) s_S[Kd] = 0.0;
    __syncthreads();
    if (tid < 32) {                       // in-place suffix sums, 32 entries at a time from the tail
      double carry = 0.0;
      for (int hi = Kd; hi > 0; hi -= 32) {
        const int idx = hi - 1 - (int)tid;                // lane 0 = last entry of the window
        double v = idx >= 0 ? s_S[idx] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double t_ = __shfl_up_sync(0xffffffffu, v, o);
          if ((int)tid >= o) v += t_;
        }
        v += carry;
        if (idx >= 0) s_S[idx] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
  }
  if (tid < 20) s_uefw[tid] = 1.0 / (log((double)(tid + 2)) / log(2.0));
  __syncthreads();

  const long long n_lists = (long long)p.Q * p.R;
  for (long long li = (long long)blockIdx.x * WARPS + warp; li < n_lists; li += (long long)gridDim.x * WARPS) {
    const int q = (int)(li / p.R);
    const raw_t* g = reinterpret_cast<const raw_t*>(p.scores) + li * p.K;
    double* out = p.out + li * QF_QPP_COLS;
    int len = p.lens[li];
    int status = 0;
    if (len < 0 || len > p.K) { status |= QF_QSTAT_BAD_LEN; len = 0; }
    const int n_top = min(len, p.top_k);
    const int k = min(p.k_cut, n_top);
    double* x = s_x[warp];

    // ---- load prefix, validate order
    int unsorted = 0;
    for (int j = lane; j < k; j += 32) x[j] = (double)Raw<ST>::f32(__ldg(g + j));
    if (research) {
      for (int j = lane; j < len; j += 32) s_all[warp][j] = Raw<ST>::f32(__ldg(g + j));
    }
    __syncwarp();
    if (research) {
      for (int j = lane + 1; j < len; j += 32) unsorted |= (s_all[warp][j - 1] < s_all[warp][j]);
    } else {
      for (int j = lane + 1; j < k; j += 32) unsorted |= (x[j - 1] < x[j]);
    }
    unsorted = __any_sync(0xffffffffu, unsorted);
    if (unsorted) status |= QF_QSTAT_UNSORTED;

    const int* tf = p.text_feats + (size_t)q * 4;
    const int n_terms = tf[0], n_unique = tf[1], max_len = tf[2], sum_len = tf[3];
    const int nT = max(1, n_terms);

    double col[QF_QPP_COLS];
#pragma unroll
    for (int c = 0; c < QF_QPP_COLS; ++c) col[c] = 0.0;

    if (k > 0) {
      auto getx = [&](int j) { return x[j]; };
      double mean;
      const double var = warp_var(getx, k, lane, &mean);
      const double kk = (double)k;
      col[0] = var;                                                      // nqc          :202-216
      {                                                                  // smv          :223-239
        const double mu = mean <= 0 ? 1.0 : mean;
        double a = 0.0;
        for (int j = lane; j < k; j += 32) {
          const double s = x[j];
          if (s > 0) a += s * fabs(log(s / mu));
        }
        col[1] = qf::warp_sum(a) / kk;
        // smv_research (SMVSpecificity.java:22-45) skips every term when the mean is <= 0 instead of falling
        // back to mu = 1; for a positive mean the two definitions coincide (avgIDF = 1 without an index)
        col[14] = mean > 0 ? col[1] : 0.0;
      }
      {                                                                  // wig          :247-262
        const double base = 1.0 / fmax(0.01, mean);
        double a = 0.0;
        for (int j = lane; j < k; j += 32) a += x[j] - base;
        col[2] = qf::warp_sum(a) / ((double)nT * kk);
      }
      {                                                                  // SigmaMax :269-293, cumnqc :400-410
        // Population variance of every prefix x[0:i], i = 2..k.  The Java recomputes each window from scratch
        // (O(k^2)); here one warp scan of y = x - x[0] and y^2 gives var_i = (S2_i - S_i^2 / i) / i.  Every window
        // contains x[0] (y = 0), so the window's spread is at least its largest |y| and the subtraction loses at
        // most a factor ~2i of precision: relative error ~1e-14 (parity tolerance 1e-9); all-equal prefixes give an
        // exact 0 as the two-pass formula does.
        const double x0 = x[0];
        double mx = 0.0, sm = 0.0, cs = 0.0, cs2 = 0.0;
        for (int base = 0; base < k; base += 32) {
          const int j = base + lane;
          const double y = j < k ? x[j] - x0 : 0.0;
          double a = y, b = y * y;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += ta; b += tb; }
          }
          a += cs;
          b += cs2;
          if (j >= 1 && j < k) {                                         // window i = j + 1
            const double n_i = (double)(j + 1);
            const double v = fmax((b - a * a / n_i) / n_i, 0.0);
            mx = fmax(mx, v);
            sm += v;
          }
          cs = __shfl_sync(0xffffffffu, a, 31);
          cs2 = __shfl_sync(0xffffffffu, b, 31);
        }
        mx = qf::warp_max(mx);
        sm = qf::warp_sum(sm);
        if (k >= 2) {
          col[3] = sqrt(mx) / sqrt((double)max(1, nT));
          col[9] = sm / (double)(k - 1);
        }
      }
      {                                                                  // SigmaX       :299-319
        const double thr = x[0] * 0.5;
        int c = 0;
        double s = 0.0;
        for (int j = lane; j < k; j += 32)
          if (x[j] >= thr) { ++c; s += x[j]; }
        c = __reduce_add_sync(0xffffffffu, c);
        s = qf::warp_sum(s);
        double sig = 0.0;
        if (c >= 1) {
          const double mu = s / (double)c;
          double v = 0.0;
          for (int j = lane; j < k; j += 32)
            if (x[j] >= thr) { const double d = x[j] - mu; v += d * d; }
          sig = sqrt(qf::warp_sum(v) / (double)c);
        }
        col[4] = (k >= 2 && c >= 2) ? sig : 0.0;
        col[16] = sig / sqrt((double)n_unique);                          // sigmax_research SigmaXSpecificity.java:19-52
      }
      {                                                                  // RSD skew     :324-342
        const double sd = sqrt(var);
        double a = 0.0;
        if (k >= 3 && !(sd < 1e-10)) {
          for (int j = lane; j < k; j += 32) {
            const double z = (x[j] - mean) / sd;
            a += z * z * z;
          }
          col[5] = qf::warp_sum(a) / kk;
        }
      }
      {                                                                  // UEF          :349-363
        const int kk20 = min(20, k);
        double u = 0.0, ws = 0.0;
        for (int i = lane; i < kk20; i += 32) {
          const double w = s_uefw[i];
          u += x[i] * w;
          ws += w;
        }
        col[6] = qf::warp_sum(u) / qf::warp_sum(ws);
      }
      {                                                                  // snqc         :417-438
        double a = 0.0;
        if (mean > 0) {
          for (int j = lane; j < k; j += 32) {
            const double r = x[j];
            if (r > 0) {
              const double f2 = ((r - mean) * (r - mean)) / r;
              // pow(pow(1, a) * pow(f2, b), g) with a = b = g = 0.33: pow(1, a) == 1 exactly; the nested power is
              // evaluated as one pow with exponent b*g (differs from the nested form by <= 2 ulp; tolerance 1e-5)
              a += pow(f2, 0.33 * 0.33);
            }
          }
          col[10] = qf::warp_sum(a) / kk;
        }
      }
      {                                                                  // dense-qpp(-m) :444-453
        double mn = INFINITY, mx = -INFINITY;
        for (int j = lane; j < k; j += 32) { mn = fmin(mn, x[j]); mx = fmax(mx, x[j]); }
        mn = qf::warp_min(mn);
        mx = qf::warp_max(mx);
        const double range = mx - mn;
        col[11] = col[12] = (k >= 2 && range > 0) ? log(1.0 + 1.0 / range) : 0.0;
      }
      // text-only proxies                                                :368-393
      col[7] = log(1.0 + (double)n_unique) + log(1.0 + (double)max_len) * 0.5;
      {
        const double avg_len = n_terms > 0 ? (double)sum_len / (double)n_terms : 3.0;
        const double diversity = n_terms > 0 ? (double)n_unique / (double)n_terms : __longlong_as_double(0x7ff8000000000000ll);
        col[8] = log(1.0 + avg_len) * diversity;
      }

      if (research) {
        const double max_idf = p.query_idf ? p.query_idf[q] : 1.0;
        const float* all = s_all[warp];
        {                                                                // wig_research  WIGSpecificity.java:18-51
          const int cutoff = min(p.wig_k, len);
          const double base = 1.0 / max_idf;
          double a = 0.0;
          for (int i = 0; i < cutoff; ++i) a += (double)all[i] - base;   // tiny, sequential like the reference
          col[15] = a / (double)(n_unique * cutoff);
        }
        {                                                                // scnqc_research NQCCalibratedSpecificity.java:31-57
          const double e = (double)0.33f;
          double a = 0.0;
          for (int j = lane; j < k; j += 32) {
            const double r = x[j];
            const double f2 = (r - mean) * (r - mean) / r;
            a += pow(pow(max_idf, e) * pow(f2, e), e);
          }
          col[17] = (qf::warp_sum(a) / kk) * max_idf;
        }
        {                                                                // rsd_bootstrap  RSDSpecificity.java:22-52
          const int mS = min(len, p.k_cut);
          unsigned short* samp = s_samp[warp];
          double acc = 0.0;
          for (int s = 0; s < p.n_samples; ++s) {
            const int32_t* perm = p.perms + (size_t)s * p.perm_len;
            // first mS entries of the permutation that are < len
            int got = 0;
            for (int base = 0; base < p.perm_len && got < mS; base += 32) {
              const int e = base + lane < p.perm_len ? __ldg(perm + base + lane) : 0x7fffffff;
              const bool ok = e >= 0 && e < len;
              const unsigned bal = __ballot_sync(0xffffffffu, ok);
              const int pos = got + __popc(bal & ((1u << lane) - 1u));
              if (ok && pos < mS) samp[pos] = (unsigned short)e;
              got += __popc(bal);
            }
            __syncwarp();
            auto gets = [&](int j) { return (double)all[samp[j]]; };
            const double est = warp_var(gets, mS, lane, nullptr);       // NQC of the sample
            // RBO(original, sample; depth len, p = 0.9f) through the suffix table (see the kernel prologue)
            const double s_len = s_S[len];
            double aggr = 0.0;
            for (int j = lane; j < mS; j += 32) {
              const int t = max(j, (int)samp[j]);
              aggr += s_S[t] - s_len;
            }
            aggr = qf::warp_sum(aggr);
            const double one_minus_p = (double)(1.0f - 0.9f);
            acc += (one_minus_p * aggr) * est;
            __syncwarp();
          }
          col[13] = p.n_samples > 0 ? acc / (double)p.n_samples : 0.0;
        }
      }
    }
    if (!research) {
#pragma unroll
      for (int c = QF_QPP_BRIDGE_COLS; c < QF_QPP_COLS; ++c) col[c] = 0.0;
    }
    // every lane holds identical col[]: spread the stores over lanes
#pragma unroll
    for (int c = 0; c < QF_QPP_COLS; ++c)
      if (lane == c) out[c] = col[c];
    if (p.out_status && lane == 31) p.out_status[li] = status;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Bridge-only kernel (the pipeline's step 03: the 13 QPPBridge columns; research columns are written as 0).
// EIGHT lanes per list, four lists per warp: k_cut = 50 scores are 7 per lane, and every reduction is 3 shuffle
// steps inside the 8-lane group instead of 5 across a warp that is 36 % idle in its second pass.  Groups are
// independent (shuffles and syncs use the group's lane mask), so lists of different length may diverge freely.
// Same formulas as qpp_kernel above (QPPBridge.java:134-453); only the summation grouping differs (tolerance 1e-9).
// ------------------------------------------------------------------------------------------------
constexpr int LPL = 8;                 // lanes per list
constexpr int GPW = 32 / LPL;          // lists per warp

__device__ __forceinline__ double gsum(unsigned m, double v) {
#pragma unroll
  for (int o = LPL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
  return v;
}
__device__ __forceinline__ double gmax(unsigned m, double v) {
#pragma unroll
  for (int o = LPL / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(m, v, o));
  return v;
}
__device__ __forceinline__ double gmin(unsigned m, double v) {
#pragma unroll
  for (int o = LPL / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(m, v, o));
  return v;
}

template <int ST>
__global__ void __launch_bounds__(NT) qpp_bridge_kernel(const QppParams p) {
  using raw_t = typename Raw<ST>::type;
  __shared__ double s_x[WARPS * GPW][MAXK];    // first k scores of each list in flight (float32-rounded, widened)
  __shared__ double s_uefw[20];                // UEF rank weights 1 / log2(i + 2)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sl = lane & (LPL - 1), gi = lane / LPL;
  const unsigned gm = ((1u << LPL) - 1u) << (gi * LPL);
  if (tid < 20) s_uefw[tid] = 1.0 / (log((double)(tid + 2)) / log(2.0));
  __syncthreads();

  double* x = s_x[warp * GPW + gi];
  const long long n_lists = (long long)p.Q * p.R;
  const long long stride = (long long)gridDim.x * WARPS * GPW;
  for (long long li = ((long long)blockIdx.x * WARPS + warp) * GPW + gi; li < n_lists; li += stride) {
    const int q = (int)(li / p.R);
    const raw_t* g = reinterpret_cast<const raw_t*>(p.scores) + li * p.K;
    double* out = p.out + li * QF_QPP_COLS;
    int len = p.lens[li];
    int status = 0;
    if (len < 0 || len > p.K) { status |= QF_QSTAT_BAD_LEN; len = 0; }
    const int k = min(p.k_cut, min(len, p.top_k));

    __syncwarp(gm);                             // previous list's x is dead
    for (int j = sl; j < k; j += LPL) x[j] = (double)Raw<ST>::f32(__ldg(g + j));
    __syncwarp(gm);
    int unsorted = 0;
    for (int j = sl + 1; j < k; j += LPL) unsorted |= (x[j - 1] < x[j]);
    if (__any_sync(gm, unsorted)) status |= QF_QSTAT_UNSORTED;

    const int* tf = p.text_feats + (size_t)q * 4;
    const int n_terms = tf[0], n_unique = tf[1], max_len = tf[2], sum_len = tf[3];
    const int nT = max(1, n_terms);

    double col[QF_QPP_BRIDGE_COLS];
#pragma unroll
    for (int c = 0; c < QF_QPP_BRIDGE_COLS; ++c) col[c] = 0.0;

    if (k > 0) {
      const double kk = (double)k;
      // one pass: sum, min, max, count/sum above half of the top score
      const double thr = x[0] * 0.5;
      double s1 = 0.0, mn = INFINITY, mx = -INFINITY, sx = 0.0;
      int cx = 0;
      for (int j = sl; j < k; j += LPL) {
        const double v = x[j];
        s1 += v;
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        if (v >= thr) { ++cx; sx += v; }
      }
      s1 = gsum(gm, s1);
      mn = gmin(gm, mn);
      mx = gmax(gm, mx);
      sx = gsum(gm, sx);
      cx = __reduce_add_sync(gm, cx);
      const double mean = s1 / kk;
      const double mu_smv = mean <= 0 ? 1.0 : mean;
      const double wig_base = 1.0 / fmax(0.01, mean);
      const double mu_x = cx >= 1 ? sx / (double)cx : 0.0;
      // second pass: everything that needs the mean(s)
      double vsum = 0.0, smv = 0.0, wig = 0.0, vx = 0.0, snqc = 0.0;
      for (int j = sl; j < k; j += LPL) {
        const double v = x[j];
        const double d = v - mean;
        vsum += d * d;
        if (v > 0) smv += v * fabs(log(v / mu_smv));
        wig += v - wig_base;
        if (v >= thr) { const double dx = v - mu_x; vx += dx * dx; }
        if (mean > 0 && v > 0) snqc += pow((d * d) / v, 0.33 * 0.33);     // see qpp_kernel: one pow, <= 2 ulp
      }
      vsum = gsum(gm, vsum);
      smv = gsum(gm, smv);
      wig = gsum(gm, wig);
      vx = gsum(gm, vx);
      snqc = gsum(gm, snqc);
      const double var = vsum / kk;
      col[0] = var;                                                      // nqc          :202-216
      col[1] = smv / kk;                                                 // smv          :223-239
      col[2] = wig / ((double)nT * kk);                                  // wig          :247-262
      {                                                                  // SigmaMax :269-293, cumnqc :400-410
        const double x0 = x[0];
        double wmx = 0.0, wsm = 0.0, cs = 0.0, cs2 = 0.0;
        for (int base = 0; base < k; base += LPL) {
          const int j = base + sl;
          const double y = j < k ? x[j] - x0 : 0.0;
          double a = y, b = y * y;
#pragma unroll
          for (int o = 1; o < LPL; o <<= 1) {
            const double ta = __shfl_up_sync(gm, a, o, LPL), tb = __shfl_up_sync(gm, b, o, LPL);
            if (sl >= o) { a += ta; b += tb; }
          }
          a += cs;
          b += cs2;
          if (j >= 1 && j < k) {                                         // window i = j + 1
            const double n_i = (double)(j + 1);
            const double v = fmax((b - a * a / n_i) / n_i, 0.0);
            wmx = fmax(wmx, v);
            wsm += v;
          }
          cs = __shfl_sync(gm, a, LPL - 1, LPL);
          cs2 = __shfl_sync(gm, b, LPL - 1, LPL);
        }
        wmx = gmax(gm, wmx);
        wsm = gsum(gm, wsm);
        if (k >= 2) {
          col[3] = sqrt(wmx) / sqrt((double)max(1, nT));
          col[9] = wsm / (double)(k - 1);
        }
      }
      {                                                                  // SigmaX       :299-319
        const double sig = cx >= 1 ? sqrt(vx / (double)cx) : 0.0;
        col[4] = (k >= 2 && cx >= 2) ? sig : 0.0;
      }
      {                                                                  // RSD skew     :324-342
        const double sd = sqrt(var);
        if (k >= 3 && !(sd < 1e-10)) {
          double a = 0.0;
          for (int j = sl; j < k; j += LPL) {
            const double z = (x[j] - mean) / sd;
            a += z * z * z;
          }
          col[5] = gsum(gm, a) / kk;
        }
      }
      {                                                                  // UEF          :349-363
        const int kk20 = min(20, k);
        double u = 0.0, ws = 0.0;
        for (int i = sl; i < kk20; i += LPL) {
          const double w = s_uefw[i];
          u += x[i] * w;
          ws += w;
        }
        col[6] = gsum(gm, u) / gsum(gm, ws);
      }
      if (mean > 0) col[10] = snqc / kk;                                 // snqc         :417-438
      {                                                                  // dense-qpp(-m) :444-453
        const double range = mx - mn;
        col[11] = col[12] = (k >= 2 && range > 0) ? log(1.0 + 1.0 / range) : 0.0;
      }
      // text-only proxies                                                :368-393
      col[7] = log(1.0 + (double)n_unique) + log(1.0 + (double)max_len) * 0.5;
      {
        const double avg_len = n_terms > 0 ? (double)sum_len / (double)n_terms : 3.0;
        const double diversity = n_terms > 0 ? (double)n_unique / (double)n_terms : __longlong_as_double(0x7ff8000000000000ll);
        col[8] = log(1.0 + avg_len) * diversity;
      }
    }
    // every lane of the group holds identical col[]: spread the stores over its lanes
#pragma unroll
    for (int c = 0; c < QF_QPP_COLS; ++c)
      if (sl == (c & (LPL - 1))) out[c] = c < QF_QPP_BRIDGE_COLS ? col[c < QF_QPP_BRIDGE_COLS ? c : 0] : 0.0;
    if (p.out_status && sl == LPL - 1) p.out_status[li] = status;
  }
}

// ------------------------------------------------------------------------------------------------
// column normalisation
// ------------------------------------------------------------------------------------------------
struct ColStat {
  unsigned long long omin, omax;
  double sum, sumsq;
};

__global__ void colnorm_init(ColStat* st, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { st[i].omin = ~0ull; st[i].omax = 0ull; st[i].sum = 0.0; st[i].sumsq = 0.0; }
}

// pass 1: min / max / sum per (r, col); pass 2: sum of squared deviations.
// A CTA owns a contiguous range of rows; thread (row group, column) keeps its column's statistics in registers while
// it walks the rows (loads coalesced across columns), so shared and global atomics are per (row group, column), not
// per element.
template <int PASS>
__global__ void __launch_bounds__(256) colnorm_reduce(const double* vals, long long Q, int RC, ColStat* st) {
  extern __shared__ unsigned char sm_raw[];
  ColStat* loc = reinterpret_cast<ColStat*>(sm_raw);
  for (int i = threadIdx.x; i < RC; i += blockDim.x) { loc[i].omin = ~0ull; loc[i].omax = 0ull; loc[i].sum = 0.0; loc[i].sumsq = 0.0; }
  __syncthreads();
  const int cpt = RC < (int)blockDim.x ? RC : (int)blockDim.x;     // columns per sweep
  const int n_rg = (int)blockDim.x / cpt;                           // row groups
  const int c_in = (int)threadIdx.x % cpt, rg = (int)threadIdx.x / cpt;
  const long long rows_per = (Q + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * rows_per;
  const long long q1 = q0 + rows_per < Q ? q0 + rows_per : Q;
  if (rg < n_rg) {
    for (int c = c_in; c < RC; c += cpt) {
      unsigned long long omin = ~0ull, omax = 0ull;
      double sum = 0.0;
      const double mean = PASS == 2 ? st[c].sum / (double)Q : 0.0;
      bool any = false;
      for (long long q = q0 + rg; q < q1; q += n_rg) {
        const double v = vals[q * RC + c];
        any = true;
        if (PASS == 1) {
          const unsigned long long o = qf::ord_f64(v + 0.0);
          omin = o < omin ? o : omin;
          omax = o > omax ? o : omax;
          sum += v;
        } else {
          const double d = v - mean;
          sum += d * d;
        }
      }
      if (any) {
        if (PASS == 1) {
          atomicMin(&loc[c].omin, omin);
          atomicMax(&loc[c].omax, omax);
          atomicAdd(&loc[c].sum, sum);
        } else {
          atomicAdd(&loc[c].sumsq, sum);
        }
      }
    }
  }
  __syncthreads();
  if (q0 < q1) {
    for (int i = threadIdx.x; i < RC; i += blockDim.x) {
      if (PASS == 1) {
        atomicMin(&st[i].omin, loc[i].omin);
        atomicMax(&st[i].omax, loc[i].omax);
        atomicAdd(&st[i].sum, loc[i].sum);
      } else {
        atomicAdd(&st[i].sumsq, loc[i].sumsq);
      }
    }
  }
}

__device__ __forceinline__ double unord_f64(unsigned long long o) {
  const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)b);
}

// float(f"{v:.6f}"): exact round-half-even of the binary value to 6 decimals, then the nearest double.
__device__ __forceinline__ double round6(double v) {
  if (!(fabs(v) < 1.0e9)) return v;
  const double pr = __dmul_rn(v, 1.0e6);
  const double err = fma(v, 1.0e6, -pr);       // v*1e6 == pr + err exactly
  double r = rint(pr);
  const double a = pr - r;                      // exact, in [-0.5, 0.5]
  if (a > 0) {
    const double c = 0.5 - a;                   // distance to the upper tie (exact)
    if (err > c) r += 1.0;
    else if (err == c && fmod(r, 2.0) != 0.0) r += 1.0;
  } else if (a < 0) {
    const double c = 0.5 + a;                   // distance to the lower tie (exact)
    if (-err > c) r -= 1.0;
    else if (-err == c && fmod(r, 2.0) != 0.0) r -= 1.0;
  }
  // a == 0: pr is an integer; |err| < 0.5 so r stands.  a == +-0.5 is covered above (c == 0).
  return __ddiv_rn(r + 0.0, 1.0e6);
}

__global__ void __launch_bounds__(256) colnorm_apply(double* vals, long long Q, int RC, const ColStat* st, int mode, int do_round6) {
  const long long total = Q * RC;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % RC);
    double v = vals[e];
    if (mode == QF_COLNORM_MINMAX) {
      const double vmin = unord_f64(st[c].omin), vmax = unord_f64(st[c].omax);
      const double scale = vmax > vmin ? __dsub_rn(vmax, vmin) : 1.0;
      v = __ddiv_rn(__dsub_rn(v, vmin), scale);
    } else if (mode == QF_COLNORM_ZSCORE) {
      const double mean = st[c].sum / (double)Q;
      const double sd = sqrt(st[c].sumsq / (double)Q);
      v = (v - mean) / (sd > 0 ? sd : 1.0);
    }
    if (do_round6) v = round6(v);
    vals[e] = v;
  }
}

__global__ void colstats_export(const ColStat* st, int RC, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < RC) {
    out[4 * i + 0] = unord_f64(st[i].omin);
    out[4 * i + 1] = unord_f64(st[i].omax);
    out[4 * i + 2] = st[i].sum;
    out[4 * i + 3] = st[i].sumsq;
  }
}

template <int PASS2>
__global__ void __launch_bounds__(256) colstats_sumsq(const double* vals, long long Q, int RC, const double* mean, ColStat* st) {
  extern __shared__ unsigned char sm_raw[];
  double* loc = reinterpret_cast<double*>(sm_raw);
  for (int i = threadIdx.x; i < RC; i += blockDim.x) loc[i] = 0.0;
  __syncthreads();
  const long long total = Q * RC;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % RC);
    const double d = vals[e] - mean[c];
    atomicAdd(&loc[c], d * d);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RC; i += blockDim.x) atomicAdd(&st[i].sumsq, loc[i]);
}

__global__ void __launch_bounds__(256) colnorm_apply_params(double* vals, long long Q, int RC, const double* base,
                                                            const double* scale, int do_round6) {
  const long long total = Q * RC;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % RC);
    double v = vals[e];
    if (base) v = __ddiv_rn(__dsub_rn(v, base[c]), scale[c]);
    if (do_round6) v = round6(v);
    vals[e] = v;
  }
}

__global__ void __launch_bounds__(256) select_weight_kernel(const double* vals, long long n_lists, int ncols, int col,
                                                            int ncols_mean, double* weights) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lists) return;
  const double* v = vals + i * ncols;
  if (col >= 0) {
    weights[i] = v[col];
  } else {
    double s = 0.0;                                  // Python sum(): left to right from 0
    for (int c = 0; c < ncols_mean; ++c) s = __dadd_rn(s, v[c]);
    weights[i] = __ddiv_rn(s, (double)ncols_mean);
  }
}

}  // namespace

extern "C" int qf_qpp_stats(const void* scores, int score_type, const int32_t* lens, const int32_t* text_feats,
                            const double* query_idf, const int32_t* perms, int n_samples, int perm_len, int Q, int R,
                            int K, int top_k, int k_cut, int wig_k, double* out, int32_t* out_status, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID, "qf_qpp_stats: bad sizes Q=%d R=%d K=%d", Q, R, K);
  QF_REQUIRE(k_cut >= 1 && k_cut <= MAXK, QF_E_UNSUPPORTED, "qf_qpp_stats: k_cut=%d outside [1,%d]", k_cut, MAXK);
  QF_REQUIRE(top_k >= 1 && wig_k >= 1, QF_E_INVALID, "qf_qpp_stats: top_k and wig_k must be >= 1");
  QF_REQUIRE(score_type >= QF_SCORE_F64 && score_type <= QF_SCORE_MICRO, QF_E_INVALID,
             "qf_qpp_stats: unknown score_type %d", score_type);
  if (perms) {
    QF_REQUIRE(K <= MAXLIST, QF_E_UNSUPPORTED, "qf_qpp_stats: bootstrap columns need K <= %d (got %d)", MAXLIST, K);
    QF_REQUIRE(n_samples >= 1 && perm_len >= K, QF_E_INVALID,
               "qf_qpp_stats: perms need n_samples >= 1 and perm_len >= K (got %d, %d)", n_samples, perm_len);
  }
  if (Q == 0) return QF_OK;
  QF_REQUIRE(scores && lens && text_feats && out, QF_E_INVALID, "qf_qpp_stats: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  QppParams p{scores, lens, text_feats, query_idf, perms, n_samples, perm_len, Q, R, K, top_k, k_cut, wig_k, out, out_status};
  const long long n_lists = (long long)Q * R;
  long long blocks = (n_lists + WARPS - 1) / WARPS;
  const long long cap = (long long)qf::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned nb = (unsigned)blocks;
  qf::prof_mark(QF_PROF_QPP_STATS, 0, st);
  if (perms) {
    switch (score_type) {
      case QF_SCORE_F64: qpp_kernel<QF_SCORE_F64, true><<<nb, NT, 0, st>>>(p); break;
      case QF_SCORE_F32: qpp_kernel<QF_SCORE_F32, true><<<nb, NT, 0, st>>>(p); break;
      default: qpp_kernel<QF_SCORE_MICRO, true><<<nb, NT, 0, st>>>(p); break;
    }
  } else {
    long long gb = (n_lists + WARPS * GPW - 1) / (WARPS * GPW);
    if (gb > cap) gb = cap;
    switch (score_type) {
      case QF_SCORE_F64: qpp_bridge_kernel<QF_SCORE_F64><<<(unsigned)gb, NT, 0, st>>>(p); break;
      case QF_SCORE_F32: qpp_bridge_kernel<QF_SCORE_F32><<<(unsigned)gb, NT, 0, st>>>(p); break;
      default: qpp_bridge_kernel<QF_SCORE_MICRO><<<(unsigned)gb, NT, 0, st>>>(p); break;
    }
  }
  qf::prof_mark(QF_PROF_QPP_STATS, 1, st);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

extern "C" size_t qf_qpp_colnorm_workspace_bytes(int R, int ncols) {
  if (R < 1 || ncols < 1) return 0;
  return sizeof(ColStat) * (size_t)R * (size_t)ncols;
}

extern "C" int qf_qpp_colnorm(double* vals, int Q, int R, int ncols, int mode, int round6_flag, void* ws,
                              size_t ws_bytes, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && ncols >= 1, QF_E_INVALID, "qf_qpp_colnorm: bad sizes");
  QF_REQUIRE(mode >= 0 && mode <= QF_COLNORM_ZSCORE, QF_E_INVALID, "qf_qpp_colnorm: unknown mode %d", mode);
  const int RC = R * ncols;
  QF_REQUIRE(RC <= 1024, QF_E_UNSUPPORTED, "qf_qpp_colnorm: R*ncols=%d exceeds 1024", RC);
  if (Q == 0) return QF_OK;
  QF_REQUIRE(vals && ws, QF_E_INVALID, "qf_qpp_colnorm: null pointer argument");
  QF_REQUIRE(ws_bytes >= qf_qpp_colnorm_workspace_bytes(R, ncols), QF_E_WORKSPACE,
             "qf_qpp_colnorm: workspace %zu < %zu bytes", ws_bytes, qf_qpp_colnorm_workspace_bytes(R, ncols));
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ColStat* cs = reinterpret_cast<ColStat*>(ws);
  const long long total = (long long)Q * RC;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)qf::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  long long rblocks = ((long long)Q + 15) / 16;          // reduce: a contiguous row range per CTA
  if (rblocks > qf::sm_count()) rblocks = qf::sm_count();
  if (mode != 0) {
    colnorm_init<<<(RC + 255) / 256, 256, 0, st>>>(cs, RC);
    colnorm_reduce<1><<<(unsigned)rblocks, 256, sizeof(ColStat) * RC, st>>>(vals, Q, RC, cs);
    if (mode == QF_COLNORM_ZSCORE)
      colnorm_reduce<2><<<(unsigned)rblocks, 256, sizeof(ColStat) * RC, st>>>(vals, Q, RC, cs);
  }
  if (mode != 0 || round6_flag)
    colnorm_apply<<<(unsigned)blocks, 256, 0, st>>>(vals, Q, RC, cs, mode, round6_flag);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

extern "C" int qf_qpp_select_weight(const double* vals, int Q, int R, int ncols, int col, int ncols_mean,
                                    double* weights, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && ncols >= 1, QF_E_INVALID, "qf_qpp_select_weight: bad sizes");
  QF_REQUIRE(col >= -1 && col < ncols, QF_E_INVALID, "qf_qpp_select_weight: column %d out of range [-1,%d)", col, ncols);
  QF_REQUIRE(col >= 0 || (ncols_mean >= 1 && ncols_mean <= ncols), QF_E_INVALID,
             "qf_qpp_select_weight: ncols_mean=%d out of range", ncols_mean);
  if (Q == 0) return QF_OK;
  QF_REQUIRE(vals && weights, QF_E_INVALID, "qf_qpp_select_weight: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const long long n = (long long)Q * R;
  select_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      vals, n, ncols, col, ncols_mean, weights);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

extern "C" int qf_qpp_colstats(const double* vals, int Q, int R, int ncols, const double* mean, double* stats, void* ws,
                               size_t ws_bytes, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && ncols >= 1, QF_E_INVALID, "qf_qpp_colstats: bad sizes");
  const int RC = R * ncols;
  QF_REQUIRE(RC <= 1024, QF_E_UNSUPPORTED, "qf_qpp_colstats: R*ncols=%d exceeds 1024", RC);
  QF_REQUIRE(stats && ws && (vals || Q == 0), QF_E_INVALID, "qf_qpp_colstats: null pointer argument");
  QF_REQUIRE(ws_bytes >= qf_qpp_colnorm_workspace_bytes(R, ncols), QF_E_WORKSPACE, "qf_qpp_colstats: workspace too small");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ColStat* cs = reinterpret_cast<ColStat*>(ws);
  const long long total = (long long)Q * RC;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)qf::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  colnorm_init<<<(RC + 255) / 256, 256, 0, st>>>(cs, RC);
  if (Q > 0) {
    long long rblocks = ((long long)Q + 15) / 16;
    if (rblocks > qf::sm_count()) rblocks = qf::sm_count();
    colnorm_reduce<1><<<(unsigned)rblocks, 256, sizeof(ColStat) * RC, st>>>(vals, Q, RC, cs);
    if (mean) colstats_sumsq<1><<<(unsigned)blocks, 256, sizeof(double) * RC, st>>>(vals, Q, RC, mean, cs);
  }
  colstats_export<<<(RC + 255) / 256, 256, 0, st>>>(cs, RC, stats);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

extern "C" int qf_qpp_colnorm_apply(double* vals, int Q, int R, int ncols, const double* base, const double* scale,
                                    int round6_flag, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && ncols >= 1, QF_E_INVALID, "qf_qpp_colnorm_apply: bad sizes");
  QF_REQUIRE((base == nullptr) == (scale == nullptr), QF_E_INVALID, "qf_qpp_colnorm_apply: base and scale go together");
  if (Q == 0) return QF_OK;
  QF_REQUIRE(vals, QF_E_INVALID, "qf_qpp_colnorm_apply: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const int RC = R * ncols;
  const long long total = (long long)Q * RC;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)qf::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  colnorm_apply_params<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(vals, Q, RC, base, scale, round6_flag);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}
