// Linear SVM fit of the late-fusion step: replaces `svm.LinearSVC().fit(X, y)` at reference
// Sheet03/combinedModel.py:34-35, i.e. LIBLINEAR's L2-regularised L2-loss dual coordinate descent, one-vs-rest.
// The statement of the algorithm (and what deliberately differs from LIBLINEAR: a deterministic visiting order, shrunk
// samples removed at the end of an epoch instead of on the spot) is in oracle/svm_fit.py; this file follows it step for
// step, in fp64.
//
// Mapping: coordinate descent is sequential in the samples of ONE class problem (every step reads the w the previous
// step wrote) and the classes are independent, so each class is one warp: lane l keeps w[l], w[l+32], ... of the
// 2D+1 = 513 weights in registers, a step is 16 FMAs per lane, an xor-butterfly (every lane ends with the same sum, in
// the same order), a handful of scalar operations done redundantly by all lanes, and 16 FMAs for the update.  The next
// sample's row (4 KB, L2-resident: X is 15 MB) and its alpha / Q_ii are loaded before the current step's arithmetic.
// 101 warps on 101 SMs is all the parallelism the reference's algorithm has; the point of running it here is that the
// fused descriptors never leave the device between evaluation and fit/predict.
#include "va_internal.h"

namespace va {

__device__ __forceinline__ uint32_t svm_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t svm_gcd(uint32_t a, uint32_t b) {
  while (b != 0) { const uint32_t t = a % b; a = b; b = t; }
  return a;
}
// visiting order of an epoch: i -> (a * i + b) mod n with gcd(a, n) == 1 (oracle/svm_fit.py::epoch_order_params)
__device__ __forceinline__ void epoch_order_params(uint32_t epoch, uint32_t n, uint32_t& a, uint32_t& b) {
  const uint32_t h = svm_mix32(epoch * 0x9E3779B1u + 0x7F4A7C15u);
  a = h % n;
  while (svm_gcd(a, n) != 1) a = (a + 1) % n;
  b = svm_mix32(h ^ 0x85EBCA6Bu) % n;
}

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Q_ii + D_ii = x_i.x_i + bias^2 + 1/(2C); one warp per sample
__global__ void svm_qd_kernel(const double* __restrict__ X, int V, int F, double bias2, double D, double* __restrict__ qd) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= V) return;
  double s = 0.0;
  for (int f = lane; f < F; f += 32) { const double x = X[(size_t)i * F + f]; s = fma(x, x, s); }
  s = warp_sum_all(s);
  if (lane == 0) qd[i] = s + bias2 + D;
}

template <int KMAX>
__global__ void __launch_bounds__(32) svm_dcd_kernel(const double* __restrict__ X, int V, int F,
                                                     const int32_t* __restrict__ class_index, int first_class,
                                                     const double* __restrict__ qd, double D, double bias, double tol,
                                                     int max_iter, double* __restrict__ alpha, int32_t* __restrict__ active,
                                                     double* __restrict__ coef, double* __restrict__ intercept,
                                                     int32_t* __restrict__ epochs) {
  const int k = blockIdx.x, cls = k + first_class, lane = threadIdx.x;
  double* al = alpha + (size_t)k * V;          // zeroed by the caller
  int32_t* idx = active + (size_t)k * V;       // the active set, in visiting-position order
  double w[KMAX], x[KMAX], xn[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) w[j] = 0.0;
  double wb = 0.0;
  for (int i = lane; i < V; i += 32) idx[i] = i;
  __syncwarp();
  uint32_t n = (uint32_t)V;                    // active samples
  double pgmax_old = INFINITY;
  int it = 0;
  while (it < max_iter) {
    double pgmax = -INFINITY, pgmin = INFINITY;
    bool any_left = false;
    if (n > 0) {
      uint32_t a, b;
      epoch_order_params((uint32_t)it, n, a, b);
      // software pipeline: sample index two steps ahead, its row / alpha / Q_ii / label one step ahead
      double a_n = 0.0, q_n = 1.0, y_n = 1.0;
      auto load_row = [&](int32_t s) {
        const double* row = X + (size_t)s * F;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) xn[j] = (lane + 32 * j < F) ? row[lane + 32 * j] : 0.0;
        a_n = al[s];
        q_n = qd[s];
        y_n = (class_index[s] == cls) ? 1.0 : -1.0;
      };
      uint32_t pos1 = b;                       // position of step i + 1 while step i runs (b < n)
      int32_t s1 = idx[pos1];
      load_row(s1);
      uint32_t pos2 = pos1 + a;                // (a * i + b) mod n without a division: a, pos < n
      if (pos2 >= n) pos2 -= n;
      int32_t s2 = (n > 1) ? idx[pos2] : 0;
      for (uint32_t i = 0; i < n; ++i) {
        const int32_t cur = s1;
        const uint32_t curpos = pos1;
        const double ai = a_n, qi = q_n, yi = y_n;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) x[j] = xn[j];
        if (i + 1 < n) {
          s1 = s2;
          pos1 = pos2;
          load_row(s1);
          if (i + 2 < n) {
            pos2 += a;
            if (pos2 >= n) pos2 -= n;
            s2 = idx[pos2];
          }
        }
        // the step's critical path: up to four independent FMA chains, then the butterfly
        constexpr int NACC = KMAX >= 4 ? 4 : KMAX;
        double acc[NACC];
#pragma unroll
        for (int j = 0; j < NACC; ++j) acc[j] = w[j] * x[j];
#pragma unroll
        for (int j = NACC; j < KMAX; ++j) acc[j % NACC] = fma(w[j], x[j], acc[j % NACC]);
        double dot = acc[0];
#pragma unroll
        for (int j = 1; j < NACC; ++j) dot += acc[j];
        dot = warp_sum_all(dot);
        const double G = yi * (dot + wb * bias) - 1.0 + ai * D;
        if (ai == 0.0 && G > pgmax_old) {       // shrinking: sits out until the active set has converged
          idx[curpos] = ~cur;                   // (every lane stores the same value) removed after the epoch
          any_left = true;
          continue;
        }
        const double PG = (ai > 0.0) ? G : fmin(G, 0.0);
        pgmax = fmax(pgmax, PG);
        pgmin = fmin(pgmin, PG);
        if (fabs(PG) > 1e-12) {
          const double nw = fmax(ai - G / qi, 0.0);
          const double d = (nw - ai) * yi;
          al[cur] = nw;                         // every lane stores the same value: each lane later reads its own store
#pragma unroll
          for (int j = 0; j < KMAX; ++j) w[j] = fma(d, x[j], w[j]);
          wb = fma(d, bias, wb);
        }
      }
    }
    ++it;
    if (pgmax - pgmin <= tol) {                 // (-inf - inf for an empty active set)
      if (n == (uint32_t)V && !any_left) break;   // the range was taken over ALL samples
      __syncwarp();
      for (int i = lane; i < V; i += 32) idx[i] = i;   // converged on the active set: check again on all samples
      __syncwarp();
      n = (uint32_t)V;
      pgmax_old = INFINITY;
      continue;
    }
    if (any_left) {                             // drop the marked entries, keep the order of the rest
      __syncwarp();
      uint32_t wpos = 0;
      for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        const int32_t v = (i < n) ? idx[i] : -1;
        const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
        if (v >= 0) idx[wpos + __popc(m & ((1u << lane) - 1u))] = v;   // wpos + rank <= i: never ahead of the reads
        wpos += __popc(m);
      }
      __syncwarp();
      n = wpos;
    }
    pgmax_old = pgmax > 0.0 ? pgmax : INFINITY;
  }
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (lane + 32 * j < F) coef[(size_t)k * F + lane + 32 * j] = w[j];
  if (lane == 0) {
    intercept[k] = wb * bias;
    epochs[k] = it;
  }
}

// X fp64 [V][F] row-major; class_index int32 [V] in [0, n_classes); n_classes == 2 fits ONE problem (positive = class 1,
// like scikit-learn's coef_ of shape [1][F]).  work: fp64 [(P + 1) * V + (P * V + 1) / 2] (Q_ii, dual variables, active
// sets as int32), P = number of problems.  Returns nullptr or an error string.
const char* svm_fit_run(const double* X, const int32_t* class_index, int V, int F, int n_classes, double C, double bias,
                        double tol, int max_iter, double* coef, double* intercept, int32_t* epochs, double* work,
                        cudaStream_t st) {
  const int K = n_classes == 2 ? 1 : n_classes, first = n_classes == 2 ? 1 : 0;
  const bool use_bias = bias > 0.0;
  const double b = use_bias ? bias : 0.0, D = 0.5 / C;
  double* qd = work;
  double* alpha = work + V;
  int32_t* active = reinterpret_cast<int32_t*>(work + (size_t)(K + 1) * V);
  if (cudaMemsetAsync(alpha, 0, (size_t)K * V * sizeof(double), st) != cudaSuccess) return "svm_fit: memset failed";
  count_launch();
  svm_qd_kernel<<<(V + 7) / 8, 256, 0, st>>>(X, V, F, b * b, D, qd);
  count_launch();
#define VA_SVM_LAUNCH(KM) \
  svm_dcd_kernel<KM><<<K, 32, 0, st>>>(X, V, F, class_index, first, qd, D, b, tol, max_iter, alpha, active, coef, intercept, epochs)
  if (F <= 64) VA_SVM_LAUNCH(2);
  else if (F <= 256) VA_SVM_LAUNCH(8);
  else if (F <= 512) VA_SVM_LAUNCH(16);
  else if (F <= 1024) VA_SVM_LAUNCH(32);
  else return "svm_fit: more than 1024 features";
#undef VA_SVM_LAUNCH
  return cudaGetLastError() == cudaSuccess ? nullptr : "svm_fit: launch failed";
}

}  // namespace va
