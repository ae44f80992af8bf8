// DTW lower-bound matrices on the device: the transformers of wildboar.distance.lb
// (DtwKeoghLowerBound LB:314-432, DtwKimLowerBound LB:198-311).  The reference fills
// the (n_query, n_fit) matrix with a PYTHON double loop over `_dtw_lb_keogh` (EL:1095-1115 ->
// cumulative_bound EL:228-260); here one thread owns one fitted sample and QB queries, the sums run in
// the reference's order (k = 0..T-1, one rounding per operation), so the matrices are bit-equal.
//
// Layout: fitted samples and their envelopes TRANSPOSED ([t][sample]) so that a warp whose lanes
// are 32 consecutive samples loads coalesced; the QB queries of a CTA (series + envelopes) are staged
// in shared memory in chunks of KC time steps and read as broadcasts.  Each loaded sample value
// is used by QB queries x 2 directions: ~10 FP64 instructions per (pair, k) against 3 B of L2
// traffic -- FP64-issue bound, like the DP kernels.
#pragma once
#include <cuda_runtime.h>
#include "metrics.cuh"

namespace wb {

constexpr int kLbQB = 8;     // queries per CTA (register accumulators: 2 * QB doubles per thread)
constexpr int kLbKC = 128;   // time steps staged per chunk
constexpr int kLbNT = 256;   // threads per CTA = fitted samples per CTA

struct LbMatArgs {
  const double* q; const double* qlo; const double* qhi;     // (nq, T) row-major: queries + envelopes
  const double* xT; const double* xloT; const double* xhiT;  // (T, nx) transposed: fitted samples + envelopes
  long long nq, nx; int T;
  double* out; long long ld;                                  // out[i * ld + j]
};

// e(v; lo, hi)^2 with the reference's branches (EL:247-253)
__device__ __forceinline__ double lb_excess_sq(double v, double lo, double hi) {
  double d = 0.0;
  if (v > hi) { const double s = v - hi; d = s * s; }
  else if (v < lo) { const double s = v - lo; d = s * s; }
  return d;
}

template <bool LEFT, bool RIGHT>
__global__ void __launch_bounds__(kLbNT) k_lb_keogh_matrix(LbMatArgs a) {
  __shared__ double sq[kLbQB][kLbKC], slo[kLbQB][kLbKC], shi[kLbQB][kLbKC];
  const long long nxb = (a.nx + kLbNT - 1) / kLbNT;
  const long long nqb = (a.nq + kLbQB - 1) / kLbQB;
  // consecutive CTAs share the sample block and differ in the query block: the transposed sample
  // tiles are re-read from L2, not HBM
  for (long long t = blockIdx.x; t < nxb * nqb; t += gridDim.x) {
    const long long xb = t / nqb, qb = t - xb * nqb;
    const long long i0 = qb * kLbQB;
    const int nqi = (int)min((long long)kLbQB, a.nq - i0);
    const long long j = xb * kLbNT + threadIdx.x;
    const bool valid = j < a.nx;
    const long long jc = valid ? j : a.nx - 1;
    double s1[kLbQB], s2[kLbQB];
#pragma unroll
    for (int u = 0; u < kLbQB; ++u) { s1[u] = 0.0; s2[u] = 0.0; }
    for (int k0 = 0; k0 < a.T; k0 += kLbKC) {
      const int kc = min(kLbKC, a.T - k0);
      __syncthreads();
      for (int e = threadIdx.x; e < kLbQB * kLbKC; e += kLbNT) {
        const int u = e / kLbKC, k = e - u * kLbKC;
        const bool in = u < nqi && k < kc;
        const long long src = (i0 + (u < nqi ? u : 0)) * a.T + k0 + (k < kc ? k : 0);
        sq[u][k] = in ? a.q[src] : 0.0;
        slo[u][k] = (in && RIGHT) ? a.qlo[src] : 0.0;
        shi[u][k] = (in && RIGHT) ? a.qhi[src] : 0.0;
      }
      __syncthreads();
      for (int k = 0; k < kc; ++k) {
        const long long off = (long long)(k0 + k) * a.nx + jc;
        const double xv = RIGHT ? a.xT[off] : 0.0;
        const double xl = LEFT ? a.xloT[off] : 0.0, xh = LEFT ? a.xhiT[off] : 0.0;
#pragma unroll
        for (int u = 0; u < kLbQB; ++u) {
          if (LEFT) s1[u] += lb_excess_sq(sq[u][k], xl, xh);          // query against the sample's envelope
          if (RIGHT) s2[u] += lb_excess_sq(xv, slo[u][k], shi[u][k]);  // sample against the query's envelope
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int u = 0; u < kLbQB; ++u) {
        if (u < nqi) {
          const double d1 = LEFT ? sqrt(s1[u]) : -WB_INF, d2 = RIGHT ? sqrt(s2[u]) : -WB_INF;
          a.out[(i0 + u) * a.ld + j] = d1 > d2 ? d1 : d2;   // LB:431 max(dist1, dist2)
        }
      }
    }
  }
}

// dtw_lb_keogh (distance/dtw.py:193-243 -> _dtw_lb_keogh EL:1095-1115 -> cumulative_bound EL:228-260 with mean 0 / std 1
// and no best-so-far): cb[i][k] = squared excess of x[i][k] over [lower[i][k], upper[i][k]], min_dist[i] = sqrt of their
// sum in time order.  One warp per series: the lanes compute the terms (coalesced), lane 0 adds them up sequentially.
__global__ void __launch_bounds__(128) k_lb_keogh_terms(const double* __restrict__ x, const double* __restrict__ lo,
                                                        const double* __restrict__ hi, long long n, int T,
                                                        double* __restrict__ min_dist, double* __restrict__ cb) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const long long base = i * T;
  for (int k = lane; k < T; k += 32) cb[base + k] = lb_excess_sq(x[base + k], lo[base + k], hi[base + k]);
  __syncwarp();
  if (lane == 0) {
    double s = 0.0;
    for (int k = 0; k < T; ++k) s += cb[base + k];
    min_dist[i] = sqrt(s);
  }
}

// DtwKimLowerBound.transform (LB:241-311): sum of squared terms over the first / last three points
// (no sqrt in the reference).  q = queries (rows of the result), x = fitted samples (columns).
__device__ __forceinline__ double kim_d(double a, double b) { const double v = a - b; return v * v; }

__global__ void __launch_bounds__(256) k_lb_kim_matrix(const double* __restrict__ q, long long nq,
                                                       const double* __restrict__ x, long long nx, int T,
                                                       double* __restrict__ out, long long ld) {
  const long long total = nq * nx;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / nx, j = e - i * nx;
    const double* Y = q + i * T;
    const double* X = x + j * T;
    const double x0 = X[0], x0_ = X[T - 1], y0 = Y[0], y0_ = Y[T - 1];
    double d = kim_d(x0, y0) + kim_d(x0_, y0_);
    if (T > 1) {
      const double x1 = X[1], x1_ = X[T - 2], y1 = Y[1], y1_ = Y[T - 2];
      d += dmin2(kim_d(x1, y0), dmin2(kim_d(x0, y1), kim_d(x1, y1)));
      d += dmin2(kim_d(x1_, y1), dmin2(kim_d(x0_, y1_), kim_d(x1_, y1_)));
      if (T > 2) {
        const double x2 = X[2], x2_ = X[T - 3], y2 = Y[2], y2_ = Y[T - 3];
        d += dmin2(kim_d(x0, y2), dmin2(kim_d(x1, y2), dmin2(kim_d(x2, y2), dmin2(kim_d(x2, y1), kim_d(x2, y0)))));
        d += dmin2(kim_d(x0_, y2_), dmin2(kim_d(x1_, y2_), dmin2(kim_d(x2_, y2_), dmin2(kim_d(x2_, y1_), kim_d(x2_, y0_)))));
      }
    }
    out[i * ld + j] = d;
  }
}

}  // namespace wb
