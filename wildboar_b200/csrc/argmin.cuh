// argmin_distance on the device: an EXACT replay of the reference's sequential scan
// (CD:1302-1345) that still runs tens of thousands of pairs in parallel.
//
// Reference semantics for one query i: scan j = 0..ny-1 with a running threshold t (INF until
// the k-heap is full, then the heap maximum); skip j when lower_bound[i,j] >= t; otherwise
// eadistance() early-abandons when a DP row's minimum exceeds T(t) and accepts iff d < t.
//
// Device scheme: the references are processed in CHUNKS of C columns.  For a chunk, every
// (query, ref) distance is computed in parallel against the threshold the query had at the
// START of the chunk (t_cs >= every t inside the chunk, so abandoning against T(t_cs) can
// only remove pairs the reference rejects too).  A replay kernel (one warp per query) then
// walks the chunk's columns in order with the exact rule and the exact heap
// (utils/_misc.pyx:18-107), so indices, distances AND the heap-array order match the reference:
//   * DTW family: rows' minima never decrease, so "abandoned" <=> "d >= t": d alone suffices.
//   * lcss/erp/edr/msm/twe: abandoning is not monotone, so the row-scan engine
//     also returns M = max over checked rows of the row minimum; the pair is rejected iff
//     M > T(t) with the exact t.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <mutex>
#include "dispatch.cuh"
#include "kernels.cuh"

namespace wb {

// Reference-side operands of the LB cascade (outward-rounded fp32 envelopes / values, transposed and permuted): they depend
// only on the reference set and the window, so a device-resident fitted set keeps them between calls (wb_fitted, one per
// device) instead of rebuilding them for every query batch (4.4 ms per call for 200 000 x 256).
struct LbCascCache {
  std::mutex mu;
  bool valid = false;
  int T = 0, w = 0, stride = 0;
  long long ny = 0;
  const double* py = nullptr;
  float2* envT = nullptr; float2* yvT = nullptr; double* y0 = nullptr; double* yL = nullptr;
};

struct ArgminIo {
  LbCascCache* casc_cache = nullptr;  // optional (fitted sets)
  int64_t k;
  const double* lower_bound;  // host, rows of this device's query block, leading dim lb_ld
  int64_t lb_ld;
  int64_t* out_idx;           // host (nq, k)
  double* out_dist;           // host (nq, k)
  int use_device_lb;
};

enum ThrKind : int { TK_SQUARE = 0, TK_IDENT = 1, TK_SCALE = 2, TK_LCSS = 3, TK_NONE = 4 };

// T(t): the early-abandon threshold eadistance() hands to the DP (EL:3205, 3526-3528, 3657, 3838)
__device__ __forceinline__ double ea_threshold(int kind, double t, double scale) {
  switch (kind) {
    case TK_SQUARE: return t * t;
    case TK_IDENT: return t;
    case TK_SCALE: return t * scale;
    case TK_LCSS: return isinf(t) ? WB_INF : scale - t * scale;
    default: return WB_INF;
  }
}

// thresholds handed to the DP kernel for a chunk; LCSS's T(t) is not monotone in t, so no
// device-side abandoning for it (the replay still applies the exact rule).
__global__ void k_thr_raw(const double* __restrict__ tau, long long n, int kind, double scale,
                          double* __restrict__ thr) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  thr[q] = (kind == TK_LCSS || kind == TK_NONE) ? WB_INF : ea_threshold(kind, tau[q], scale);
}

struct HeapEl { long long index; double value; };

__device__ inline void heap_shift_down(long long* hi, double* hv, int startpos, int pos) {
  const long long ni = hi[pos]; const double nv = hv[pos];
  while (pos > startpos) {
    const int parent = (pos - 1) >> 1;
    if (nv > hv[parent]) { hi[pos] = hi[parent]; hv[pos] = hv[parent]; pos = parent; continue; }
    break;
  }
  hi[pos] = ni; hv[pos] = nv;
}
__device__ inline void heap_shift_up(long long* hi, double* hv, int pos, int endpos) {
  const int startpos = pos;
  const long long ni = hi[pos]; const double nv = hv[pos];
  int child = 2 * pos + 1;
  while (child < endpos) {
    const int right = child + 1;
    if (right < endpos && hv[child] < hv[right]) child = right;
    hi[pos] = hi[child]; hv[pos] = hv[child];
    pos = child;
    child = 2 * pos + 1;
  }
  hi[pos] = ni; hv[pos] = nv;
  heap_shift_down(hi, hv, startpos, pos);
}
// utils/_misc.pyx:76-89
__device__ inline void heap_push(long long* hi, double* hv, int& n, int cap, long long index, double value) {
  if (n == 0) { hi[0] = index; hv[0] = value; n = 1; }
  else if (n < cap) { hi[n] = index; hv[n] = value; n++; heap_shift_down(hi, hv, 0, n - 1); }
  else if (hv[0] > value) { hi[0] = index; hv[0] = value; heap_shift_up(hi, hv, 0, cap); }
}

struct ReplayArgs {
  const double* d;   // (nq, ld) distances of this chunk (INF = abandoned / pruned)
  const double* m;   // (nq, ld) row-minimum maxima or nullptr (DTW family)
  const double* lb;  // (nq, ld) user lower bounds of this chunk or nullptr
  long long ld, nq, c0, ncols;
  int k, kind;
  double scale;
  double* tau; long long* hidx; double* hval; int* hn;
};

__global__ void __launch_bounds__(128) k_replay(ReplayArgs a) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = wid; q < a.nq; q += nw) {
    double t = a.tau[q];
    int n = a.hn[q];
    long long* hi = a.hidx + q * a.k;
    double* hv = a.hval + q * a.k;
    const double* drow = a.d + q * a.ld;
    for (long long jj = 0; jj < a.ncols; jj += 32) {
      const long long j = jj + lane;
      const double dv = (j < a.ncols) ? drow[j] : WB_INF;
      unsigned mask = __ballot_sync(0xffffffffu, dv < t);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const double ds = __shfl_sync(0xffffffffu, dv, src);
        const long long js = jj + src;
        bool acc = ds < t;
        if (acc && a.lb) acc = !(a.lb[q * a.ld + js] >= t);
        if (acc && a.m) acc = !(a.m[q * a.ld + js] > ea_threshold(a.kind, t, a.scale));
        if (acc) {
          if (lane == 0) {
            heap_push(hi, hv, n, a.k, a.c0 + js, ds);
            t = (n == a.k) ? hv[0] : WB_INF;
          }
          t = __shfl_sync(0xffffffffu, t, 0);
          n = __shfl_sync(0xffffffffu, n, 0);
        }
      }
    }
    if (lane == 0) { a.tau[q] = t; a.hn[q] = n; }
    __syncwarp();
  }
}

__global__ void k_fill(double* p, long long n, double v) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) p[e] = v;
}

// ------------------------------------------------------------------------------------------
// On-device lower-bound cascade for dtw: LB_Kim (first/last point) ->
// LB_Keogh(query, envelope(ref)) -> LB_Keogh(ref, envelope(query)) -> survivors go to the
// early-abandoning DP.  Envelope half-width = R-1 (the band is |i-j| <= R-1), the tightest
// valid one.  Pruning compares against the chunk-start threshold with a 1e-9 relative safety
// margin, so rounding in the O(T) sums can never drop a pair the exact scan would accept; it
// therefore NEVER changes the result (the replay only ever sees "pruned" = +INF = rejected).
// ------------------------------------------------------------------------------------------
// Time order of the lower-bound sums.  Summing the time steps in natural order makes the early exit wait for
// the part of the series where the excess over the envelope happens to be (for random walks: the end).  The
// cascade therefore visits time step perm[p] = (p * stride) mod T at position p, with stride ~ 0.618 T coprime
// to T (a golden-ratio sequence: every prefix is spread evenly over the series), and all arrays it reads are
// stored in that order.  Only the pruning decision depends on these sums (1e-9 safety margin), never a result.
inline int lb_time_stride(int T) {
  if (T <= 2) return 1;
  int s = (int)(0.6180339887498949 * T);
  if (s < 1) s = 1;
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  while (s > 1 && gcd(s, T) != 1) --s;
  return s;
}

// lower/upper[p] = min/max t[k-w .. k+w] (clipped) at k = perm[p]; rows = series (queries); xp = the series in
// the same order
__global__ void k_envelope_rows(const double* __restrict__ x, long long n, int T, int w, int stride,
                                double* __restrict__ xp, double* __restrict__ lo, double* __restrict__ hi) {
  const long long total = n * (long long)T;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long s = e / T;
    const int pos = (int)(e - s * T);
    const int k = (int)(((long long)pos * stride) % T);
    const double* p = x + s * T;
    const int a = max(0, k - w), b = min(T - 1, k + w);
    double l = p[a], h = p[a];
    for (int q = a + 1; q <= b; ++q) { const double v = p[q]; l = fmin(l, v); h = fmax(h, v); }
    if (xp) xp[e] = p[k];
    lo[e] = l; hi[e] = h;
  }
}
// same, but written TRANSPOSED ([position][series]) together with the transposed series, so that a
// warp whose lanes are 32 consecutive references reads them coalesced
__global__ void k_envelope_T(const double* __restrict__ y, long long n, int T, int w, int stride, double* __restrict__ yT,
                             double* __restrict__ loT, double* __restrict__ hiT) {
  const long long total = n * (long long)T;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)(e / n);
    const long long s = e - (long long)pos * n;
    const int k = (int)(((long long)pos * stride) % T);
    const double* p = y + s * T;
    const int a = max(0, k - w), b = min(T - 1, k + w);
    double l = p[a], h = p[a];
    for (int q = a + 1; q <= b; ++q) { const double v = p[q]; l = fmin(l, v); h = fmax(h, v); }
    yT[e] = p[k]; loT[e] = l; hiT[e] = h;
  }
}

// Cascade operands, all fp32 and rounded OUTWARD so that every LB_Keogh term computed from them (with round-down
// arithmetic) is a rigorous lower bound of the exact fp64 term for ANY data: a larger envelope and a value interval
// [v_down, v_up] can only shrink the excess over the envelope.  fp32 because the prune kernel is instruction bound:
// 2 FADD + FMNMX + FFMA per direction and time step instead of 7 FP64-pipe instructions + 4 selects.
//   references: envT[pos][j] = (lower_down, upper_up), yvT[pos][j] = (y_down, y_up); y0 / yL: first / last sample (fp64, LB_Kim)
//   queries   : qf[i][pos]   = (x_down, x_up, lower_down, upper_up)
// pos = position in the permuted time order (time step (pos * stride) mod T).
__global__ void k_envelope_casc(const double* __restrict__ y, long long n, int T, int w, int stride, float2* __restrict__ envT,
                                float2* __restrict__ yvT, double* __restrict__ y0, double* __restrict__ yL) {
  const long long total = n * (long long)T;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)(e / n);
    const long long s = e - (long long)pos * n;
    const int k = (int)(((long long)pos * stride) % T);
    const double* p = y + s * T;
    const int a = max(0, k - w), b = min(T - 1, k + w);
    double l = p[a], h = p[a];
    for (int q = a + 1; q <= b; ++q) { const double v = p[q]; l = fmin(l, v); h = fmax(h, v); }
    envT[e] = make_float2(__double2float_rd(l), __double2float_ru(h));
    yvT[e] = make_float2(__double2float_rd(p[k]), __double2float_ru(p[k]));
    if (pos == 0) { y0[s] = p[0]; yL[s] = p[T - 1]; }
  }
}
__global__ void k_query_casc(const double* __restrict__ x, long long n, int T, int w, int stride, float4* __restrict__ qf) {
  const long long total = n * (long long)T;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long s = e / T;
    const int pos = (int)(e - s * T);
    const int k = (int)(((long long)pos * stride) % T);
    const double* p = x + s * T;
    const int a = max(0, k - w), b = min(T - 1, k + w);
    double l = p[a], h = p[a];
    for (int q = a + 1; q <= b; ++q) { const double v = p[q]; l = fmin(l, v); h = fmax(h, v); }
    qf[e] = make_float4(__double2float_rd(p[k]), __double2float_ru(p[k]), __double2float_rd(l), __double2float_ru(h));
  }
}

#ifndef WB_LB_STRAGGLERS
#define WB_LB_STRAGGLERS 2
#endif
#ifndef WB_LB_STRAGGLER_AFTER
#define WB_LB_STRAGGLER_AFTER 63
#endif
constexpr int kLbStragglers = WB_LB_STRAGGLERS, kLbStragglerAfter = WB_LB_STRAGGLER_AFTER;

struct LbArgs {
  const double* x;     // (nq, T) queries, natural order (LB_Kim reads the first / last sample)
  const float4* qf;    // (nq, T) (x_down, x_up, lower_down, upper_up), permuted time order
  const float2* envT;  // (T, ny) (lower_down, upper_up) of the references, rows in the same order
  const float2* yvT;   // (T, ny) (y_down, y_up)
  const double* y0; const double* yL;  // (ny) first / last sample of the references
  long long nq, ny, c0, nc; int T;
  const double* tau;  // per query, distance domain
  double* d; long long ld;  // chunk matrix: +INF = pruned, -1 = survivor (to be filled by the DP)
  unsigned long long* n_kim; unsigned long long* n_keogh;  // pruning statistics
};

__global__ void __launch_bounds__(256) k_lb_prune(LbArgs a) {
  const int lane = threadIdx.x & 31;
  const long long nyb = (a.nc + 31) / 32;
  const long long ntask = a.nq * nyb;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int T = a.T;
  unsigned long long c_kim = 0, c_keogh = 0;  // per-warp statistics (lane 0), one atomic per warp at the end
  for (long long t = wid; t < ntask; t += nw) {
    const long long i = t / nyb;
    const long long jl = (t - i * nyb) * 32 + lane;
    const bool valid = jl < a.nc;
    const long long j = a.c0 + (valid ? jl : a.nc - 1);
    const double tau = a.tau[i];
    const double lim = isinf(tau) ? WB_INF : tau * tau * (1.0 + 1e-9);  // prune only if LB^2 > lim
    const double* q = a.x + i * T;
    bool pruned = false;
    if (!isinf(lim)) {
      // LB_Kim: every warping path contains (0,0) and (T-1,T-1)
      const double d0 = q[0] - a.y0[j];
      const double d1 = q[T - 1] - a.yL[j];
      double lb = d0 * d0 + (T > 1 ? d1 * d1 : 0.0);
      pruned = lb > lim;
      c_kim += __popc(__ballot_sync(0xffffffffu, pruned && valid));
      bool p2 = pruned;
      if (!__all_sync(0xffffffffu, pruned)) {
        // LB_Keogh in BOTH directions at once (time steps in the permuted order): s1 = query against the reference's
        // envelope, s2 = reference against the query's envelope, in round-down fp32 on outward-rounded operands (rigorous
        // lower bounds of the fp64 sums).  A lane is pruned as soon as EITHER sum exceeds the limit, so the warp leaves
        // at max over lanes of min(k1, k2) instead of running direction 1 to the end for every lane that only direction
        // 2 can prune (ncu: 4800 instructions per warp task before, 80 % of the warps ran all T steps of direction 1).
        const float limf = __double2float_ru(lim);
        float s1 = 0.0f, s2 = 0.0f;
        const float2* env = a.envT + j;
        const float2* yv = a.yvT + j;
        const float4* qf = a.qf + i * T;
        const long long ny = a.ny;
        // one time step: excess of the query sample over the reference envelope (>= 0 parts of x - upper and lower - x)
        // and of the reference sample over the query envelope
        auto step = [&](const float2 lh, const float2 yy, const float4 qq) {
          const float e1 = fmaxf(fmaxf(__fsub_rd(qq.x, lh.y), __fsub_rd(lh.x, qq.y)), 0.0f);
          s1 = __fmaf_rd(e1, e1, s1);
          const float e2 = fmaxf(fmaxf(__fsub_rd(yy.x, qq.w), __fsub_rd(qq.z, yy.y)), 0.0f);
          s2 = __fmaf_rd(e2, e2, s2);
        };
        int k = 0;
        // blocks of 8 steps (pointer increments, all 24 loads of a block issued up front), then one warp vote: leave
        // when every lane is pruned -- or when only a straggler or two are left after a good part of the series:
        // finishing their sums would keep the whole warp busy, handing them to the (early-abandoning) DP is cheaper.
        // Pruning is optional, so this cannot change a result.
        for (; k + 8 <= T; k += 8) {
          float2 lh[8], yy[8]; float4 qq[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { lh[u] = env[u * ny]; yy[u] = yv[u * ny]; qq[u] = qf[k + u]; }
#pragma unroll
          for (int u = 0; u < 8; ++u) step(lh[u], yy[u], qq[u]);
          env += 8 * ny; yv += 8 * ny;
          const int alive = __popc(__ballot_sync(0xffffffffu, !(pruned || s1 > limf || s2 > limf)));
          if (alive == 0 || (alive <= kLbStragglers && k + 7 >= kLbStragglerAfter)) { k = T; break; }
        }
        for (; k < T; ++k) { step(*env, *yv, qf[k]); env += ny; yv += ny; }
        p2 = pruned || s1 > limf || s2 > limf;
        c_keogh += __popc(__ballot_sync(0xffffffffu, p2 && !pruned && valid));
        pruned = p2;
      }
    }
    if (valid) a.d[i * a.ld + jl] = pruned ? WB_INF : -1.0;
  }
  if (lane == 0) {
    if (c_kim) atomicAdd(a.n_kim, c_kim);
    if (c_keogh) atomicAdd(a.n_keogh, c_keogh);
  }
}

// survivors per query row -> exclusive scan -> (i, j) list sorted by (i, j)
__global__ void __launch_bounds__(128) k_row_count(const double* __restrict__ d, long long nq, long long nc, long long ld,
                                                   int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = wid; q < nq; q += nw) {
    int c = 0;
    for (long long jj = 0; jj < nc; jj += 32) {
      const long long j = jj + lane;
      c += __popc(__ballot_sync(0xffffffffu, j < nc && d[q * ld + j] < 0.0));
    }
    if (lane == 0) counts[q] = c;
  }
}
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ counts, long long nq, int* __restrict__ starts,
                                                      int* __restrict__ total, unsigned long long* __restrict__ grand_total) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const long long per = (nq + 1023) / 1024;
  const long long lo = tid * per, hi = min(nq, lo + per);
  int s = 0;
  for (long long q = lo; q < hi; ++q) s += counts[q];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int k = 0; k < 1024; ++k) { const int v = part[k]; part[k] = acc; acc += v; }
    *total = acc;
    atomicAdd(grand_total, (unsigned long long)acc);
  }
  __syncthreads();
  int acc = part[tid];
  for (long long q = lo; q < hi; ++q) { starts[q] = acc; acc += counts[q]; }
}
__global__ void __launch_bounds__(128) k_fill_list(const double* __restrict__ d, long long nq, long long nc, long long ld,
                                                   const int* __restrict__ starts, int2* __restrict__ list) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = wid; q < nq; q += nw) {
    int base = starts[q];
    for (long long jj = 0; jj < nc; jj += 32) {
      const long long j = jj + lane;
      const bool sv = j < nc && d[q * ld + j] < 0.0;
      const unsigned m = __ballot_sync(0xffffffffu, sv);
      if (sv) list[base + __popc(m & ((1u << lane) - 1u))] = make_int2((int)q, (int)j);
      base += __popc(m);
    }
  }
}

// LaunchFn(r0, nrows, c0, ncols, out, ld, out_m, thr, stats) -> int
template <class WS, class DI, class Call, class LaunchFn>
int run_argmin(WS& ws, const DI& di, Call& c, const ArgminIo& io, wb_stats* stats, LaunchFn launch) {
  (void)di;
  cudaStream_t st = ws.stream;
  const long long nq = c.nx, ny = c.ny;
  const int k = (int)io.k;
  int kind; double scale = 1.0;
  const bool dtwfam = is_dtw_family(c.metric);
  // adtw with a negative penalty (AmercingDtwMetric accepts any float, EL:3349): the row minima are no longer monotone,
  // so a pair whose final distance is below the threshold can still be abandoned by the reference's eadistance (a row
  // minimum > t * t, EL:3390-3396).  It is then replayed like the non-DTW metrics: row-minimum maxima M from the
  // row-ordered engines, rejected iff M > t * t.
  const bool adtw_neg = c.metric == M_ADTW && c.p.p < 0;
  if (dtwfam) kind = TK_SQUARE;
  else if (c.metric == M_LCSS || c.metric == M_WLCSS) { kind = TK_LCSS; scale = (double)std::min(c.Tx, c.Ty); }
  else if (c.metric == M_EDR) { kind = TK_SCALE; scale = (double)std::max(c.Tx, c.Ty); }
  else kind = TK_IDENT;

  double *tau = nullptr, *thr = nullptr, *hval = nullptr, *dbuf = nullptr, *mbuf = nullptr, *lbuf = nullptr;
  long long* hidx = nullptr; int* hn = nullptr;
  // Columns per chunk.  The thresholds a chunk is pruned with are those at its start, so the FIRST chunks should be small
  // (they run dense or nearly so) and the later ones large (few launches: with one query every chunk is ~6 launches of
  // almost no work).  The chunk grows 4x per step from 1024 columns up to C = min(32768, 4 Mi / nq) -- nq * C values per
  // buffer; for the cfg4 share (2500 queries) that is 1024 then 1600, as before; for one query 1024, 4096, 16384, 32768, ...
  long long C = (4LL << 20) / std::max<long long>(nq, 1);
  C = std::max<long long>(32, std::min<long long>(32768, (C / 32) * 32));
  long long c_first = std::min<long long>(C, 1024);
  if (const char* e = getenv("WILDBOAR_CUDA_ARGMIN_CHUNK")) {  // tuning / test knob: fixed columns per chunk
    const long long v = atoll(e);
    if (v >= 32) { C = (v / 32) * 32; c_first = C; }
  }
  C = std::min<long long>(C, ((ny + 31) / 32) * 32);
  c_first = std::min(c_first, C);
  if (ws.alloc(&tau, (size_t)nq) || ws.alloc(&thr, (size_t)nq) || ws.alloc(&hval, (size_t)nq * k) ||
      ws.alloc(&hidx, (size_t)nq * k) || ws.alloc(&hn, (size_t)nq) || ws.alloc(&dbuf, (size_t)nq * C)) return 1;
  if ((!dtwfam || adtw_neg) && ws.alloc(&mbuf, (size_t)nq * C)) return 1;
  if (io.lower_bound && ws.alloc(&lbuf, (size_t)nq * C)) return 1;
  // ---- optional on-device lower-bound cascade (dtw, equal lengths) ----
  const bool cascade = io.use_device_lb && c.metric == M_DTW && c.ptx == c.pty && c.ptx >= 2 && !c.degenerate &&
                       nq * C < 2000000000LL;
  float4* qf = nullptr; float2 *envT = nullptr, *yvT = nullptr;
  double *y0 = nullptr, *yL = nullptr;
  int *counts = nullptr, *starts = nullptr, *list_len = nullptr; int2* list = nullptr;
  unsigned long long* lbstat = nullptr;  // [0] kim-pruned, [1] keogh-pruned, [2] survivors
  if (cascade) {
    const int T = c.ptx, w = std::max(c.R - 1, 0);
    if (ws.alloc(&qf, (size_t)nq * T) || ws.alloc(&counts, (size_t)nq) ||
        ws.alloc(&starts, (size_t)nq) || ws.alloc(&list_len, 1) || ws.alloc(&list, (size_t)nq * C) ||
        ws.alloc(&lbstat, 3)) return 1;
    if (cudaMemsetAsync(lbstat, 0, 3 * sizeof(unsigned long long), st) != cudaSuccess) return 1;
    const int stride = lb_time_stride(T);
    k_query_casc<<<1024, 256, 0, st>>>(c.px, nq, T, w, stride, qf);
    LbCascCache* cc = io.casc_cache;
    bool cached = false;
    if (cc) {
      // built ONCE per fitted set, by the first call, under the lock and finished before anyone else may read it; a later
      // call with another window / length leaves it alone (somebody may be reading it) and builds its own operands
      std::lock_guard<std::mutex> lk(cc->mu);
      if (!cc->valid && !cc->envT) {
        if (cudaMallocAsync((void**)&cc->envT, sizeof(float2) * (size_t)ny * T, st) != cudaSuccess ||
            cudaMallocAsync((void**)&cc->yvT, sizeof(float2) * (size_t)ny * T, st) != cudaSuccess ||
            cudaMallocAsync((void**)&cc->y0, sizeof(double) * (size_t)ny, st) != cudaSuccess ||
            cudaMallocAsync((void**)&cc->yL, sizeof(double) * (size_t)ny, st) != cudaSuccess) { cudaGetLastError(); return 1; }
        k_envelope_casc<<<2048, 256, 0, st>>>(c.py, ny, T, w, stride, cc->envT, cc->yvT, cc->y0, cc->yL);
        if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
        cc->T = T; cc->w = w; cc->stride = stride; cc->ny = ny; cc->py = c.py; cc->valid = true;
      }
      if (cc->valid && cc->T == T && cc->w == w && cc->stride == stride && cc->ny == ny && cc->py == c.py) {
        envT = cc->envT; yvT = cc->yvT; y0 = cc->y0; yL = cc->yL;
        cached = true;
      }
    }
    if (!cached) {
      if (ws.alloc(&envT, (size_t)ny * T) || ws.alloc(&yvT, (size_t)ny * T) || ws.alloc(&y0, (size_t)ny) || ws.alloc(&yL, (size_t)ny)) return 1;
      k_envelope_casc<<<2048, 256, 0, st>>>(c.py, ny, T, w, stride, envT, yvT, y0, yL);
    }
  }
  k_fill<<<256, 256, 0, st>>>(tau, nq, WB_INF);
  if (cudaMemsetAsync(hval, 0, sizeof(double) * nq * k, st) != cudaSuccess ||
      cudaMemsetAsync(hidx, 0, sizeof(long long) * nq * k, st) != cudaSuccess ||
      cudaMemsetAsync(hn, 0, sizeof(int) * nq, st) != cudaSuccess) return 1;

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  int rc = 0;
  long long c_cur = c_first;
  for (long long c0 = 0; c0 < ny && !rc; ) {
    const long long nc = std::min(c_cur, ny - c0);
    const long long c0_next = c0 + nc;
    c_cur = std::min(C, c_cur * 4);
    k_thr_raw<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(tau, nq, kind, scale, thr);
    if (c.degenerate) {
      // ddtw with T < 3: eadistance() returns False for every pair (EL:3297-3298)
      k_fill<<<256, 256, 0, st>>>(dbuf, nq * C, WB_INF);
    } else if (cascade && c0 > 0) {
      // chunk 0 has no threshold yet (tau = INF): nothing can be pruned, run it densely
      LbArgs la;
      la.x = c.px; la.qf = qf; la.envT = envT; la.yvT = yvT; la.y0 = y0; la.yL = yL;
      la.nq = nq; la.ny = ny; la.c0 = c0; la.nc = nc; la.T = c.ptx; la.tau = tau; la.d = dbuf; la.ld = C;
      la.n_kim = lbstat; la.n_keogh = lbstat + 1;
      k_lb_prune<<<148 * 8, 256, 0, st>>>(la);
      k_row_count<<<148 * 4, 128, 0, st>>>(dbuf, nq, nc, C, counts);
      k_scan_counts<<<1, 1024, 0, st>>>(counts, nq, starts, list_len, lbstat + 2);
      k_fill_list<<<148 * 4, 128, 0, st>>>(dbuf, nq, nc, C, starts, list);
      c.mode = PM_LIST; c.list = list; c.list_len = list_len;
      rc = launch(0, nq, c0, nc, dbuf, C, nullptr, thr, nullptr);
      c.mode = PM_PAIRWISE; c.list = nullptr; c.list_len = nullptr;
      if (stats) stats->launches += 5;
      if (rc) break;
    } else {
      rc = launch(0, nq, c0, nc, dbuf, C, mbuf, (kind == TK_NONE || kind == TK_LCSS) ? nullptr : thr, stats);
      if (rc) break;
    }
    if (lbuf) {
      if (cudaMemcpy2DAsync(lbuf, sizeof(double) * C, io.lower_bound + c0, sizeof(double) * io.lb_ld,
                            sizeof(double) * nc, nq, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = 1; break; }
    }
    ReplayArgs ra;
    ra.d = dbuf; ra.m = mbuf; ra.lb = lbuf; ra.ld = C; ra.nq = nq; ra.c0 = c0; ra.ncols = nc;
    ra.k = k; ra.kind = kind; ra.scale = scale; ra.tau = tau; ra.hidx = hidx; ra.hval = hval; ra.hn = hn;
    const long long blocks = std::max<long long>(1, std::min<long long>((nq + 3) / 4, 148 * 16));
    k_replay<<<(unsigned)blocks, 128, 0, st>>>(ra);
    if (stats) stats->launches += 2;
    c0 = c0_next;
  }
  cudaEventRecord(e1, st);
  if (!rc) {
    if (cudaMemcpyAsync(io.out_idx, hidx, sizeof(long long) * nq * k, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(io.out_dist, hval, sizeof(double) * nq * k, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
    float f = 0;
    if (!rc && stats && cudaEventElapsedTime(&f, e0, e1) == cudaSuccess) stats->kernel_ms += f;
    if (!rc && stats && cascade) {
      unsigned long long h[3] = {0, 0, 0};
      if (cudaMemcpy(h, lbstat, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess) {
        // pairs / cells = what actually went through the DP (dense first chunk + survivors)
        const long long cpp = stats->pairs > 0 ? stats->cells / stats->pairs : 0;
        stats->pairs += (long long)h[2];
        stats->cells += (long long)h[2] * cpp;
        stats->lb_kim_pruned = (long long)h[0];
        stats->lb_keogh_pruned = (long long)h[1];
      }
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

}  // namespace wb
