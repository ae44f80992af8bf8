// __global__ kernels around the DP engines + small prologue kernels.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include "engine_rowscan.cuh"
#include "engine_strip.cuh"
#include "engine_band.cuh"
#include "engine_coop.cuh"

namespace wb {

enum PairMode : int {
  PM_PAIRWISE = 0,  // out[i][j] = d(x_i, y_j)                       (CD:1144-1205)
  PM_SELF = 1,      // j > i only, also written to out[j][i]          (CD:1208-1267)
  PM_PAIRED = 2,    // out[i] = d(x_i, y_i)  (caller already swapped)  (CD:1597-1652)
  PM_LIST = 3,      // out[i][j] for the (i, j) of a device-resident survivor list (argmin cascade)
  PM_LISTP = 4,     // out[e] = d(x_i, y_j) for list entry e = (i, j)   (DBA: members against their centre)
};

// F = arithmetic type of the DP (double: bit-exact mode, float: fp32 mode).  Results, per-series
// scalars and thresholds cross the kernel boundary as doubles in both modes.
template <class F>
struct KArgsT {
  const F* x;  // (nx, Tx) dense, first operand (rows of the DP)
  const F* y;  // (ny, Ty) dense, second operand (columns of the DP)
  long long nx, ny;
  int Tx, Ty;
  Geom g;
  int NS;            // strip engine: boundary-buffer slots per pair (strip_ring_slots(g, W))
  const double* sx;  // per-sample scalars of x (erp gap sums / edr std) or nullptr
  const double* sy;
  const double* sy2; // second per-sample scalar of y (scaled dtw: window std) or nullptr
  double* out;
  long long ld;       // out[i * ld + j]
  double* out_m;      // optional: max over checked rows of the row minimum (row-scan engine)
  const double* thr;  // optional per-x-row early-abandon threshold, RAW dp domain
  long long thr_ld;   // thr_div > 0: the threshold of pair (i, j) is thr[i * thr_ld + j / thr_div] --
  long long thr_div;  //   one per (x row, group of thr_div consecutive y series), e.g. per (subsequence, sample) in a scan
  unsigned long long* counter;  // persistent-grid work counter (zeroed before launch)
  long long ntasks;   // warp tasks
  long long nyb;      // ceil(ny / lt)
  int lt;             // lanes of a warp task that carry a pair (32; fewer when there are not enough pairs to give every
                      //   resident warp slot a full task: more, emptier warps hide latency that fewer, fuller ones cannot)
  int mode;
  long long row0;     // PM_SELF: global index of local x row 0 (row-sharded self join)
  int mirror;         // PM_SELF: also write element (j, i) of the full n x n matrix that `out` is a row block of (`out` = row
                      //   `row0` of it, ld = n): out[(j - row0) * ld + (i + row0)]  (single-device self join)
  const int2* list;   // PM_LIST: pairs to evaluate, in any order (the cascade appends them as its warps finish)
  const int* list_len;  // PM_LIST: number of pairs (device resident: no host round trip)
  F* gring;           // strip engine, GRING variant: boundary buffers in global memory, [warp][slot][lane]
  F* scratch;         // row-scan engine: 2 rows per thread, interleaved
  long long sstride;  // = total threads
  int srows;          // elements per scratch row
  // multivariate dim="mean" (DI:1289-1297, np.mean over the per-dimension matrices): the launches of
  // dimensions 1.. ADD to the stored value (dimension order = numpy's reduction order along axis 0) and
  // the last one divides by the number of dimensions
  int acc;            // 1: out = out + d
  double div;         // != 0: out = (...) / div
  long long ys;       // elements between consecutive y series (Ty for dense rows; 1 = every sliding window of one long buffer)
  long long npairs;   // cooperative engine: pairs of this launch (nx * ny, nx for PM_PAIRED, list length for PM_LISTP)
  int yil;            // 1: y is interleaved in groups of 32 series (k_interleave32): element t of series j sits at
                      //    ((j >> 5) * Ty + t) * 32 + (j & 31), so the 32 lanes of a warp task read consecutive addresses
                      //    (row-scan and band kernels; dense rows of many short series are uncoalesced otherwise)
};
using KArgs = KArgsT<double>;

// One warp task = 32 consecutive pairs.  Returns false when the whole task is empty.
template <class A>
__device__ __forceinline__ long long task_count(const A& a) {
  return (a.mode == PM_LIST || a.mode == PM_LISTP) ? ((long long)__ldg(a.list_len) + 31) / 32 : a.ntasks;
}

template <class A>
__device__ __forceinline__ bool decode_task(const A& a, long long t, int lane, long long& i, long long& j,
                                            bool& valid) {
  if (a.mode == PM_LIST || a.mode == PM_LISTP) {
    const long long n = __ldg(a.list_len);
    long long e = t * 32 + lane;
    valid = e < n;
    if (!valid) e = n - 1;
    const int2 p = a.list[e];
    i = p.x; j = p.y;
    return true;
  }
  const int lt = a.lt;
  const int sub = lane < lt ? lane : lane - lt * (lane / lt);  // surplus lanes repeat a pair of the task (their results are dropped)
  if (a.mode == PM_PAIRED) {
    i = t * lt + sub;
    valid = lane < lt && i < a.nx;
    if (i >= a.nx) i = a.nx - 1;
    j = i;
    return true;
  }
  // x row fastest: the ~2000 warps that are resident at any time then share ONE block of 32 y rows
  // (L1/L2 resident, 128 KB) and differ in the x row (one broadcast load per DP row), instead of
  // touching every y row at once (41 MB for cfg3, which the boundary buffers evict from L2)
  const long long jb = t / a.nx;
  i = t - jb * a.nx;
  j = jb * lt + sub;
  valid = lane < lt && j < a.ny;
  if (j >= a.ny) j = a.ny - 1;
  if (a.mode == PM_SELF) {
    const long long ig = i + a.row0;
    if (jb * lt + (lt - 1) <= ig) return false;
    if (j <= ig) valid = false;
  }
  return true;
}

// dim="mean": sum the per-dimension distances in dimension order, divide after the last one
template <class A>
__device__ __forceinline__ double combine_dims(const A& a, const double* po, double d) {
  if (a.acc) d = *po + d;
  if (a.div != 0.0) d = d / a.div;
  return d;
}

template <class A>
__device__ __forceinline__ double* result_ptr(const A& a, long long t, int lane, long long i, long long j) {
  if (a.mode == PM_PAIRED) return &a.out[i];
  if (a.mode == PM_LISTP) return &a.out[t * 32 + lane];
  return &a.out[i * a.ld + j];
}

__device__ __forceinline__ long long next_task(unsigned long long* counter, int lane) {
  unsigned long long t = 0;
  if (lane == 0) t = atomicAdd(counter, 1ULL);
  return (long long)__shfl_sync(0xffffffffu, t, 0);
}

// ---- strip engine: thread per pair, boundary ring in shared memory [warp][slot][lane] ----
template <class M, int W, int NT, int MINB, bool EA, int NR = 2, bool GRING = false>
__global__ void __launch_bounds__(NT, MINB) k_strip(KArgsT<typename M::real> a, M m) {
  using F = typename M::real;
  extern __shared__ double smem_raw[];
  F* smem = reinterpret_cast<F*>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // boundary buffers of this warp's 32 pairs: shared memory, or (tall bands that would leave too
  // few warps resident) L2-resident global memory
  F* bnd = GRING ? a.gring + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * a.NS * 32 + lane
                 : smem + (size_t)warp * a.NS * 32 + lane;
  const long long ntasks = task_count(a);
  for (;;) {
    const long long t = next_task(a.counter, lane);
    if (t >= ntasks) break;
    long long i, j;
    bool valid;
    if (!decode_task(a, t, lane, i, j, valid)) continue;
    M mm = m;
    PairCtx pc;
    pc.sx = a.sx ? a.sx[i] : 0.0;
    pc.sy = a.sy ? a.sy[j] : 0.0;
    pc.sy2 = a.sy2 ? a.sy2[j] : 0.0;
    mm.begin_pair(pc);
    const F ab = EA ? (F)a.thr[a.thr_div > 0 ? i * a.thr_ld + j / a.thr_div : i] : Num<F>::inf();
    const double d = (double)strip_pair<M, W, EA, NR, 32>(a.g, mm, a.x + i * a.Tx, a.y + j * a.ys, bnd, 32, ab);
    if (valid) {
      // results are written once and never re-read by the kernel: streaming stores keep them from
      // displacing the boundary buffers / y tiles in L2
      double* const po = result_ptr(a, t, lane, i, j);
      const double r = combine_dims(a, po, d);
      __stcs(po, r);
      if (a.mode == PM_SELF && a.mirror) __stcs(&a.out[(j - a.row0) * a.ld + (i + a.row0)], r);
    }
    __syncwarp();
  }
}

// ---- row-scan engine: thread per pair, two scratch rows per thread in global memory ----
// YS = 32: y interleaved in groups of 32 series (KArgs::yil); a separate instantiation so that the plain kernel keeps its
// register budget
template <class M, int NT, int YS = 1>
__global__ void __launch_bounds__(NT) k_rowscan(KArgsT<typename M::real> a, M m) {
  using F = typename M::real;
  const int lane = threadIdx.x & 31;
  const long long gtid = (long long)blockIdx.x * NT + threadIdx.x;
  F* b0 = a.scratch + gtid;
  F* b1 = a.scratch + (long long)a.srows * a.sstride + gtid;
  const long long ntasks = task_count(a);
  for (;;) {
    const long long t = next_task(a.counter, lane);
    if (t >= ntasks) break;
    long long i, j;
    bool valid;
    if (!decode_task(a, t, lane, i, j, valid)) continue;
    M mm = m;
    PairCtx pc;
    pc.sx = a.sx ? a.sx[i] : 0.0;
    pc.sy = a.sy ? a.sy[j] : 0.0;
    pc.sy2 = a.sy2 ? a.sy2[j] : 0.0;
    mm.begin_pair(pc);
    const F md = a.thr ? (F)a.thr[a.thr_div > 0 ? i * a.thr_ld + j / a.thr_div : i] : Num<F>::inf();
    F mmax = F(0);
    const F* const yp = (YS == 32) ? a.y + interleave32_base(j, a.Ty) : a.y + j * a.ys;
    const double d = (double)rowscan_pair<M, YS>(a.g, mm, a.x + i * a.Tx, yp, b0, b1, a.sstride, md, &mmax);
    if (valid) {
      double* const po = result_ptr(a, t, lane, i, j);
      const double r = combine_dims(a, po, d);
      *po = r;
      if (a.mode != PM_PAIRED && a.mode != PM_LISTP) {
        if (a.out_m) a.out_m[i * a.ld + j] = (double)mmax;
        if (a.mode == PM_SELF && a.mirror) a.out[(j - a.row0) * a.ld + (i + a.row0)] = r;
      } else if (a.mode == PM_LISTP && a.out_m) {
        a.out_m[t * 32 + lane] = (double)mmax;  // one entry per list element, like the distances
      }
    }
    __syncwarp();
  }
}

// ---- band-register engine: thread per pair, the previous band row in HB registers (equal lengths, H <= HB) ----
template <class M, int HB, int NT, int YS = 1, bool BLK = false>
__global__ void __launch_bounds__(NT) k_band(KArgsT<typename M::real> a, M m) {
  using F = typename M::real;
  const int lane = threadIdx.x & 31;
  const long long ntasks = task_count(a);
  for (;;) {
    const long long t = next_task(a.counter, lane);
    if (t >= ntasks) break;
    long long i, j;
    bool valid;
    if (!decode_task(a, t, lane, i, j, valid)) continue;
    M mm = m;
    PairCtx pc;
    pc.sx = a.sx ? a.sx[i] : 0.0;
    pc.sy = a.sy ? a.sy[j] : 0.0;
    pc.sy2 = a.sy2 ? a.sy2[j] : 0.0;
    mm.begin_pair(pc);
    const F md = a.thr ? (F)a.thr[a.thr_div > 0 ? i * a.thr_ld + j / a.thr_div : i] : Num<F>::inf();
    F mmax = F(0);
    const F* const yp = (YS == 32) ? a.y + interleave32_base(j, a.Ty) : a.y + j * a.ys;
    const double d = (double)band_pair<M, HB, YS, BLK>(a.g, mm, a.x + i * a.Tx, yp, md, &mmax);
    if (valid) {
      double* const po = result_ptr(a, t, lane, i, j);
      const double r = combine_dims(a, po, d);
      *po = r;
      if (a.mode != PM_PAIRED && a.mode != PM_LISTP) {
        if (a.out_m) a.out_m[i * a.ld + j] = (double)mmax;
        if (a.mode == PM_SELF && a.mirror) a.out[(j - a.row0) * a.ld + (i + a.row0)] = r;
      } else if (a.mode == PM_LISTP && a.out_m) {
        a.out_m[t * 32 + lane] = (double)mmax;
      }
    }
    __syncwarp();
  }
}

// ---- cooperative engine: G lanes per pair, the band row in the lanes' registers, neighbours by warp shuffle ----
template <class M, int W, int U>
struct CoopDevGroup {
  using F = typename M::real;
  CoopLane<M, W, U>& L;
  const F* x; const F* y;
  int G;
  CoopTmp<M, W, U> tmp;
  template <class Fn> __device__ __forceinline__ void each(Fn f) { f(L, x, y, tmp); }
  __device__ __forceinline__ void xchg_left() { L.left_in = __shfl_up_sync(0xffffffffu, L.last_out, 1, G); }
  __device__ __forceinline__ void xchg_up() { L.up_in = __shfl_down_sync(0xffffffffu, L.c0, 1, G); }
};

// G (power of two, <= 32) lanes cooperate on one pair; a warp task = 32 / G consecutive pairs (pair e = i * ny + j: the
// groups of a warp share the x row).  All pairs of a launch share the geometry, so every group of every warp runs the same
// step sequence (full-mask shuffles).
template <class M, int W, int U, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_coop(KArgsT<typename M::real> a, M m, int G, CoopLayout lay) {
  using F = typename M::real;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int slot = lane / G;
  const int PG = 32 / G;
  const long long npairs = (a.mode == PM_LISTP) ? (long long)__ldg(a.list_len) : a.npairs;
  const long long ntasks = (npairs + PG - 1) / PG;
  for (;;) {
    const long long t = next_task(a.counter, lane);
    if (t >= ntasks) break;
    long long e = t * PG + slot;
    bool valid = e < npairs;
    if (!valid) e = npairs - 1;
    long long i, j;
    if (a.mode == PM_LISTP) { const int2 p = a.list[e]; i = p.x; j = p.y; }
    else if (a.mode == PM_PAIRED) { i = e; j = e; }
    else { i = e / a.ny; j = e - i * a.ny; }
    if (a.mode == PM_SELF && j <= i + a.row0) valid = false;
    if (__ballot_sync(0xffffffffu, valid) == 0u) continue;
    M mm = m;
    PairCtx pc;
    pc.sx = a.sx ? a.sx[i] : 0.0;
    pc.sy = a.sy ? a.sy[j] : 0.0;
    pc.sy2 = a.sy2 ? a.sy2[j] : 0.0;
    mm.begin_pair(pc);
    CoopLane<M, W, U> L;
    const F* const xp = a.x + i * a.Tx;
    const F* const yp = a.y + j * a.ys;
    L.init(a.g, mm, lay, gl, xp, yp);
    CoopDevGroup<M, W, U> grp{L, xp, yp, G, {}};
    coop_run<M, W, U>(grp, a.g, mm, lay);
    if (valid && L.holds_result(a.g)) {
      const double d = (double)mm.finish(L.result(a.g), a.g);
      double* const po = (a.mode == PM_PAIRED) ? &a.out[i] : (a.mode == PM_LISTP ? &a.out[e] : &a.out[i * a.ld + j]);
      const double r = combine_dims(a, po, d);
      __stcs(po, r);
      if (a.mode == PM_SELF && a.mirror) __stcs(&a.out[(j - a.row0) * a.ld + (i + a.row0)], r);
    }
    __syncwarp();
  }
}

// ---- prologues ----
// EL:3220-3225: d[k] = ((q[k+1]-q[k]) + ((q[k+2]-q[k])/2))/2, k = 0..T-3
__global__ void k_slope(const double* __restrict__ q, long long n, int T, double* __restrict__ d) {
  const long long total = n * (long long)(T - 2);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long s = e / (T - 2);
    const int k = (int)(e - s * (T - 2));
    const double* p = q + s * T + k;
    d[e] = ((p[1] - p[0]) + ((p[2] - p[0]) / 2)) / 2;
  }
}

// fp32 mode: operands are converted once per call
__global__ void k_to_float(const double* __restrict__ src, long long n, float* __restrict__ dst) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    dst[e] = (float)src[e];
}

// kind 0: erp gap sum  sum_t |x[t] - g| (EL:1295-1303, sequential order)
// kind 1: std of the series (utils/_stats.pyx:22-42: sequential sums, threshold 1e-13)
// stride: elements between consecutive series (T for dense rows; 1 = every sliding window of a flat buffer)
__global__ void k_series_stat(const double* __restrict__ x, long long n, int T, int kind, double g,
                              double* __restrict__ out, long long stride) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double* p = x + s * stride;
  if (kind == 0) {
    double acc = 0;
    for (int t = 0; t < T; ++t) acc += fabs(p[t] - g);
    out[s] = acc;
  } else {
    double ex = 0, ex2 = 0;
    for (int t = 0; t < T; ++t) { const double v = p[t]; ex += v; ex2 += v * v; }
    const double mean = ex / (double)T;
    ex2 = ex2 / (double)T - mean * mean;
    out[s] = ex2 > 1e-13 ? sqrt(ex2) : 0.0;
  }
}

// ---- scaled subsequence search: running window statistics of every sample, in the reference's order (EL:375-401,
// 472-474: ex / ex2 accumulate sample by sample and the oldest sample is subtracted again) ----
// mean[i * T + w], stdv[i * T + w] for the windows w = 0 .. T - m of sample i; one thread per sample.
__global__ void k_window_stats(const double* __restrict__ x, long long n, int T, int m, double* __restrict__ mean,
                               double* __restrict__ stdv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = x + i * T;
  double ex = 0.0, ex2 = 0.0;
  for (int t = 0; t < T; ++t) {
    const double v = p[t];
    ex += v;
    ex2 += v * v;
    if (t >= m - 1) {
      const int w = t - (m - 1);
      const double mu = ex / (double)m;
      const double tmp = ex2 / (double)m - mu * mu;
      mean[i * T + w] = mu;
      stdv[i * T + w] = tmp > 0 ? sqrt(tmp) : 1.0;
      const double old = p[w];
      ex -= old;
      ex2 -= old * old;
    }
  }
}

// LB_Kim of the UCR-suite scan, `constant_lower_bound` EL:158-225, for every window of every sample (same layout as the
// DP values: entry i * T + w).  Term by term as the reference writes it -- its third term lists dist(t_y1, s_y1) twice
// and never dist(t_y1, s_y0), so the value can exceed the DTW distance; the scan skips a window when it is >= the running
// minimum (EL:413), which makes it part of the observable result.  sn = z-normalised subsequence (m >= 3 values).
__global__ void k_ucr_kim(const double* __restrict__ x, long long n, int T, int m, const double* __restrict__ sn,
                          const double* __restrict__ mean, const double* __restrict__ stdv, double* __restrict__ lb,
                          long long sn_stride = 0) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = T - m + 1;
  if (e >= n * (long long)nw) return;
  const long long i = e / nw;
  const int w = (int)(e - i * nw);
  sn += i * sn_stride;  // 0: one subsequence for every sample; m: subsequence i for sample i
  const double* t = x + i * T + w;
  const double mu = mean[i * T + w], sd = stdv[i * T + w];
  auto d2 = [](double a, double b) { const double q = a - b; return q * q; };
  const double t_x0 = (t[0] - mu) / sd, t_y0 = (t[m - 1] - mu) / sd;
  const double s_x0 = sn[0], s_y0 = sn[m - 1];
  double v = d2(t_x0, s_x0) + d2(t_y0, s_y0);
  const double t_x1 = (t[1] - mu) / sd, s_x1 = sn[1];
  v += dmin2(dmin2(d2(t_x1, s_x0), d2(t_x0, s_x1)), d2(t_x1, s_x1));
  const double t_y1 = (t[m - 2] - mu) / sd, s_y1 = sn[m - 2];
  v += dmin2(dmin2(d2(t_y1, s_y1), d2(t_y0, s_y1)), d2(t_y1, s_y1));
  const double t_x2 = (t[2] - mu) / sd, s_x2 = sn[2];
  v += dmin2(dmin2(d2(t_x0, s_x2), dmin2(dmin2(d2(t_x1, s_x2), d2(t_x2, s_x2)), d2(t_x2, s_x1))), d2(t_x2, s_x0));
  const double t_y2 = (t[m - 3] - mu) / sd, s_y2 = sn[m - 3];
  v += dmin2(dmin2(d2(t_y0, s_y2), dmin2(dmin2(d2(t_y1, s_y2), d2(t_y2, s_y2)), d2(t_y2, s_y1))), d2(t_y2, s_y0));
  lb[i * T + w] = v;
}

// subsequence search: (minimum, window) of the replayed scan -> (sqrt when the scan ran in the squared-cost domain, index)
__global__ void k_finish_scan(const double* __restrict__ hval, const long long* __restrict__ hidx, long long n,
                              double* __restrict__ out_dist, long long* __restrict__ out_idx, long long ld, int apply_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_dist[i * ld] = apply_sqrt ? sqrt(hval[i]) : hval[i];
  out_idx[i * ld] = hidx[i];
}

// grouped form: query q = g * nr + i (subsequence g of the launch, sample i of the pass) -> out[i * ld + ks[g]]
__global__ void k_finish_scan_group(const double* __restrict__ hval, const long long* __restrict__ hidx, long long nr, long long ng,
                                    const int* __restrict__ ks, double* __restrict__ out_dist, long long* __restrict__ out_idx,
                                    long long ld, int apply_sqrt) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nr * ng) return;
  const long long g = q / nr, i = q - g * nr;
  const long long o = i * ld + ks[g];
  out_dist[o] = apply_sqrt ? sqrt(hval[q]) : hval[q];
  out_idx[o] = hidx[q];
}

// ---- generic scaled subsequence metrics (ScaledSubsequenceMetricWrap, CD:470-551) ----
// Window statistics with the reference's IncStats (utils/_stats.pyx:45-93: Welford add / remove in scan order, variance
// below 1e-13 -> 0 -> std 1).  mean[i * nw + w], stdv[i * nw + w], nw = T - m + 1; one thread per sample.
__global__ void k_inc_window_stats(const double* __restrict__ x, long long n, int T, int m, double* __restrict__ mean,
                                   double* __restrict__ stdv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long nw = T - m + 1;
  inc_window_stats_one(x + i * T, T, m, mean + i * nw, stdv + i * nw);
}

// dst[((e >> 5) * T + t) * 32 + (e & 31)] = src[e * T + t]: n dense series -> groups of 32 interleaved series (KArgs::yil)
__global__ void k_interleave32(const double* __restrict__ src, long long n, int T, double* __restrict__ dst) {
  const long long total = interleave32_size(n, T);  // the last group is padded (its missing series are never read)
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    long long e; int t;
    interleave32_source(o, T, &e, &t);
    if (e < n) dst[o] = src[e * T + t];
  }
}

// out[(i * nw + w) * m + j] = (x[i * T + w + j] - mean) / std: the z-normalised windows as dense rows (CD:526-527)
__global__ void k_normalise_windows(const double* __restrict__ x, long long n, int T, int m, const double* __restrict__ mean,
                                    const double* __restrict__ stdv, double* __restrict__ out) {
  const int nw = T - m + 1;
  const long long total = n * (long long)nw * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long win = e / m;
    const int j = (int)(e - win * m);
    const long long i = win / nw;
    const int w = (int)(win - i * nw);
    out[e] = (x[i * T + w + j] - mean[win]) / stdv[win];
  }
}

// ---- subsequence matches / distance profile (SubsequenceMetric._matches, CD:311-372 and the *_subsequence_matches of EL) ----
// pair list of one pass: entry e = i * nw + w pairs subsequence (paired ? sub0 + i : 0) with window w of sample i, whose
// first element sits at y[i * ystride + w] (flat series, stride-1 windows) or row i * nw + w (ystride == nw: dense rows)
__global__ void k_profile_list(int2* __restrict__ list, long long n, int nw, long long ystride, int sub0, int paired) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / nw;
    const int w = (int)(e - i * nw);
    list[e] = make_int2(paired ? sub0 + (int)i : 0, (int)(i * ystride + w));
  }
}
// pair list of the first c1 windows of every (subsequence g, sample i) query of a scan pass: entry e = (g * nr + i) * c1 + w
// pairs subsequence g with window w of sample i, whose first element sits at y[i * ystride + w]
__global__ void k_scan_head_list(int2* __restrict__ list, long long n, long long nr, int c1, long long ystride) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long q = e / c1;
    const int w = (int)(e - q * c1);
    const long long g = q / nr, i = q - g * nr;
    list[e] = make_int2((int)g, (int)(i * ystride + w));
  }
}
// out[e] = the window's distance when the reference reports it, NaN otherwise: d <= thr_d (strict: d < thr_d), not abandoned
// (M > thr_m), not skipped by scaled_dtw's LB_Kim prefilter (kim >= thr_d; kim laid out [i * ldk + w]).
__global__ void k_profile_select(const double* __restrict__ d, const double* __restrict__ mm, const double* __restrict__ kim,
                                 long long n, int nw, long long ldk, double thr_d, double thr_m, int strict, int apply_sqrt,
                                 double* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const double v = d[e];
    bool ok = strict ? (v < thr_d) : (v <= thr_d);
    if (ok && mm) ok = !(mm[e] > thr_m);
    if (ok && kim) { const long long i = e / nw; ok = kim[i * ldk + (e - i * nw)] < thr_d; }
    out[e] = ok ? (apply_sqrt ? sqrt(v) : v) : __longlong_as_double(0x7ff8000000000000LL);
  }
}

// ---- subsequence search: first minimum over the windows of every sample (EL:622-660: `dist < min_dist`, strict) ----
// raw: DP values of ALL windows of the flat (n, T) buffer, window w of sample i at raw[i * T + w]; only w < nw are
// windows of sample i (the rest straddle two samples).  One warp per sample.
// ks != nullptr (grouped launch): row i = g * nr + s is sample s of the group's subsequence g and goes to out[s * ld + ks[g]].
__global__ void __launch_bounds__(128) k_window_min(const double* __restrict__ raw, long long n, int T, int nw,
                                                    double* __restrict__ out_dist, long long* __restrict__ out_idx,
                                                    long long ld, int apply_sqrt, const int* __restrict__ ks = nullptr,
                                                    long long nr = 0) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const double* p = raw + i * T;
  double best = WB_INF;
  int bidx = 0x7fffffff;
  for (int w = lane; w < nw; w += 32) {
    const double v = p[w];
    if (v < best) { best = v; bidx = w; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  if (lane == 0) {
    // every window +inf (cannot happen for finite input): the reference leaves the index untouched; report 0
    long long o = i * ld;
    if (ks) { const long long g = i / nr; o = (i - g * nr) * ld + ks[g]; }
    out_dist[o] = apply_sqrt ? sqrt(best) : best;
    out_idx[o] = bidx == 0x7fffffff ? 0 : bidx;
  }
}

// ---- FP64 issue-rate microbenchmark (roofline denominator) ----
// mix 0: 8 independent DADD chains.  mix 1: the DTW cell's instruction mix on 4 independent
// chains: sub, mul, add + two compare/selects (min) per "cell".
__global__ void __launch_bounds__(256) k_fp64_peak(int mix, int iters, double seed, double* out,
                                                    unsigned long long* cyc) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double inc = seed * 1e-9 + 1e-7;
  const unsigned long long c0 = clock64();
  if (mix == 0) {
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
      // volatile so that the clock reads bracket the arithmetic
      asm volatile("add.f64 %0, %0, %8; add.f64 %1, %1, %8; add.f64 %2, %2, %8; add.f64 %3, %3, %8;"
                   "add.f64 %4, %4, %8; add.f64 %5, %5, %8; add.f64 %6, %6, %8; add.f64 %7, %7, %8;"
                   : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(inc));
    }
  } else {
    double y0 = seed * 0.5, y1 = seed * 0.25, y2 = seed * 0.125, y3 = seed * 0.0625;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
      double v, m;
      v = a0 - y0; m = a4 < a0 ? a4 : a0; m = m < a1 ? m : a1; a0 = m + v * v;
      v = a1 - y1; m = a5 < a1 ? a5 : a1; m = m < a2 ? m : a2; a1 = m + v * v;
      v = a2 - y2; m = a6 < a2 ? a6 : a2; m = m < a3 ? m : a3; a2 = m + v * v;
      v = a3 - y3; m = a7 < a3 ? a7 : a3; m = m < a0 ? m : a0; a3 = m + v * v;
    }
  }
  const unsigned long long c1 = clock64();
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  out[gtid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (gtid == 0) cyc[0] = c1 - c0;
}

}  // namespace wb
