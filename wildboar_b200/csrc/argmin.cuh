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
#include "dispatch.cuh"
#include "kernels.cuh"

namespace wb {

struct ArgminIo {
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

// LaunchFn(r0, nrows, c0, ncols, out, ld, out_m, thr, stats) -> int
template <class WS, class DI, class Call, class LaunchFn>
int run_argmin(WS& ws, const DI& di, Call& c, const ArgminIo& io, wb_stats* stats, LaunchFn launch) {
  (void)di;
  cudaStream_t st = ws.stream;
  const long long nq = c.nx, ny = c.ny;
  const int k = (int)io.k;
  int kind; double scale = 1.0;
  const bool dtwfam = is_dtw_family(c.metric);
  if (dtwfam) kind = (c.metric == M_ADTW && c.p.p < 0) ? TK_NONE : TK_SQUARE;
  else if (c.metric == M_LCSS || c.metric == M_WLCSS) { kind = TK_LCSS; scale = (double)std::min(c.Tx, c.Ty); }
  else if (c.metric == M_EDR) { kind = TK_SCALE; scale = (double)std::max(c.Tx, c.Ty); }
  else kind = TK_IDENT;

  double *tau = nullptr, *thr = nullptr, *hval = nullptr, *dbuf = nullptr, *mbuf = nullptr, *lbuf = nullptr;
  long long* hidx = nullptr; int* hn = nullptr;
  long long C = (4LL << 20) / std::max<long long>(nq, 1);
  C = std::max<long long>(32, std::min<long long>(4096, (C / 32) * 32));
  C = std::min<long long>(C, ((ny + 31) / 32) * 32);
  if (ws.alloc(&tau, (size_t)nq) || ws.alloc(&thr, (size_t)nq) || ws.alloc(&hval, (size_t)nq * k) ||
      ws.alloc(&hidx, (size_t)nq * k) || ws.alloc(&hn, (size_t)nq) || ws.alloc(&dbuf, (size_t)nq * C)) return 1;
  if (!dtwfam && ws.alloc(&mbuf, (size_t)nq * C)) return 1;
  if (io.lower_bound && ws.alloc(&lbuf, (size_t)nq * C)) return 1;
  k_fill<<<256, 256, 0, st>>>(tau, nq, WB_INF);
  if (cudaMemsetAsync(hval, 0, sizeof(double) * nq * k, st) != cudaSuccess ||
      cudaMemsetAsync(hidx, 0, sizeof(long long) * nq * k, st) != cudaSuccess ||
      cudaMemsetAsync(hn, 0, sizeof(int) * nq, st) != cudaSuccess) return 1;

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  int rc = 0;
  for (long long c0 = 0; c0 < ny && !rc; c0 += C) {
    const long long nc = std::min(C, ny - c0);
    k_thr_raw<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(tau, nq, kind, scale, thr);
    if (c.degenerate) {
      // ddtw with T < 3: eadistance() returns False for every pair (EL:3297-3298)
      k_fill<<<256, 256, 0, st>>>(dbuf, nq * C, WB_INF);
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
  }
  cudaEventRecord(e1, st);
  if (!rc) {
    if (cudaMemcpyAsync(io.out_idx, hidx, sizeof(long long) * nq * k, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(io.out_dist, hval, sizeof(double) * nq * k, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
    float f = 0;
    if (!rc && stats && cudaEventElapsedTime(&f, e0, e1) == cudaSuccess) stats->kernel_ms += f;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

}  // namespace wb
