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
//
// What keeps the chunks cheap (dtw, ddtw, adtw with p >= 0):
//   * LB cascade: LB_Kim, then LB_Keogh in both directions, in rigorous round-down fp32 on outward-rounded operands,
//     as a register tile -- 4 queries x 32 references per warp, the query rows staged in shared memory by bulk copies
//     (cp.async.bulk + mbarrier, two tasks ahead), stragglers finished one pair per lane (k_lb_prune_tile); the survivors
//     are appended to the DP's work list by the pass itself.
//   * threshold seeding (k = 1; 1 < k <= 8 where the caller consumes the neighbours as a set or sorted): the exact
//     distances to sketch-nearest candidates bound the final threshold from the start (k_seed_candidates).
//   * host-resident references are uploaded piecewise on the copy stream while earlier chunks compute (ensure_refs).
// Tuning / test knobs (environment, read per call): WILDBOAR_CUDA_ARGMIN_CHUNK / _ARGMIN_FIRST (chunk widths), _NO_SEED,
// _SEED_MIN (references needed to seed), _PIPED_UPLOAD_KB (piece size, 0 = whole upload first), _LB_Q / _LB_BS / _LB_MINB /
// _LB_RB (tile shape: queries per warp, steps per register block, CTAs per SM, reference blocks per CTA task),
// _LB_STRAG "n,after" (straggler rule), _LB_KEEP (never drop a pass that does not prune), _ENVELOPE_PLAIN.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "dispatch.cuh"
#include "kernels.cuh"
#include "stage.hpp"

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
  int use_device_lb;          // bit 0: on-device LB cascade (dtw); bit 1: neighbour-SET mode (include/wb_cuda.h, wb_cuda_argmin)
  // Pipelined upload of the references (second operand in HOST memory, dtw / wdtw / adtw in fp64): the scan starts as soon
  // as the first chunk's references are resident; the remaining rows are copied piecewise on `up_stream` while earlier
  // chunks compute (run_argmin, `ensure_refs`).  y_host == nullptr: the references are already on the device.
  const double* y_host = nullptr;  // (ny, T) rows, `y_hs` elements apart
  int64_t y_hs = 0;
  double* y_dev = nullptr;         // destination == the call's dense y (ny, T)
  cudaStream_t up_stream = nullptr;
  // pageable y_host: two page-locked staging buffers of piped_piece_bytes() each (filled by HostCopyPool, stage.hpp); nullptr:
  // y_host is page-locked already, or no staging memory was to be had -- the pieces are copied straight from y_host
  char* stage_buf[2] = {nullptr, nullptr};
};

// bytes per upload piece (test knob WILDBOAR_CUDA_PIPED_UPLOAD_KB forces many pieces on small inputs; 0 disables the pipeline)
inline long long piped_piece_bytes() {
  if (const char* e = getenv("WILDBOAR_CUDA_PIPED_UPLOAD_KB")) return std::max<long long>(0, atoll(e)) << 10;
  return 16LL << 20;
}

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
// `seed` (optional, raw domain): an upper bound of the final threshold known in advance (run_argmin, threshold seeding)
__global__ void k_thr_raw(const double* __restrict__ tau, long long n, int kind, double scale,
                          double* __restrict__ thr, const double* __restrict__ seed = nullptr) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double t = (kind == TK_LCSS || kind == TK_NONE) ? WB_INF : ea_threshold(kind, tau[q], scale);
  if (seed) t = fmin(t, seed[q]);
  thr[q] = t;
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
  // neighbour-set mode (ArgminIo::use_device_lb bit 1): amb[q] = number of pairs left out of the heap although their
  // distance EQUALS the heap's current maximum (rejected by the strict `<`, or evicted while a twin stayed); reset when the
  // maximum drops.  Non-zero after the last chunk <=> which of the tied pairs is in the final set depends on the scan's
  // history (in the reference as well) -- the caller repeats those queries with the exact scan.  nullptr: not tracked.
  int* amb = nullptr;
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
    int amb = a.amb ? a.amb[q] : 0;
    // eight groups of 32 columns per trip: the loads of a trip are issued together (one warp walks a whole row, and with
    // one load in flight at a time the kernel ran at the latency of a dependent chain: 33 us for a 33 MB chunk matrix)
    for (long long j0 = 0; j0 < a.ncols; j0 += 256) {
      double dvs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long j = j0 + u * 32 + lane;
        dvs[u] = (j < a.ncols) ? __ldcs(drow + j) : WB_INF;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long jj = j0 + u * 32;
        const double dv = dvs[u];
        // t only decreases while the group is walked, so the candidates under the group-start t are a superset
        unsigned mask = __ballot_sync(0xffffffffu, a.amb ? (dv <= t && dv < WB_INF) : (dv < t));
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const double ds = __shfl_sync(0xffffffffu, dv, src);
          const long long js = jj + src;
          bool acc = ds < t;
          if (a.amb && !acc && ds == t && n == a.k) ++amb;  // tied with the current maximum and left out
          if (acc && a.lb) acc = !(a.lb[q * a.ld + js] >= t);
          if (acc && a.m) acc = !(a.m[q * a.ld + js] > ea_threshold(a.kind, t, a.scale));
          if (acc) {
            const double t_old = t;
            const int n_old = n;
            if (lane == 0) {
              heap_push(hi, hv, n, a.k, a.c0 + js, ds);
              t = (n == a.k) ? hv[0] : WB_INF;
            }
            t = __shfl_sync(0xffffffffu, t, 0);
            n = __shfl_sync(0xffffffffu, n, 0);
            if (a.amb) {
              if (n_old == a.k && t == t_old) ++amb;  // the evicted maximum had a twin that stays
              else if (t != t_old) amb = 0;           // the maximum dropped: earlier ties are below... above it now
            }
          }
        }
      }
    }
    if (a.amb && lane == 0) a.amb[q] = amb;
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
// series [s0, s0 + ns) of the n references (the whole set, or the piece that has just been uploaded)
__global__ void k_envelope_casc(const double* __restrict__ y, long long n, long long s0, long long ns, int T, int w, int stride,
                                float2* __restrict__ envT, float2* __restrict__ yvT, double* __restrict__ y0, double* __restrict__ yL) {
  const long long total = ns * (long long)T;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)(e / ns);
    const long long s = s0 + (e - (long long)pos * ns);
    const int k = (int)(((long long)pos * stride) % T);
    const double* p = y + s * T;
    const int a = max(0, k - w), b = min(T - 1, k + w);
    double l = p[a], h = p[a];
    for (int q = a + 1; q <= b; ++q) { const double v = p[q]; l = fmin(l, v); h = fmax(h, v); }
    const long long o = (long long)pos * n + s;
    envT[o] = make_float2(__double2float_rd(l), __double2float_ru(h));
    yvT[o] = make_float2(__double2float_rd(p[k]), __double2float_ru(p[k]));
    if (pos == 0) { y0[s] = p[0]; yL[s] = p[T - 1]; }
  }
}
// The same operands through a shared-memory tile: a CTA takes 32 consecutive series, reads their rows coalesced into a
// tile (row stride T + 1: lanes = series hit distinct banks), and writes the transposed arrays with lanes = series.  The
// plain kernel above reads every window straight from global memory with lanes = series, i.e. 32 sectors per load
// (5.1 ms for 200 000 x 256, 240 GB/s; this one moves the 1.2 GB at memory speed).
__global__ void __launch_bounds__(256) k_envelope_casc_tile(const double* __restrict__ y, long long n, long long s0, long long ns, int T,
                                                            int w, int stride, float2* __restrict__ envT, float2* __restrict__ yvT,
                                                            double* __restrict__ y0, double* __restrict__ yL) {
  extern __shared__ double env_tile[];  // 32 x (T + 1)
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int ldt = T + 1;
  const long long nblk = (ns + 31) / 32;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long sb = s0 + blk * 32;
    const int nsb = (int)min(32LL, s0 + ns - sb);
    const double* src = y + sb * T;
    for (int e = tid; e < nsb * T; e += 256) {
      const int r = e / T;
      env_tile[r * ldt + (e - r * T)] = src[e];
    }
    __syncthreads();
    if (lane < nsb) {
      const double* p = env_tile + lane * ldt;
      for (int pos = wrp; pos < T; pos += 8) {
        const int k = (int)(((long long)pos * stride) % T);
        const int a = max(0, k - w), b = min(T - 1, k + w);
        double l = p[a], h = p[a];
        for (int q = a + 1; q <= b; ++q) { const double v = p[q]; l = fmin(l, v); h = fmax(h, v); }
        const long long o = (long long)pos * n + sb + lane;
        envT[o] = make_float2(__double2float_rd(l), __double2float_ru(h));
        yvT[o] = make_float2(__double2float_rd(p[k]), __double2float_ru(p[k]));
        if (pos == 0) { y0[sb + lane] = p[0]; yL[sb + lane] = p[T - 1]; }
      }
    }
    __syncthreads();
  }
}
inline void launch_envelope_casc(cudaStream_t st, const double* y, long long n, long long s0, long long ns, int T, int w, int stride,
                                 float2* envT, float2* yvT, double* y0, double* yL) {
  const size_t smem = (size_t)32 * (T + 1) * sizeof(double);
  if (smem <= ((size_t)200 << 10) && !getenv("WILDBOAR_CUDA_ENVELOPE_PLAIN")) {
    cudaFuncSetAttribute(k_envelope_casc_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long nblk = (ns + 31) / 32;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)220 << 10) / smem));
    k_envelope_casc_tile<<<(unsigned)std::max<long long>(1, std::min<long long>(148LL * per_sm, nblk)), 256, smem, st>>>(y, n, s0, ns, T, w, stride, envT, yvT, y0, yL);
  } else {
    const long long nb = std::max<long long>(1, std::min<long long>(2048, (ns * T + 255) / 256));
    k_envelope_casc<<<(unsigned)nb, 256, 0, st>>>(y, n, s0, ns, T, w, stride, envT, yvT, y0, yL);
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
  const double* thr2;  // per query: the chunk's abandon threshold in the DP's raw (squared) domain, INF = none yet
  double* d; long long ld;  // chunk matrix, pre-filled with +INF by the caller (= pruned); the pass writes nothing into it:
                            // the survivors go to the work list and the DP stores their distances
  unsigned long long* n_kim; unsigned long long* n_keogh;  // pruning statistics
  // survivors are appended to `list` as (query, chunk column) through the cursor `list_len` (zeroed by the caller; one atomic
  // per (query, block) subtask): the DP's work list comes out of this pass directly -- no pass over the chunk matrix to
  // count and collect them (4 ms of a 36 ms call).  The order of the list is whatever order the warps finish in; every
  // pair is evaluated independently and lands in d[i][j], so the result does not depend on it.
  int2* list; int* list_len;
  unsigned long long* n_surv;  // statistics
  int strag_n, strag_after;  // a (query, block) subtask with <= strag_n unpruned lanes after step strag_after hands them to the DP
};

// ---- k_lb_prune: one warp = one query x 32 consecutive references (lane = reference) ----
// Task mapping: the eight warps of a CTA take EIGHT CONSECUTIVE QUERIES against the SAME block of 32 references, so the block's
// envelope / value rows (2 x 256 B per time step) come from L2 once and from L1 for the other seven warps.  With one query
// per CTA-wide row of reference blocks (the first version) every warp streamed its own rows from L2: 4 GB per launch of the
// cfg4 share = 9.4 TB/s, i.e. the pass ran at the L2's throughput (profiles/r02e_ncu_argmin_cfg4.csv: L1 hit 12 %,
// long_scoreboard 10.8 per issued instruction, issue active 36 %; now L1 hit 65 %, 0.435 -> 0.28 ms per launch).  The warps
// are not synchronised -- each leaves its task when its own 32 pairs are decided -- they merely start together.
// Kept as the fallback of k_lb_prune_tile for series too long for its shared-memory query tiles.
// warp-aggregated append of this warp's survivors (query i, chunk columns jl) to the DP's work list; returns their number
__device__ __forceinline__ int lb_append(const LbArgs& a, int lane, bool sv, long long i, long long jl) {
  const unsigned m = __ballot_sync(0xffffffffu, sv);
  if (!m) return 0;
  int base = 0;
  if (lane == 0) base = atomicAdd(a.list_len, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (sv) a.list[base + __popc(m & ((1u << lane) - 1u))] = make_int2((int)i, (int)jl);
  return __popc(m);
}
__global__ void __launch_bounds__(256) k_lb_prune(LbArgs a) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long nyb = (a.nc + 31) / 32;
  const long long nqg = (a.nq + wpb - 1) / wpb;
  const long long nct = nqg * nyb;
  const int T = a.T;
  unsigned c_kim = 0, c_keogh = 0, c_surv = 0;  // per-warp statistics (lane 0; a warp sees far fewer than 2^32 pairs per launch), one atomic per warp at the end
  for (long long ct = blockIdx.x; ct < nct; ct += gridDim.x) {
    const long long qg = ct / nyb;
    const long long i = qg * wpb + wib;
    if (i >= a.nq) continue;
    const long long jl = (ct - qg * nyb) * 32 + lane;
    const bool valid = jl < a.nc;
    const long long j = a.c0 + (valid ? jl : a.nc - 1);
    const double t2 = a.thr2[i];
    const double lim = isinf(t2) ? WB_INF : t2 * (1.0 + 1e-9);  // prune only if LB^2 > lim
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
        // blocks of 8 steps, software pipelined in halves of 4: while one half is being summed the loads of the next are in
        // flight, then one warp vote per block: leave when every lane is pruned -- or when only a straggler or two are left
        // after a good part of the series: finishing their sums would keep the whole warp busy, handing them to the
        // (early-abandoning) DP is cheaper.  Pruning is optional, so this cannot change a result.
        if (T >= 8) {
          float2 lhA[4], yyA[4], lhB[4], yyB[4]; float4 qqA[4], qqB[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { lhA[u] = env[u * ny]; yyA[u] = yv[u * ny]; qqA[u] = qf[u]; }
          for (; k + 8 <= T; k += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { lhB[u] = env[(4 + u) * ny]; yyB[u] = yv[(4 + u) * ny]; qqB[u] = qf[k + 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) step(lhA[u], yyA[u], qqA[u]);
            env += 8 * ny; yv += 8 * ny;
            if (k + 16 <= T) {  // first half of the next block (wasted when the warp leaves below)
#pragma unroll
              for (int u = 0; u < 4; ++u) { lhA[u] = env[u * ny]; yyA[u] = yv[u * ny]; qqA[u] = qf[k + 8 + u]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) step(lhB[u], yyB[u], qqB[u]);
            const int alive = __popc(__ballot_sync(0xffffffffu, !(pruned || s1 > limf || s2 > limf)));
            if (alive == 0 || (alive <= a.strag_n && k + 7 >= a.strag_after)) { k = T; break; }
          }
        }
        for (; k < T; ++k) { step(*env, *yv, qf[k]); env += ny; yv += ny; }
        p2 = pruned || s1 > limf || s2 > limf;
        c_keogh += __popc(__ballot_sync(0xffffffffu, p2 && !pruned && valid));
        pruned = p2;
      }
    }
    c_surv += lb_append(a, lane, valid && !pruned, i, jl);
  }
  if (lane == 0) {
    if (c_kim) atomicAdd(a.n_kim, (unsigned long long)c_kim);
    if (c_keogh) atomicAdd(a.n_keogh, (unsigned long long)c_keogh);
    if (c_surv) atomicAdd(a.n_surv, (unsigned long long)c_surv);
  }
}

// ---- k_lb_prune_tile<Q>: register tile of Q queries x 32 references per warp, query tiles staged in shared memory by TMA ----
// A CTA task is a group of Q CONSECUTIVE QUERIES against a range of reference blocks.  The Q queries' cascade rows (Q x T
// float4, contiguous in qf) arrive in shared memory by ONE bulk copy (cp.async.bulk + mbarrier), issued two tasks ahead
// into one of three buffers; the warps of the CTA draw reference blocks of the task from a shared counter.  A warp holds
// the block's envelope / value rows of eight time steps in registers (loaded once, the next block of eight in flight
// meanwhile) and runs all Q queries over them, the query samples coming as broadcast float4 reads from shared memory:
// 512 / Q + 16 B per pair-step instead of 528 B move through the SM's load path, and no load of the inner loop waits on L2.
// Every (query, block) subtask keeps its own sums and its own vote and leaves exactly where k_lb_prune leaves (same sums,
// same vote positions: the pruning statistics are bit-identical); the warp moves on when all Q are decided.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WB_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WB_MBAR_DONE;\n"
      "bra WB_MBAR_WAIT;\n"
      "WB_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

struct LbQMeta { double lim, x0, xL; };  // per query of a tile: prune limit (INF: no threshold yet), first / last sample

// a (query, reference) pair whose block of 32 was left by the straggler rule: continued lane-per-pair at the end of the CTA task
struct LbStrag { int q, jl, k; float s1, s2; };
// Task buffers per CTA.  A buffer is refilled when the LAST warp has left its task, so with two buffers a warp that is a whole
// task ahead of the slowest one finds nothing staged: ncu's source view had 10 % of the pass's samples (and 8 % of its
// instructions) in the mbarrier wait.  With three the warps run up to two tasks apart, and because they draw reference blocks
// from the task's counter the fast ones simply take more of the next task.
constexpr int kLbBufs = 3;
// queue entries per task buffer: every (query, block) subtask of a task can queue strag_n pairs, so nothing overflows
inline int lb_tile_qcap(int Q, int rb_per_task, int strag_n) { return Q * rb_per_task * std::max(strag_n, 0); }
inline size_t lb_tile_smem(int Q, int T, int qcap) {
  return kLbBufs * ((size_t)Q * T * sizeof(float4) + Q * sizeof(LbQMeta) + (size_t)qcap * sizeof(LbStrag)) + 128;  // + mbarriers, counters
}

// Q queries per warp task, BS time steps per register block (4: 3 CTAs of 8 warps per SM; 8: 2), MINB = CTAs per SM the
// register budget is cut for.  No CTA-wide barrier in the task loop: a buffer is handed back through a shared counter,
// and the LAST warp to finish a task stages the CTA's task after next into the buffer it has just freed.
template <int Q, int BS, int MINB>
__global__ void __launch_bounds__(256, MINB) k_lb_prune_tile(LbArgs a, int rb_per_task, int qcap) {
  static_assert(BS == 4 || BS == 8, "block of 4 or 8 time steps");
  extern __shared__ __align__(128) unsigned char lb_smem[];
  const int T = a.T;
  constexpr int NB = kLbBufs;
  float4* const qs = reinterpret_cast<float4*>(lb_smem);                        // [NB][Q * T]
  LbQMeta* const meta = reinterpret_cast<LbQMeta*>(qs + NB * (size_t)Q * T);    // [NB][Q]
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(meta + NB * Q);  // [NB]
  int* const s_next = reinterpret_cast<int*>(mbar + NB);                        // [NB] next reference block of the task
  int* const s_done = s_next + NB;                                              // [NB] warps that have finished the task
  int* const s_qn = s_done + NB;                                                // [NB] stragglers queued by the task
  LbStrag* const s_q = reinterpret_cast<LbStrag*>(s_qn + NB + (NB & 1));        // [NB][qcap]
  const int tid = threadIdx.x, lane = tid & 31, wpb = blockDim.x >> 5;
  const long long nyb = (a.nc + 31) / 32;
  const long long nqg = (a.nq + Q - 1) / Q;
  const long long nsplit = (nyb + rb_per_task - 1) / rb_per_task;
  const long long per = (nyb + nsplit - 1) / nsplit;
  const long long nct = nqg * nsplit;
  const long long ny = a.ny;
  unsigned c_kim = 0, c_keogh = 0, c_surv = 0;  // per-warp statistics: far fewer than 2^32 pairs per warp and launch

  // stage task `ct` into buffer b (one whole warp): lanes 0..Q-1 write the per-query scalars, lane 0 resets the block
  // counter and starts the bulk copy of the Q query rows; its arrive (release) publishes all of it with the data
  auto stage = [&](long long ct, int b) {
    const long long qg = ct / nsplit;
    const long long i0 = qg * Q;
    if (lane < Q) {
      LbQMeta mq; mq.lim = WB_INF; mq.x0 = 0.0; mq.xL = 0.0;
      const long long i = i0 + lane;
      if (i < a.nq) {
        const double t2 = a.thr2[i];
        mq.lim = isinf(t2) ? WB_INF : t2 * (1.0 + 1e-9);  // prune only if LB^2 > lim
        mq.x0 = a.x[i * T]; mq.xL = a.x[i * T + T - 1];
      }
      meta[b * Q + lane] = mq;
    }
    __syncwarp();
    if (lane == 0) {
      const unsigned bytes = (unsigned)(min((long long)Q, a.nq - i0) * T * (long long)sizeof(float4));
      s_next[b] = (int)((ct - qg * nsplit) * per);
      s_done[b] = 0;
      s_qn[b] = 0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer's last generic reads precede the async write
      mbar_expect_tx(&mbar[b], bytes);
      bulk_g2s(qs + (size_t)b * Q * T, a.qf + i0 * T, bytes, &mbar[b]);
    }
  };

  if (tid == 0) {
    for (int b = 0; b < NB; ++b) mbar_init(&mbar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {
    for (int b = 0; b < NB; ++b)
      if ((long long)blockIdx.x + (long long)b * gridDim.x < nct) stage((long long)blockIdx.x + (long long)b * gridDim.x, b);
  }
  int it = 0;
  for (long long ct = blockIdx.x; ct < nct; ct += gridDim.x, ++it) {
    const int b = it % NB;
    mbar_wait(&mbar[b], (unsigned)((it / NB) & 1));
    const long long qg = ct / nsplit;
    const long long i0 = qg * Q;
    const int rb_hi = (int)min(nyb, ((ct - qg * nsplit) + 1) * per);
    const float4* const qf = qs + (size_t)b * Q * T;
    const LbQMeta* const mt = meta + b * Q;
    for (;;) {
      int rb = 0;
      if (lane == 0) rb = atomicAdd(&s_next[b], 1);
      rb = __shfl_sync(0xffffffffu, rb, 0);
      if (rb >= rb_hi) break;
      const long long jl = (long long)rb * 32 + lane;
      const bool valid = jl < a.nc;
      const long long j = a.c0 + (valid ? jl : a.nc - 1);
      const float2* env = a.envT + j;
      const float2* yv = a.yvT + j;
      float2 lhA[BS], yyA[BS], lhB[BS], yyB[BS];
      auto load_block = [&](float2 (&lh)[BS], float2 (&yy)[BS]) {
#pragma unroll
        for (int u = 0; u < BS; ++u) { lh[u] = env[u * ny]; yy[u] = yv[u * ny]; }
      };
      const double y0 = a.y0[j], yL = a.yL[j];
      if (T >= 8) load_block(lhA, yyA);  // in flight during the LB_Kim tests (wasted when they prune the whole block for all Q)
      float limf[Q], s1[Q], s2[Q];
      unsigned active = 0;   // bit q (warp uniform): query q is still summing
      unsigned prm = 0;      // bit q (per lane): this lane's pair with query q is pruned
      unsigned dfr = 0;      // bit q (per lane): this lane's pair with query q was queued as a straggler (decided later)
      // LB_Kim: every warping path contains (0,0) and (T-1,T-1)
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        s1[q] = 0.0f; s2[q] = 0.0f; limf[q] = 0.0f;
        const LbQMeta mq = mt[q];
        if (isinf(mq.lim)) continue;  // no threshold yet (or no such query): nothing is pruned
        const double d0 = mq.x0 - y0;
        const double d1 = mq.xL - yL;
        const double lb = d0 * d0 + (T > 1 ? d1 * d1 : 0.0);
        const bool pk = lb > mq.lim;
        c_kim += __popc(__ballot_sync(0xffffffffu, pk && valid));
        if (pk) prm |= 1u << q;
        if (!__all_sync(0xffffffffu, pk)) { active |= 1u << q; limf[q] = __double2float_ru(mq.lim); }
      }
      if (active) {
        // LB_Keogh in BOTH directions at once (time steps in the permuted order): s1 = query against the reference's
        // envelope, s2 = reference against the query's envelope, in round-down fp32 on outward-rounded operands (rigorous
        // lower bounds of the fp64 sums).  A lane is pruned as soon as EITHER sum exceeds the limit.
        auto step = [&](float& t1, float& t2, const float2 lh, const float2 yy, const float4 qq) {
          const float e1 = fmaxf(fmaxf(__fsub_rd(qq.x, lh.y), __fsub_rd(lh.x, qq.y)), 0.0f);
          t1 = __fmaf_rd(e1, e1, t1);
          const float e2 = fmaxf(fmaxf(__fsub_rd(yy.x, qq.w), __fsub_rd(qq.z, yy.y)), 0.0f);
          t2 = __fmaf_rd(e2, e2, t2);
        };
        int k = 0;
        if (T >= 8) {
          // BS steps of every active query over one register block (the query samples: four broadcast reads from shared
          // memory issued together); after every 8th step one warp vote per query: leave when every lane is pruned -- or
          // when only a straggler or two are left after a good part of the series (handed to the early-abandoning DP)
          auto sum_block = [&](const float2 (&lh)[BS], const float2 (&yy)[BS], int kb, bool vote) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              if (!((active >> q) & 1u)) continue;
              const float4* qq = qf + q * T + kb;
#pragma unroll
              for (int h = 0; h < BS; h += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = qq[h + u];
#pragma unroll
                for (int u = 0; u < 4; ++u) step(s1[q], s2[q], lh[h + u], yy[h + u], v[u]);
              }
              if (vote) {
                const bool open = !(((prm >> q) & 1u) || s1[q] > limf[q] || s2[q] > limf[q]);
                const int alive = __popc(__ballot_sync(0xffffffffu, open));
                if (alive == 0) active &= ~(1u << q);
                else if (alive <= a.strag_n && kb + BS - 1 >= a.strag_after) {
                  // a straggler or two: the warp leaves, the open pairs are queued with their sums and continued one pair
                  // per lane when the CTA task ends (the queue holds strag_n entries per subtask of the task, so it cannot overflow)
                  active &= ~(1u << q);
                  if (open && valid && kb + BS < T) {
                    const int slot = atomicAdd(&s_qn[b], 1);
                    if (slot < qcap) {
                      LbStrag e; e.q = q; e.jl = (int)jl; e.k = kb + BS; e.s1 = s1[q]; e.s2 = s2[q];
                      s_q[b * qcap + slot] = e;
                      dfr |= 1u << q;
                    }
                  }
                }
              }
            }
          };
          // 16 steps per trip for BS = 8, 8 for BS = 4: the two register blocks swap roles without moves, the block after
          // the one being summed is in flight (its loads are wasted when the warp leaves after this block)
          while (k + 8 <= T && active) {
            env += BS * ny; yv += BS * ny;
            if (BS == 4 || k + 16 <= T) load_block(lhB, yyB);
            sum_block(lhA, yyA, k, BS == 8);
            k += BS;
            if (BS == 8 && !(k + 8 <= T && active)) break;
            env += BS * ny; yv += BS * ny;
            if (k + BS + 8 <= T) load_block(lhA, yyA);  // only when another trip follows
            sum_block(lhB, yyB, k, true);
            k += BS;
          }
        }
        // T % 8 last steps for the queries that ran to the end
        if (active) {
          for (; k < T; ++k) {
            const float2 lh1 = *env, yy1 = *yv;
            env += ny; yv += ny;
#pragma unroll
            for (int q = 0; q < Q; ++q)
              if ((active >> q) & 1u) step(s1[q], s2[q], lh1, yy1, qf[q * T + k]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const long long i = i0 + q;
        if (i >= a.nq) continue;
        const bool pk = (prm >> q) & 1u;
        // limf = 0 with zero sums where no Keogh pass ran (no threshold, or LB_Kim pruned all 32): 0 > 0 is false
        const bool p2 = pk || s1[q] > limf[q] || s2[q] > limf[q];
        const bool later = (dfr >> q) & 1u;  // queued: written by whoever drains the queue
        c_keogh += __popc(__ballot_sync(0xffffffffu, p2 && !pk && valid));
        c_surv += lb_append(a, lane, valid && !p2 && !later, i, jl);
      }
    }
    // hand the buffer back; the last warp drains the task's straggler queue and refills the buffer for the CTA's task after next
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence_block();
      last = atomicAdd(&s_done[b], 1) == wpb - 1;
      __threadfence_block();
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    __syncwarp();  // lane 0's fence + atomic order the other warps' queue entries before every lane's reads below
    if (last) {
      // stragglers of this task, one pair per lane: the rest of the two LB_Keogh sums from where the block left them.  The
      // task's reference rows are hot in L2, its query rows still in this buffer; 32 undecided pairs cost one warp a few
      // thousand instructions here against 32 full DPs (80 000 instructions each) if they were handed on.
      const int nqd = min(s_qn[b], qcap);
      for (int base = 0; base < nqd; base += 32) {
        const bool has = base + lane < nqd;
        LbStrag e; e.q = 0; e.jl = 0; e.k = T; e.s1 = 0.0f; e.s2 = 0.0f;
        if (has) e = s_q[b * qcap + base + lane];
        const float limq = __double2float_ru(mt[e.q].lim);
        const float4* qq = qf + e.q * T;
        const long long j = a.c0 + e.jl;
        const float2* env = a.envT + (long long)e.k * ny + j;
        const float2* yv = a.yvT + (long long)e.k * ny + j;
        float t1 = e.s1, t2 = e.s2;
        int k = e.k;
        auto step1 = [&](const float2 lh, const float2 yy, const float4 v) {
          const float e1 = fmaxf(fmaxf(__fsub_rd(v.x, lh.y), __fsub_rd(lh.x, v.y)), 0.0f);
          t1 = __fmaf_rd(e1, e1, t1);
          const float e2 = fmaxf(fmaxf(__fsub_rd(yy.x, v.w), __fsub_rd(v.z, yy.y)), 0.0f);
          t2 = __fmaf_rd(e2, e2, t2);
        };
        while (k + 8 <= T && !(t1 > limq || t2 > limq)) {
          float2 lh[8], yy[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { lh[u] = env[u * ny]; yy[u] = yv[u * ny]; }
#pragma unroll
          for (int u = 0; u < 8; ++u) step1(lh[u], yy[u], qq[k + u]);
          env += 8 * ny; yv += 8 * ny; k += 8;
        }
        if (!(t1 > limq || t2 > limq)) {
          for (; k < T; ++k) { step1(*env, *yv, qq[k]); env += ny; yv += ny; }
        }
        const bool p2 = t1 > limq || t2 > limq;
        c_keogh += __popc(__ballot_sync(0xffffffffu, has && p2));
        c_surv += lb_append(a, lane, has && !p2, i0 + e.q, e.jl);
      }
      if (ct + (long long)NB * gridDim.x < nct) stage(ct + (long long)NB * gridDim.x, b);
    }
  }
  if (lane == 0) {
    if (c_kim) atomicAdd(a.n_kim, (unsigned long long)c_kim);
    if (c_keogh) atomicAdd(a.n_keogh, (unsigned long long)c_keogh);
    if (c_surv) atomicAdd(a.n_surv, (unsigned long long)c_surv);
  }
}

// ------------------------------------------------------------------------------------------
// Threshold seeding (k = 1).  The cascade prunes with the threshold a query has at the START of a chunk, and that starts
// at +INF: the first chunks run dense or nearly so, and on the cfg4 share 2.1 % of all pairs reach the DP although only
// 0.08 % would with the final thresholds.  Any exact distance d(q, y_j*) is an upper bound of the final nearest-neighbour
// distance, so pruning / abandoning against min(running threshold, d*^2) can only remove pairs with d > d* >= the final
// minimum: for k = 1 the result (first index of the minimal distance, and that distance) cannot change -- the candidate
// itself and every tie survive (LB^2 <= d^2 = seed, column minima <= d^2 = seed, both tests are strict).  For k > 1 the
// SET of neighbours would be safe too but not the heap-array order the reference returns, so seeding is k = 1 only.
// Candidates: the 8 nearest references under the Euclidean distance of a 16-segment piecewise-aggregate sketch (fp32;
// a heuristic, it only decides how tight the seed is), then their exact DP values (raw, squared domain) -- 8 nq pairs.
// ------------------------------------------------------------------------------------------
constexpr int kSeedP = 16;   // sketch dimensions
constexpr int kSeedNC = 8;   // candidates per query
constexpr int kSeedQB = 8;   // queries per CTA of the candidate search

// sketch of series [0, n): mean of segment p = [p T / P, (p + 1) T / P); TR: transposed ([p][series]) for the references
template <bool TR>
__global__ void k_seed_sketch(const double* __restrict__ x, long long n, int T, float* __restrict__ out) {
  const long long total = n * kSeedP;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long s = e / kSeedP;
    const int p = (int)(e - s * kSeedP);
    const int a = (int)((long long)p * T / kSeedP), b = (int)((long long)(p + 1) * T / kSeedP);
    double acc = 0.0;
    for (int t = a; t < b; ++t) acc += x[s * T + t];
    const float v = (b > a) ? (float)(acc / (double)(b - a)) : 0.0f;
    out[TR ? (long long)p * n + s : e] = v;
  }
}

// per query the kSeedNC best of the 256 threads' own best reference (thread t scans references t, t + 256, ...)
// blockIdx.y = slice of the references (few queries: one CTA per query group would scan all of them alone -- 1.5 ms for
// 200 000 references); every (query, slice) contributes its own kSeedNC candidates
__global__ void __launch_bounds__(256) k_seed_candidates(const float* __restrict__ qp, const float* __restrict__ rp, long long nq,
                                                         long long S, int2* __restrict__ cl) {
  __shared__ float sq[kSeedQB][kSeedP];
  __shared__ float sv[8];
  __shared__ int st[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long q0 = (long long)blockIdx.x * kSeedQB;
  if (tid < kSeedQB * kSeedP) {
    const long long q = q0 + tid / kSeedP;
    sq[tid / kSeedP][tid % kSeedP] = q < nq ? qp[q * kSeedP + tid % kSeedP] : 0.0f;
  }
  __syncthreads();
  float best[kSeedQB]; int bidx[kSeedQB];
#pragma unroll
  for (int q = 0; q < kSeedQB; ++q) { best[q] = 3.0e38f; bidx[q] = -1; }
  const int nsl = gridDim.y, sl = blockIdx.y;
  const long long j_lo = S * sl / nsl, j_hi = S * (sl + 1) / nsl;
  for (long long j = j_lo + tid; j < j_hi; j += 256) {
    float r[kSeedP];
#pragma unroll
    for (int p = 0; p < kSeedP; ++p) r[p] = rp[(long long)p * S + j];
#pragma unroll
    for (int q = 0; q < kSeedQB; ++q) {
      float d = 0.0f;
#pragma unroll
      for (int p = 0; p < kSeedP; ++p) { const float v = r[p] - sq[q][p]; d = fmaf(v, v, d); }
      if (d < best[q]) { best[q] = d; bidx[q] = (int)j; }
    }
  }
#pragma unroll
  for (int q = 0; q < kSeedQB; ++q) {
    float v = best[q];
    for (int r = 0; r < kSeedNC; ++r) {
      // block-wide argmin of (v, thread)
      float wv = v; int wt = tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
        const int ot = __shfl_xor_sync(0xffffffffu, wt, o);
        if (ov < wv || (ov == wv && ot < wt)) { wv = ov; wt = ot; }
      }
      if (lane == 0) { sv[warp] = wv; st[warp] = wt; }
      __syncthreads();
      float bv = sv[0]; int bt = st[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) if (sv[w] < bv) { bv = sv[w]; bt = st[w]; }
      if (tid == bt && q0 + q < nq) {
        cl[((q0 + q) * nsl + sl) * kSeedNC + r] = make_int2((int)(q0 + q), bidx[q] < 0 ? (int)j_lo : bidx[q]);
        v = 3.4e38f;  // taken
      }
      __syncthreads();
    }
  }
}

// seed2[q] = the kth smallest (1-based, kth <= kSeedNC) of the query's candidate DP values
__global__ void k_seed_min(const double* __restrict__ cd, long long nq, int per_query, int kth, double* __restrict__ seed2) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double best[kSeedNC];  // ascending
#pragma unroll
  for (int r = 0; r < kSeedNC; ++r) best[r] = WB_INF;
  for (int e = 0; e < per_query; ++e) {
    double v = cd[q * per_query + e];
#pragma unroll
    for (int r = 0; r < kSeedNC; ++r) {
      if (v < best[r]) { const double tmp = best[r]; best[r] = v; v = tmp; }
    }
  }
  double out = best[0];
#pragma unroll
  for (int r = 1; r < kSeedNC; ++r) if (r < kth) out = best[r];
  seed2[q] = out;
}
__global__ void k_set_int(int* p, int v) { *p = v; }

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
  long long* hidx = nullptr; int* hn = nullptr; int* amb = nullptr;
  // Columns per chunk.  The thresholds a chunk is pruned with are those at its start, so the FIRST chunks should be small
  // (they run dense or nearly so) and the later ones large (few launches: with one query every chunk is ~6 launches of
  // almost no work).  The chunk grows 4x per step from 1024 columns up to C = min(32768, 4 Mi / nq) -- nq * C values per
  // buffer; for the cfg4 share (2500 queries) that is 128, 512, then 1600; for one query 1024, 4096, 16384, 32768, ...
  // Threshold seeding (k = 1, see k_seed_candidates) is decided here because it changes the schedule: with thresholds that
  // are close to final from the start, a stale chunk-start threshold costs little, so the chunks are four times as wide
  // (32 Mi values per buffer: an eighth of the launches; cfg4 share, seeded: kernels 55.7 ms at 1600 columns, 36.1 at 6400, 32.8 at 12 800) and there is no dense first chunk to keep small.
  // pipelined upload: the candidates are drawn from the first few pieces (the scan waits for them; with the staged copy a
  // piece of 16 MB arrives in half a millisecond)
  long long seed_pieces = 4;
  if (const char* e = getenv("WILDBOAR_CUDA_SEED_PIECES")) seed_pieces = std::max<long long>(1, atoll(e));
  const long long seed_from = io.y_host ? std::min<long long>(ny, seed_pieces * std::max<long long>(1, piped_piece_bytes() / (long long)(sizeof(double) * c.Ty))) : ny;
  long long seed_min = 2048;
  if (const char* e = getenv("WILDBOAR_CUDA_SEED_MIN")) seed_min = std::max<long long>(256, atoll(e));  // test knob
  // The cascade bounds the banded DTW of the PREPARED operands: dtw itself, ddtw (DTW of the slope series; R as
  // eadistance() takes it), and adtw with a penalty >= 0 (its cost is DTW's plus non-negative penalties, so every lower
  // bound of DTW is one of adtw; the survivors and the seeds go through adtw's own recurrence).  wdtw's weights start near
  // zero on the diagonal (w(0) = 0.0017 for g = 0.05, T = 256), which leaves nothing of the bound.
  const bool lb_metric = c.metric == M_DTW || c.metric == M_DDTW || (c.metric == M_ADTW && c.p.p >= 0);
  const bool lb_on = (io.use_device_lb & 1) != 0 && lb_metric;
  // neighbour-set mode: the caller needs the k nearest as a SET (class votes), so the thresholds may be seeded for k > 1 as
  // well -- with the kth smallest candidate distance, an upper bound of the final kth distance: only pairs outside the final
  // set are removed, the heap-array order is this scan's, and queries whose set is history dependent are counted (amb)
  const bool set_mode = (io.use_device_lb & 2) != 0 && k > 1 && k <= kSeedNC;
  // A caller-supplied lower_bound matrix rules seeding out: pairs with lower_bound >= threshold are SKIPPED by the
  // reference whatever their distance (the elastic ensemble masks each sample's own column with +inf that way,
  // ensemble/_elastic.py), so the distance to a candidate is no bound on what the scan can still accept.
  const bool will_seed = lb_on && !io.lower_bound && c.ptx == c.pty && c.ptx >= 2 && !c.degenerate && (k == 1 || set_mode) &&
                         seed_from >= seed_min && nq * (long long)kSeedNC < 2000000000LL && !getenv("WILDBOAR_CUDA_NO_SEED");
  // ... and the unseeded cascade gains as much from wide chunks once its thresholds have settled (heap-order k = 5 on the cfg4
  // share: 54.5 ms at 1664 columns, 31.3 ms at 12 800 with the first chunks still 128, 512, 2048, ...: a chunk costs its
  // launches, the staler thresholds add 7 % to the DP's pairs).  Without the cascade every pair of a chunk goes through the
  // DP and only its early abandoning profits from fresh thresholds, so those scans keep the narrow chunks.
  const bool casc_ok = lb_on && c.ptx == c.pty && c.ptx >= 2 && !c.degenerate;
  long long C = ((casc_ok ? 32LL : 4LL) << 20) / std::max<long long>(nq, 1);
  C = std::max<long long>(32, std::min<long long>(32768, (C / 32) * 32));
  // first chunk: it runs without thresholds (every pair in full), so just enough pairs to fill the device -- 256 Ki --
  // between 128 and 1024 columns (cfg4 share, 2500 queries: 128 columns; kernels 81 -> 75 ms against 1024)
  long long c_first = std::min<long long>(C, std::max<long long>(128, std::min<long long>(1024, (((256LL << 10) / std::max<long long>(nq, 1)) / 32) * 32)));
  if (const char* e = getenv("WILDBOAR_CUDA_ARGMIN_CHUNK")) {  // tuning / test knob: fixed columns per chunk
    const long long v = atoll(e);
    if (v >= 32) { C = (v / 32) * 32; c_first = C; }
  }
  if (const char* e = getenv("WILDBOAR_CUDA_ARGMIN_FIRST")) {  // tuning knob: columns of the first (dense) chunk
    const long long v = atoll(e);
    if (v >= 32) c_first = (v / 32) * 32;
  }
  C = std::min<long long>(C, ((ny + 31) / 32) * 32);
  c_first = std::min(c_first, C);
  if (ws.alloc(&tau, (size_t)nq) || ws.alloc(&thr, (size_t)nq) || ws.alloc(&hval, (size_t)nq * k) ||
      ws.alloc(&hidx, (size_t)nq * k) || ws.alloc(&hn, (size_t)nq) || ws.alloc(&dbuf, (size_t)nq * C)) return 1;
  if ((!dtwfam || adtw_neg) && ws.alloc(&mbuf, (size_t)nq * C)) return 1;
  if (io.lower_bound && ws.alloc(&lbuf, (size_t)nq * C)) return 1;
  // ---- optional on-device lower-bound cascade (dtw, equal lengths) ----
  const bool cascade = lb_on && c.ptx == c.pty && c.ptx >= 2 && !c.degenerate &&
                       nq * C < 2000000000LL;
  float4* qf = nullptr; float2 *envT = nullptr, *yvT = nullptr;
  double *y0 = nullptr, *yL = nullptr;
  int* list_len = nullptr; int2* list = nullptr;
  unsigned long long* lbstat = nullptr;  // [0] kim-pruned, [1] keogh-pruned, [2] survivors
  if (cascade) {
    const int T = c.ptx, w = std::max(c.R - 1, 0);
    if (ws.alloc(&qf, (size_t)nq * T) || ws.alloc(&list_len, 1) || ws.alloc(&list, (size_t)nq * C) ||
        ws.alloc(&lbstat, 3)) return 1;
    if (cudaMemsetAsync(lbstat, 0, 3 * sizeof(unsigned long long), st) != cudaSuccess) return 1;
    const int stride = lb_time_stride(T);
    k_query_casc<<<1024, 256, 0, st>>>(c.px, nq, T, w, stride, qf);
    LbCascCache* cc = c.metric == M_DDTW ? nullptr : io.casc_cache;  // ddtw: the operands are this call's slope buffers
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
        launch_envelope_casc(st, c.py, ny, 0, ny, T, w, stride, cc->envT, cc->yvT, cc->y0, cc->yL);
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
      // host-resident references: the operands are built piece by piece as the rows arrive (ensure_refs)
      if (!io.y_host) launch_envelope_casc(st, c.py, ny, 0, ny, T, w, stride, envT, yvT, y0, yL);
    }
  }
  // ---- pipelined upload of host-resident references ----
  long long up_done = io.y_host ? 0 : ny;  // references [0, up_done) are resident (and their cascade operands built)
  cudaEvent_t up_ev = nullptr;
  if (io.y_host) {
    // the destination was allocated in stream order on `st`: the copy stream may touch it only after that point
    if (cudaEventCreateWithFlags(&up_ev, cudaEventDisableTiming) != cudaSuccess) return 1;
    if (cudaEventRecord(up_ev, st) != cudaSuccess || cudaStreamWaitEvent(io.up_stream, up_ev, 0) != cudaSuccess) { cudaEventDestroy(up_ev); return 1; }
  }
  // make references [0, upto) resident before work that reads them is enqueued on `st`.  Pieces of >= 16 MB beyond what the
  // chunk needs: a copy from pageable memory blocks the HOST until it is staged, the device meanwhile runs the chunks
  // enqueued before it (a piece is several chunks of work, and is copied in less time than they take).
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  bool stage_used[2] = {false, false};
  int stage_n = 0;
  const bool staged = io.y_host && io.stage_buf[0] && io.stage_buf[1];
  if (staged && (cudaEventCreateWithFlags(&stage_ev[0], cudaEventDisableTiming) != cudaSuccess ||
                 cudaEventCreateWithFlags(&stage_ev[1], cudaEventDisableTiming) != cudaSuccess)) return 1;
  auto ensure_refs = [&](long long upto) -> int {
    if (upto <= up_done) return 0;
    const long long Ty = c.Ty;
    const long long piece = std::max<long long>(1, piped_piece_bytes() / (long long)(sizeof(double) * Ty));
    const long long to = std::min<long long>(ny, std::max(upto, up_done + piece));
    const long long rows = to - up_done;
    double* dst = io.y_dev + up_done * Ty;
    const double* src = io.y_host + up_done * io.y_hs;
    cudaError_t e = cudaSuccess;
    if (staged) {
      // pageable source: piece-sized parts through the two page-locked buffers -- several threads copy a part in, one DMA
      // takes it to the device while the next part is being copied
      for (long long r0 = 0; r0 < rows && e == cudaSuccess; r0 += piece) {
        const long long nr = std::min(piece, rows - r0);
        const int sb = stage_n++ & 1;
        if (stage_used[sb] && cudaEventSynchronize(stage_ev[sb]) != cudaSuccess) return 1;  // its last DMA has left the buffer
        HostCopyPool::get().copy_rows(io.stage_buf[sb], (const char*)(src + r0 * io.y_hs), (size_t)nr, sizeof(double) * (size_t)Ty,
                                      sizeof(double) * (size_t)io.y_hs);
        e = cudaMemcpyAsync(dst + r0 * Ty, io.stage_buf[sb], sizeof(double) * nr * Ty, cudaMemcpyHostToDevice, io.up_stream);
        if (e == cudaSuccess) e = cudaEventRecord(stage_ev[sb], io.up_stream);
        stage_used[sb] = true;
      }
    } else {
      e = (io.y_hs == Ty)
          ? cudaMemcpyAsync(dst, src, sizeof(double) * rows * Ty, cudaMemcpyHostToDevice, io.up_stream)
          : cudaMemcpy2DAsync(dst, sizeof(double) * Ty, src, sizeof(double) * io.y_hs, sizeof(double) * Ty, rows, cudaMemcpyHostToDevice, io.up_stream);
    }
    if (e != cudaSuccess || cudaEventRecord(up_ev, io.up_stream) != cudaSuccess || cudaStreamWaitEvent(st, up_ev, 0) != cudaSuccess) return 1;
    if (cascade && envT) {
      launch_envelope_casc(st, c.py, ny, up_done, rows, c.ptx, std::max(c.R - 1, 0), lb_time_stride(c.ptx), envT, yvT, y0, yL);
    }
    up_done = to;
    return 0;
  };
  k_fill<<<256, 256, 0, st>>>(tau, nq, WB_INF);
  if (cudaMemsetAsync(hval, 0, sizeof(double) * nq * k, st) != cudaSuccess ||
      cudaMemsetAsync(hidx, 0, sizeof(long long) * nq * k, st) != cudaSuccess ||
      cudaMemsetAsync(hn, 0, sizeof(int) * nq, st) != cudaSuccess) return 1;
  if (set_mode && (ws.alloc(&amb, (size_t)nq) || cudaMemsetAsync(amb, 0, sizeof(int) * nq, st) != cudaSuccess)) return 1;

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  int rc = 0;
  // ---- threshold seeding (k = 1; see k_seed_candidates) ----
  double* seed2 = nullptr;
  if (cascade && will_seed) {
    // references the candidates are drawn from: all of them when they are resident, the first piece of a pipelined upload
    if (io.y_host) rc = ensure_refs(seed_from);
    const long long S = up_done;
    if (!rc && S >= seed_min) {
      // reference slices per query group: enough CTAs to fill the device when the queries are few, each slice >= 256 references
      const long long ngroups = (nq + kSeedQB - 1) / kSeedQB;
      const int nsl = (int)std::max<long long>(1, std::min<long long>({32LL, (2LL * 148 + ngroups - 1) / ngroups, S / 256}));
      const long long ncand = nq * nsl * kSeedNC;
      float *rp = nullptr, *qp = nullptr; int2* cl = nullptr; int* cl_len = nullptr; double* cd = nullptr;
      if (ws.alloc(&rp, (size_t)S * kSeedP) || ws.alloc(&qp, (size_t)nq * kSeedP) || ws.alloc(&cl, (size_t)ncand) ||
          ws.alloc(&cl_len, 1) || ws.alloc(&cd, (size_t)ncand) || ws.alloc(&seed2, (size_t)nq)) rc = 1;
      if (!rc) {
        const int T = c.ptx;
        k_seed_sketch<true><<<(unsigned)std::min<long long>(148 * 16, (S * kSeedP + 255) / 256), 256, 0, st>>>(c.py, S, T, rp);
        k_seed_sketch<false><<<(unsigned)std::min<long long>(148 * 16, (nq * kSeedP + 255) / 256), 256, 0, st>>>(c.px, nq, T, qp);
        k_seed_candidates<<<dim3((unsigned)ngroups, (unsigned)nsl), 256, 0, st>>>(qp, rp, nq, S, cl);
        k_set_int<<<1, 1, 0, st>>>(cl_len, (int)ncand);
        const int mode0 = c.mode, raw0 = c.raw;
        c.mode = PM_LISTP; c.list = cl; c.list_len = cl_len; c.list_n = ncand; c.raw = 1;
        rc = launch(0, nq, 0, ny, cd, 1, nullptr, nullptr, stats);
        c.mode = mode0; c.list = nullptr; c.list_len = nullptr; c.list_n = 0; c.raw = raw0;
        k_seed_min<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(cd, nq, nsl * kSeedNC, k, seed2);
        if (stats) stats->launches += 5;
        // with thresholds from the start there is no dense first chunk to keep small
        c_first = C;
      }
    }
  }
  long long c_cur = c_first;
  // The cascade pays for itself only where the bounds bite: ddtw of noise-like series (the slopes of a random walk) has
  // envelopes that contain almost every sample, nothing is pruned and the pass is 7 % on top of the DP.  After the first
  // few thousand columns (thresholds have settled) the pruned fraction is read back ONCE; below a fifth the remaining
  // chunks run without the pass -- pruning is optional, so the result cannot change.
  bool casc_on = cascade, casc_checked = false;
  long long casc_cols = 0;
  const long long casc_check_after = std::min<long long>(4096, std::max<long long>(ny / 4, 1));
  for (long long c0 = 0; c0 < ny && !rc; ) {
    if (casc_on && !casc_checked && casc_cols >= casc_check_after && !getenv("WILDBOAR_CUDA_LB_KEEP")) {
      casc_checked = true;
      unsigned long long h[3] = {0, 0, 0};
      if (cudaMemcpyAsync(h, lbstat, sizeof h, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess &&
          (double)(h[0] + h[1]) < 0.2 * (double)nq * (double)casc_cols)
        casc_on = false;
    }
    const long long nc = std::min(c_cur, ny - c0);
    const long long c0_next = c0 + nc;
    c_cur = std::min(C, c_cur * 4);
    if ((rc = ensure_refs(c0_next))) break;
    k_thr_raw<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(tau, nq, kind, scale, thr, seed2);
    if (c.degenerate) {
      // ddtw with T < 3: eadistance() returns False for every pair (EL:3297-3298)
      k_fill<<<256, 256, 0, st>>>(dbuf, nq * C, WB_INF);
    } else if (casc_on && (c0 > 0 || seed2)) {
      casc_cols += nc;
      // chunk 0 has no threshold yet (tau = INF) unless the thresholds were seeded: nothing can be pruned, run it densely
      LbArgs la;
      la.x = c.px; la.qf = qf; la.envT = envT; la.yvT = yvT; la.y0 = y0; la.yL = yL;
      la.nq = nq; la.ny = ny; la.c0 = c0; la.nc = nc; la.T = c.ptx; la.thr2 = thr; la.d = dbuf; la.ld = C;
      la.n_kim = lbstat; la.n_keogh = lbstat + 1; la.n_surv = lbstat + 2; la.list = list; la.list_len = list_len;
      la.strag_n = kLbStragglers; la.strag_after = kLbStragglerAfter;
      if (const char* e = getenv("WILDBOAR_CUDA_LB_STRAG")) {  // tuning knob "n,after"
        int n_ = 0, a_ = 0;
        if (sscanf(e, "%d,%d", &n_, &a_) == 2) { la.strag_n = n_; la.strag_after = a_; }
      }
      if (cudaMemsetAsync(list_len, 0, sizeof(int), st) != cudaSuccess) { rc = 1; break; }
      // every pair of the chunk starts out as "pruned"; only the survivors' entries are overwritten (by the DP).  One streaming
      // fill instead of a 256-byte store per (query, block) subtask from inside the pass
      k_fill<<<148 * 8, 256, 0, st>>>(dbuf, nq * C, WB_INF);
      {
        // register-tiled pass with shared-memory query tiles when they fit (three buffers of Q x T float4), else one query per warp
        const char* lbq_env = getenv("WILDBOAR_CUDA_LB_Q");  // tuning / test knob: 0 = the one-query kernel
        int lbq = lbq_env ? atoi(lbq_env) : 4;
        const char* rbt_env = getenv("WILDBOAR_CUDA_LB_RB");
        const int rbt = (rbt_env && atoi(rbt_env) > 0) ? std::min(atoi(rbt_env), 64) : 32;  // reference blocks per CTA task (8 / 16 / 32: 29.8 / 28.9 / 28.5 ms, r02bb)
        la.strag_n = std::min(la.strag_n, 8);
        while (lbq > 1 && lb_tile_smem(lbq, c.ptx, lb_tile_qcap(lbq, rbt, la.strag_n)) > (size_t)96 << 10) lbq >>= 1;
        if (nq < 2 || lbq < 2) lbq = 0;
        const int qcap = lb_tile_qcap(lbq, rbt, la.strag_n);
        const char* bs_env = getenv("WILDBOAR_CUDA_LB_BS");
        const int lbs = (bs_env && atoi(bs_env) == 8) ? 8 : 4;  // time steps per register block
        auto go = [&](auto kern, int per_sm) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << 10);
          kern<<<148 * per_sm, 256, lb_tile_smem(lbq, c.ptx, qcap), st>>>(la, rbt, qcap);
        };
        const char* mb_env = getenv("WILDBOAR_CUDA_LB_MINB");
        const int mb = (mb_env && atoi(mb_env) == 2) ? 2 : 3;  // CTAs per SM the register budget is cut for (BS = 4)
        if (lbq >= 8) go(k_lb_prune_tile<8, 4, 2>, 2);
        else if (lbq >= 4) { if (lbs == 8) go(k_lb_prune_tile<4, 8, 2>, 2); else if (mb == 2) go(k_lb_prune_tile<4, 4, 2>, 2); else go(k_lb_prune_tile<4, 4, 3>, 3); }
        else if (lbq >= 2) go(k_lb_prune_tile<2, 4, 3>, 3);
        else k_lb_prune<<<148 * 8, 256, 0, st>>>(la);
      }
      c.mode = PM_LIST; c.list = list; c.list_len = list_len;
      rc = launch(0, nq, c0, nc, dbuf, C, nullptr, thr, nullptr);
      c.mode = PM_PAIRWISE; c.list = nullptr; c.list_len = nullptr;
      if (stats) stats->launches += 2;  // the LB pass, the DP
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
    ra.k = k; ra.kind = kind; ra.scale = scale; ra.tau = tau; ra.hidx = hidx; ra.hval = hval; ra.hn = hn; ra.amb = amb;
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
    if (!rc && stats && amb) {
      std::vector<int> h((size_t)nq);
      if (cudaMemcpy(h.data(), amb, sizeof(int) * nq, cudaMemcpyDeviceToHost) == cudaSuccess)
        for (long long q = 0; q < nq; ++q) stats->ambiguous += h[(size_t)q] != 0;
    }
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
  if (up_ev) {
    if (rc || staged) cudaStreamSynchronize(io.up_stream);  // nothing may still be copying from / into a buffer the caller is about to free
    cudaEventDestroy(up_ev);
  }
  for (int sb = 0; sb < 2; ++sb) if (stage_ev[sb]) cudaEventDestroy(stage_ev[sb]);
  return rc;
}

}  // namespace wb
