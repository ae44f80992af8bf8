// DTW alignment paths and the DBA barycentre update on the device.
//
// Reference: `_dtw_alignment` (_elastic.pyx:1011-1073) fills a full (Tx, Ty) matrix per pair,
// `dtw_mapping` (distance/dtw.py:385-413) walks it back in Python, and `_mm_dtw_average`
// (distance/dtw.py:655-690) accumulates the barycentre with a Python loop over every path cell.
//
// B200-first formulation: the matrix is never materialised.  One thread per (a, b) pair runs the
// banded forward pass with two scratch rows (lane-interleaved, like the row-scan engine) and records per
// cell only the 2-bit MOVE the back-walk would take there (np.argmin([diag, up, left]): first minimum
// wins), packed four to a byte in [task][row][byte][lane] order so a warp writes 32 contiguous bytes.
// The same thread then walks back and emits the path as one column range [lo, hi] per row (a monotone
// path covers a contiguous run of columns in every row; `indicator.nonzero()` enumerates exactly these
// cells row by row, ascending).  The barycentre update is a second kernel with one thread per
// (cluster, time step) that adds the path cells in the reference's order (samples ascending, columns
// ascending), so the new centres are bit-equal to the reference's.
#pragma once
#include <cuda_runtime.h>
#include "metrics.cuh"

namespace wb {

struct PathArgs {
  const double* a;   // (na, Ta) dense: first operand (rows of the DP)
  const double* b;   // (nb, Tb) dense: second operand (columns)
  const int* ia;     // per pair: row of `a` (nullptr: pair index)
  const int* ib;     // per pair: row of `b` (nullptr: pair index)
  long long n_pairs;
  Geom g;            // Tx = Ta, Ty = Tb, R = max(floor(max(Ta, Tb) r), 1)  (dtw.py:38-40)
  const double* w;   // centre of the signed weight table w[d] = weight[|d|], or nullptr
  double* scratch;   // 2 rows of (Tb + 1) per thread, element j at scratch[j * sstride + gtid]
  long long sstride;
  unsigned char* moves;  // [task][row][HB][32] packed 2-bit moves
  int HB;                // bytes per row: ceil(H / 4)
  int* lo;           // (n_pairs, Ta) first column of the path in each row
  int* hi;           // (n_pairs, Ta) last column
  double* cost;      // optional (n_pairs): D[Ta-1][Tb-1]
  double* D;         // optional (n_pairs, Ta, Tb): the alignment matrix, +inf outside the band
};

// launch: grid * block >= n_pairs rounded up to 32; `task0` = first warp task of this launch
__global__ void __launch_bounds__(128) k_dtw_paths(PathArgs p) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long long task = gtid >> 5;
  long long pair = gtid;
  const bool valid = pair < p.n_pairs;
  if (!valid) pair = p.n_pairs - 1;
  const int Ta = p.g.Tx, Tb = p.g.Ty;
  const double* x = p.a + (long long)(p.ia ? p.ia[pair] : pair) * Ta;
  const double* y = p.b + (long long)(p.ib ? p.ib[pair] : pair) * Tb;
  double* prev = p.scratch + gtid;
  double* cur = prev + (long long)(Tb + 1) * p.sstride;
  const long long ss = p.sstride;
  unsigned char* mv = p.moves + (task * Ta * p.HB) * 32 + lane;
  double* Dp = p.D ? p.D + pair * (long long)Ta * Tb : nullptr;
  const double INF = WB_INF;

  int pjs = 0, pje = 0;  // band of the previous row
  for (int i = 0; i < Ta; ++i) {
    const int js = imax2(0, i - p.g.a), je = imin2(Tb, i + p.g.max_len);
    const double xi = x[i];
    double left = INF;
    double diag = (i == 0) ? 0.0 : ((js > 0 && js - 1 >= pjs) ? prev[(long long)(js - 1) * ss] : INF);
    unsigned pack = 0;
    unsigned char* mrow = mv + (long long)i * p.HB * 32;
    const int jb0 = js - (i - p.g.a);  // band-relative index of the first cell
    if (Dp && valid) for (int j = 0; j < js; ++j) Dp[(long long)i * Tb + j] = INF;
    for (int j = js; j < je; ++j) {
      const double up = (i > 0 && j < pje) ? prev[(long long)j * ss] : INF;
      const double v = xi - y[j];
      double c = v * v;
      if (p.w) c = c * p.w[i - j];
      const double d = dmin2(dmin2(up, left), diag) + c;   // min(min(x, y), z) + v * v * w, EL:1063-1068
      // move of the back-walk at (i, j): argmin([diag, up, left]), first minimum (dtw.py:402-409)
      const unsigned move = (diag <= up && diag <= left) ? 0u : (up <= left ? 1u : 2u);
      const int jb = jb0 + (j - js);
      pack |= move << (2 * (jb & 3));
      if ((jb & 3) == 3) { mrow[(long long)(jb >> 2) * 32] = (unsigned char)pack; pack = 0; }
      cur[(long long)j * ss] = d;
      if (Dp && valid) Dp[(long long)i * Tb + j] = d;
      diag = up;
      left = d;
    }
    {
      const int jbl = jb0 + (je - js);  // one past the last cell
      if (jbl & 3) mrow[(long long)(jbl >> 2) * 32] = (unsigned char)pack;
    }
    if (Dp && valid) for (int j = je; j < Tb; ++j) Dp[(long long)i * Tb + j] = INF;
    // (0, 0) has no predecessor; its stored move is never read
    double* t = prev; prev = cur; cur = t;
    pjs = js; pje = je;
  }
  if (!valid) return;
  if (p.cost) p.cost[pair] = prev[(long long)(Tb - 1) * ss];

  // back-walk (dtw.py:396-411)
  int* lo = p.lo + pair * Ta;
  int* hi = p.hi + pair * Ta;
  int i = Ta - 1, j = Tb - 1;
  hi[i] = j;
  while (i > 0 || j > 0) {
    lo[i] = j;
    const int jb = j - (i - p.g.a);
    const unsigned byte = mv[((long long)i * p.HB + (jb >> 2)) * 32];
    const unsigned move = (byte >> (2 * (jb & 3))) & 3u;
    if (move == 0) { --i; --j; hi[i] = j; }
    else if (move == 1) { --i; hi[i] = j; }
    else --j;
  }
  lo[0] = 0;
}

// One majorize-minimize step (dtw.py:666-680) for all clusters at once.  Thread (c, m): over the members
// q of cluster c (ascending sample index) and the path columns of row m (ascending):
//   V += w;  z += X[i][x] * w;        new_mean[c][m] = z / V
struct DbaArgs {
  const double* X;      // (n, T) dense samples
  int T;
  const int* member;    // (n_m) sample index of member q (cluster by cluster, ascending inside a cluster)
  const long long* off; // (K + 1) member ranges
  const double* sw;     // optional per-SAMPLE weights (indexed by sample index)
  const int* lo;        // (n_m, Tm)
  const int* hi;
  int K, Tm;
  double* mean_out;     // (K, Tm)
};

__global__ void __launch_bounds__(128) k_dba_update(DbaArgs a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)a.K * a.Tm) return;
  const int c = (int)(e / a.Tm), m = (int)(e - (long long)c * a.Tm);
  double z = 0.0, V = 0.0;
  for (long long q = a.off[c]; q < a.off[c + 1]; ++q) {
    const int i = a.member[q];
    const double w = a.sw ? a.sw[i] : 1.0;
    const int l = a.lo[q * a.Tm + m], h = a.hi[q * a.Tm + m];
    const double* xs = a.X + (long long)i * a.T;
    for (int x = l; x <= h; ++x) { V += w; z += xs[x] * w; }
  }
  a.mean_out[e] = z / V;
}

}  // namespace wb
