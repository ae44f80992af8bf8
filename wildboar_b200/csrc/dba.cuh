// DTW alignment paths and the DBA barycentre update on the device.
//
// Reference: `_dtw_alignment` (_elastic.pyx:1011-1073) fills a full (Tx, Ty) matrix per pair,
// `dtw_mapping` (distance/dtw.py:385-413) walks it back in Python, and `_mm_dtw_average`
// (distance/dtw.py:655-690) accumulates the barycentre with a Python loop over every path cell.
//
// B200-first formulation: the matrix is never materialised.  One thread per (a, b) pair runs the
// banded forward pass over ONE in-place band row in shared memory ([slot][thread], conflict free) and
// records per cell only the 2-bit MOVE the back-walk would take there (np.argmin([diag, up, left]): first minimum
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
  double* scratch;   // fallback for very tall bands: H + 1 slots per thread, slot s at scratch[s * sstride + gtid]
  long long sstride;
  unsigned char* moves;  // [task][row][HB][32] packed 2-bit moves
  int HB;                // bytes per row: ceil(H / 4)
  int* lo;           // (n_pairs, Ta) first column of the path in each row
  int* hi;           // (n_pairs, Ta) last column
  double* cost;      // optional (n_pairs): D[Ta-1][Tb-1]
  double* D;         // optional (n_pairs, Ta, Tb): the alignment matrix, +inf outside the band
};

// One band row lives IN PLACE in `row` (shared memory when it fits, else global scratch), indexed by the
// band-relative column jb = j - (i - a):  before row i is processed slot jb holds D[i-1][j-1] (the diagonal
// neighbour) and slot jb+1 holds D[i-1][j] (the upper neighbour); the cell overwrites slot jb.  Slots a row does
// not write are never written by an earlier row either (the band only moves right), so they still hold the +inf
// they were initialised with -- exactly the sentinels the reference plants around its band.  Per cell: one load,
// one store, no second scratch row.
template <bool SMEM>
__global__ void __launch_bounds__(128) k_dtw_paths(PathArgs p) {
  extern __shared__ double path_smem[];
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long long task = gtid >> 5;
  long long pair = gtid;
  const bool valid = pair < p.n_pairs;
  if (!valid) pair = p.n_pairs - 1;
  const int Ta = p.g.Tx, Tb = p.g.Ty;
  const double* __restrict__ x = p.a + (long long)(p.ia ? p.ia[pair] : pair) * Ta;
  const double* __restrict__ y = p.b + (long long)(p.ib ? p.ib[pair] : pair) * Tb;
  double* const row = SMEM ? path_smem + threadIdx.x : p.scratch + gtid;
  const long long rs = SMEM ? (long long)blockDim.x : p.sstride;
  unsigned char* mv = p.moves + (task * Ta * p.HB) * 32 + lane;
  double* Dp = p.D ? p.D + pair * (long long)Ta * Tb : nullptr;
  const double INF = WB_INF;
  const int H = p.g.H;
  for (int s = 0; s <= H; ++s) row[s * rs] = INF;

  double last = INF;
  for (int i = 0; i < Ta; ++i) {
    const int js = imax2(0, i - p.g.a), je = imin2(Tb, i + p.g.max_len);
    const double xi = x[i];
    const int jb0 = js - (i - p.g.a);  // band-relative index of the first cell
    double left = INF;
    double diag = (i == 0) ? 0.0 : row[jb0 * rs];
    unsigned pack = 0;
    unsigned char* mrow = mv + (long long)i * p.HB * 32;
    double* slot = row + (long long)jb0 * rs;
    if (Dp && valid) for (int j = 0; j < js; ++j) Dp[(long long)i * Tb + j] = INF;
    // the cell: reads its neighbours from registers, records the move, returns D[i][j]
    auto cell = [&](int j, double up, double yj) {
      const double v = xi - yj;
      double c = v * v;
      if (p.w) c = c * p.w[i - j];
      const double d = dmin2(dmin2(up, left), diag) + c;   // min(min(x, y), z) + v * v * w, EL:1063-1068
      // move of the back-walk at (i, j): argmin([diag, up, left]), first minimum (dtw.py:402-409)
      const unsigned move = (diag <= up && diag <= left) ? 0u : (up <= left ? 1u : 2u);
      const int jb = jb0 + (j - js);
      pack |= move << (2 * (jb & 3));
      if ((jb & 3) == 3) { mrow[(long long)(jb >> 2) * 32] = (unsigned char)pack; pack = 0; }
      if (Dp && valid) Dp[(long long)i * Tb + j] = d;
      diag = up;
      left = d;
      return d;
    };
    int j = js;
    // groups of four, software pipelined: the upper neighbours (old row, slots jb+1 .. jb+4) and the y samples of
    // the NEXT group are loaded before the current group's dependent min/add chain and stores (slots jb .. jb+3,
    // never the ones being prefetched), so the L1/L2 latency of y overlaps the arithmetic
    if (j + 4 <= je) {
      double u0 = slot[rs], u1 = slot[2 * rs], u2 = slot[3 * rs], u3 = slot[4 * rs];
      double y0 = y[j], y1 = y[j + 1], y2 = y[j + 2], y3 = y[j + 3];
      for (; j + 8 <= je; j += 4) {
        const double n0 = slot[5 * rs], n1 = slot[6 * rs], n2 = slot[7 * rs], n3 = slot[8 * rs];
        const double z0 = y[j + 4], z1 = y[j + 5], z2 = y[j + 6], z3 = y[j + 7];
        const double d0 = cell(j, u0, y0);
        const double d1 = cell(j + 1, u1, y1);
        const double d2 = cell(j + 2, u2, y2);
        const double d3 = cell(j + 3, u3, y3);
        slot[0] = d0; slot[rs] = d1; slot[2 * rs] = d2; slot[3 * rs] = d3;
        slot += 4 * rs;
        u0 = n0; u1 = n1; u2 = n2; u3 = n3;
        y0 = z0; y1 = z1; y2 = z2; y3 = z3;
      }
      const double d0 = cell(j, u0, y0);
      const double d1 = cell(j + 1, u1, y1);
      const double d2 = cell(j + 2, u2, y2);
      const double d3 = cell(j + 3, u3, y3);
      slot[0] = d0; slot[rs] = d1; slot[2 * rs] = d2; slot[3 * rs] = d3;
      slot += 4 * rs;
      j += 4;
    }
    for (; j < je; ++j) {
      const double d = cell(j, slot[rs], y[j]);
      *slot = d;
      slot += rs;
    }
    {
      const int jbl = jb0 + (je - js);  // one past the last cell
      if (jbl & 3) mrow[(long long)(jbl >> 2) * 32] = (unsigned char)pack;
    }
    if (Dp && valid) for (int j = je; j < Tb; ++j) Dp[(long long)i * Tb + j] = INF;
    last = left;  // D[i][je-1]; for the last row je == Tb
  }
  if (!valid) return;
  if (p.cost) p.cost[pair] = last;

  // back-walk (dtw.py:396-411); (0, 0) has no predecessor, its stored move is never read
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
  // members in batches of 8: indices, path ranges and the first path value of the whole batch are loaded before
  // the (order-preserving) accumulation, so eight dependent load chains are in flight instead of one
  constexpr int B = 8;
  const long long q1 = a.off[c + 1];
  for (long long q0 = a.off[c]; q0 < q1; q0 += B) {
    int ii[B], l[B], h[B];
    double v0[B];
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const long long q = q0 + k < q1 ? q0 + k : q1 - 1;
      ii[k] = a.member[q];
      l[k] = a.lo[q * a.Tm + m];
      h[k] = a.hi[q * a.Tm + m];
    }
#pragma unroll
    for (int k = 0; k < B; ++k) v0[k] = a.X[(long long)ii[k] * a.T + l[k]];
#pragma unroll
    for (int k = 0; k < B; ++k) {
      if (q0 + k < q1) {
        const double w = a.sw ? a.sw[ii[k]] : 1.0;
        const double* xs = a.X + (long long)ii[k] * a.T;
        V += w; z += v0[k] * w;
        for (int x = l[k] + 1; x <= h[k]; ++x) { V += w; z += xs[x] * w; }
      }
    }
  }
  a.mean_out[e] = z / V;
}

}  // namespace wb
