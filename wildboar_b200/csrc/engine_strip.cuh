// Column-strip banded-DP engine: ONE THREAD PER (x_i, y_j) PAIR.
//
// Design (B200-first, not the reference's two-row scan, EL:869-940):
//   * All pairs of one call share (Tx, Ty, R), so every thread of a warp walks the identical
//     band geometry: zero divergence, no shuffles.
//   * The band is cut into strips of W consecutive columns.  Inside a strip the thread keeps
//     the previous DP row of the strip (W doubles) and the strip's y values in REGISTERS and
//     sweeps the strip's rows top to bottom; the only memory traffic per row is one read of
//     the left neighbour column D[i][j0-1] and one write of the strip's last column
//     D[i][j0+W-1] -- a ring of H+1 doubles per pair ("boundary column") that lives in shared
//     memory, laid out [slot][lane] so a warp's accesses are conflict free.
//   * Per row of W cells: 1 LDS (boundary in) + 1 STS (boundary out) + 1 load of x[i]
//     (same address for the whole warp in pairwise mode) => the FP64 pipe, not the LSU, is the
//     limiter.  ILP comes from the independent min(up, diag) / cost terms of the W cells.
//
// The function is __host__ __device__: tests/hostsim compiles the very same code with g++
// and checks it bit-for-bit against the oracle without a GPU (test infrastructure only; the
// product library exposes no CPU path).
#pragma once
#include "metrics.cuh"

namespace wb {

template <class M>
WB_HD int row_js(const Geom& g, int i) {
  int v = i - g.a;
  if (M::kMsmBand) return v <= 1 ? 0 : v;  // column 0 is always evaluated (EL:1625-1628)
  return v > 0 ? v : 0;
}
template <class M>
WB_HD int row_je(const Geom& g, int i) {
  int v = i + g.max_len;
  if (M::kMsmBand && i == 0) v += 1;  // row 0 fills one cell beyond the band (EL:1615-1617)
  return v < g.Ty ? v : g.Ty;
}

// Can the strip engine reproduce the reference for this geometry?  (Everything else goes to
// the row-scan engine, engine_rowscan.cuh.)
template <class M>
inline bool strip_supported(const Geom& g, int W) {
  if (g.Tx < 2 || g.Ty < 2 || W < 2) return false;
  if (M::kMsmBand && g.H < 3) return false;  // stale-left rule needs a genuine cell two rows up
  return true;
}

// Ring slots needed per pair: every in-band cell of a boundary column (H of them), plus the
// left-sentinel slot that the producer strip appends below it.
WB_HD int strip_ring_slots(const Geom& g) { return g.H + 1; }

// bnd: ring base for this pair; slot s lives at bnd[s * bs].
// abandon: early-abandon threshold on the RAW dp value (pre-finish); only honoured for
//          policies whose column minima lower-bound the result (DTW family).  Returns +INF
//          when abandoned.
template <class M, int W>
WB_HD double strip_pair(const Geom& g, const M& m, const double* __restrict__ x,
                        const double* __restrict__ y, double* bnd, int bs, int NS, double abandon) {
  const int Tx = g.Tx, Ty = g.Ty;
  double result = 0.0;

  for (int j0 = 0; j0 < Ty; j0 += W) {
    const int wv = imin2(W, Ty - j0);
    const bool last_strip = (j0 + W >= Ty);
    const int jl = j0 + wv - 1;
    int i_lo = imax2(0, j0 - g.max_len + 1);
    if (M::kMsmBand && j0 == g.max_len) i_lo = 0;
    const int i_hi = imin2(Tx - 1, jl + g.a);

    typename M::Col cols[W];
#pragma unroll
    for (int c = 0; c < W; ++c) {
      int jj = imin2(j0 + c, Ty - 1);
      double yj = y[jj];
      double yjm = (jj > 0) ? y[jj - 1] : 0.0;
      cols[c] = m.col(jj, yj, yjm);
    }
    double prev[W];
#pragma unroll
    for (int c = 0; c < W; ++c) prev[c] = m.usent();

    int sl = i_lo % NS;
    double Dg = 0.0;
    if (j0 > 0) {
      if (i_lo == 0) Dg = m.prev_init();
      else { int sp = (sl == 0) ? NS - 1 : sl - 1; Dg = bnd[sp * bs]; }
    }
    double xi = x[i_lo];
    double xim = (M::kNeedPrevX && i_lo > 0) ? x[i_lo - 1] : 0.0;
    double stale = m.lsent();
    double colmin = WB_INF;

    for (int i = i_lo; i <= i_hi; ++i) {
      const double xnext = x[imin2(i + 1, Tx - 1)];  // prefetch next row's sample
      const int js = row_js<M>(g, i), je = row_je<M>(g, i);
      const int clo = imax2(js - j0, 0), chi = imin2(je - j0, wv);
      double left, diag;
      if (j0 == 0) { left = m.left0(i); diag = m.diag0(i); }
      else { left = bnd[sl * bs]; diag = Dg; Dg = left; }
      if (i == 0) {
#pragma unroll
        for (int c = 0; c < W; ++c) if (c < chi) prev[c] = m.prev_init();
      }
      const typename M::Row rw = m.row(i, xi, xim);
      const int cst = M::kMsmBand ? (row_js<M>(g, i + 1) - 1 - j0) : -1;
      double stale_next = stale;
#pragma unroll
      for (int c = 0; c < W; ++c) {
        const double up = prev[c];
        if (M::kMsmBand && c == cst) stale_next = up;
        if (c >= clo && c < chi) {
          if (c > 0 && c == clo) left = M::kMsmBand ? stale : m.lsent();
          const double d = m.cell(up, left, diag, rw, cols[c], i, j0 + c);
          prev[c] = d;
          left = d;
        }
        diag = up;
      }
      stale = stale_next;
      if (!last_strip && chi == W) {
        const double b = prev[W - 1];
        bnd[sl * bs] = b;
        if (M::kColumnMinBound) colmin = dmin2(colmin, b);
      }
      sl = (sl + 1 == NS) ? 0 : sl + 1;
      xim = xi;
      xi = xnext;
    }
    if (!last_strip) {
      if (jl + g.a + 1 <= Tx - 1) bnd[sl * bs] = M::kMsmBand ? stale : m.lsent();
      if (M::kColumnMinBound && colmin > abandon) return WB_INF;
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c) if (c == wv - 1) result = prev[c];
    }
  }
  return m.finish(result, g);
}

}  // namespace wb
