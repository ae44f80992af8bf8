// Row-scan banded-DP engine: one thread per pair, two scratch rows in global memory
// (interleaved across threads so a warp's row accesses coalesce).
//
// This is the GENERAL engine: it keeps the reference's two alternating scratch rows as real
// buffers, so every quirk that depends on buffer contents (sentinels outside the band, MSM's
// stale left edge, EL:1592-1647) falls out of the data flow for ANY geometry (T = 1, R = 1,
// unequal lengths ...).  It also tracks the per-row minima that the reference's early
// abandoning tests (EL:933, 1175, 1339, 1489, 1639, 1821), which argmin needs for an exact
// replay of the sequential scan (CD:1302-1345).  The strip engine (engine_strip.cuh) is the
// fast path; this one is the fallback for geometries it does not cover and the engine behind
// argmin for the non-DTW metrics.
#pragma once
#include "metrics.cuh"

namespace wb {

// rows whose minimum takes part in early abandoning: dtw family + msm check rows >= 1 only
template <class M> struct EaFromRow1 { static constexpr bool value = M::kColumnMinBound || M::kMsmBand; };

// b0/b1: scratch rows, element j at b[j * bs]; each needs max(Tx,Ty)+1 elements.
// min_dist: abandon when a checked row's minimum exceeds it (raw dp domain); WB_INF disables.
// row_min_max (optional): max over checked rows of the row minimum (for replay).
// YS: element stride of y (1: a plain series; 32: series interleaved in groups of 32 so that a warp's lanes read
// consecutive addresses, kernels.cuh `KArgs::yil`)
template <class M, int YS = 1>
WB_HD typename M::real rowscan_pair(const Geom& g, const M& m, const typename M::real* __restrict__ x,
                                    const typename M::real* __restrict__ yp, typename M::real* b0, typename M::real* b1,
                                    long long bs, typename M::real min_dist, typename M::real* row_min_max) {
  using F = typename M::real;
  struct YView { const F* p; WB_HD F operator[](int t) const { return p[(long long)t * YS]; } };
  const YView y{yp};
  const int Tx = g.Tx, Ty = g.Ty;
  F* prev = b0;
  F* cost = b1;
  F mmax = -Num<F>::inf();
  int i_first = 0;
  F cy = F(0);

  if (M::kMsmBand) {
    // explicit first row incl. the one cell beyond the band (EL:1611-1617); first column is
    // the running sum cy (EL:1620-1622)
    typename M::Row r0 = m.row(0, x[0], F(0));
    typename M::Col c0 = m.col(0, y[0], F(0));
    F v = m.cell(Num<F>::inf(), Num<F>::inf(), F(0), r0, c0, m.dv(0, 0));
    prev[0] = v;
    int n0 = imin2(Ty, g.max_len + 1);
    for (int j = 1; j < n0; ++j) {
      typename M::Col cj = m.col(j, y[j], y[j - 1]);
      v = m.cell(Num<F>::inf(), v, Num<F>::inf(), r0, cj, m.dv(0, j));
      prev[(long long)j * bs] = v;
    }
    cy = prev[0];
    i_first = 1;
  } else {
    int n0 = imin2(Ty, g.max_len);
    for (int j = 0; j < n0; ++j) prev[(long long)j * bs] = m.prev_init();
    if (g.max_len < Ty) prev[(long long)g.max_len * bs] = m.prev_init();
  }

  for (int i = i_first; i < Tx; ++i) {
    int js = imax2(0, i - g.a);
    const int je = imin2(Ty, i + g.max_len);
    const F xi = x[i];
    const F xim = (i > 0) ? x[i - 1] : F(0);
    const typename M::Row rw = m.row(i, xi, xim);
    F rowmin = Num<F>::inf();
    F left, diag;
    if (M::kMsmBand) {
      typename M::Col c0 = m.col(0, y[0], F(0));
      cy = m.cell(cy, Num<F>::inf(), Num<F>::inf(), rw, c0, m.dv(i, 0));  // up-branch only: cy[i-1] + cost(X[i],X[i-1],Y[0])
      cost[0] = cy;
      rowmin = cy;
      js = imax2(1, js);
      left = cost[(long long)(js - 1) * bs];   // NOT reset by the reference: stale or cy
      diag = prev[(long long)(js - 1) * bs];
    } else {
      if (js > 0) {
        cost[(long long)(js - 1) * bs] = m.lsent();
        left = m.lsent();
        diag = prev[(long long)(js - 1) * bs];
      } else {
        left = m.left0(i);
        diag = m.diag0(i);
      }
    }
    for (int j = js; j < je; ++j) {
      const F up = prev[(long long)j * bs];
      const F yj = y[j];
      const F yjm = (j > 0) ? y[j - 1] : F(0);
      const typename M::Col cj = m.col(j, yj, yjm);
      const F d = m.cell(up, left, diag, rw, cj, m.dv(i, j));
      cost[(long long)j * bs] = d;
      rowmin = dmin2(rowmin, d);
      left = d;
      diag = up;
    }
    if (!(EaFromRow1<M>::value && i == 0)) {
      mmax = dmax2(mmax, rowmin);
      if (rowmin > min_dist) { if (row_min_max) *row_min_max = mmax; return Num<F>::inf(); }
    }
    if (je < Ty) cost[(long long)je * bs] = m.usent();
    F* t = cost; cost = prev; prev = t;
  }
  if (row_min_max) *row_min_max = mmax;
  return m.finish(prev[(long long)(Ty - 1) * bs], g);
}

}  // namespace wb
