// Band-register banded-DP engine: one thread per pair, the previous DP row of the Sakoe-Chiba band in REGISTERS.
//
// For equal-length pairs the band of row i is the H = 2R - 1 columns j = i - a + k, k = 0 .. H-1 (a = R - 1): in the band
// coordinate k the three neighbours of a cell are
//     up   = D[i-1][j]   = P[k + 1]      diag = D[i-1][j-1] = P[k]      left = D[i][j-1] = the cell just computed,
// so ONE register array P[HB] (HB >= H, compile time) holds the previous row and is updated in place, left to right: a
// cell overwrites exactly the entry it has just consumed as its diagonal.  No scratch rows in memory (the row-scan engine
// moves 16 bytes per cell through L1/L2 and is bound by that), rows are still visited top to bottom, so the row minima the
// reference's early abandoning tests (EL:933, 1175, 1339, 1489, 1639, 1821) are available: this engine serves every caller
// that needs them -- argmin and the subsequence scans for the metrics whose abandoning is not monotone -- whenever the band
// is narrow (H <= 32); wide bands and unequal lengths stay on the row-scan engine, results without row minima on the strip
// engine.
//
// Everything the reference's two scratch rows hold outside the band is reproduced from the same policy constants the
// row-scan engine uses (engine_rowscan.cuh is the executable specification; tests/hostsim runs both on the host and
// compares values AND row-minimum maxima): `prev_init` above row 0, `usent` in the cell one past the previous row's band,
// `lsent` left of the band, `left0/diag0` at column 0; for MSM the always-evaluated column 0 (running sum `cy`), the
// explicit row 0 with its one cell beyond the band, and the never-reset left edge (the value that sat in that scratch
// cell two rows earlier = band coordinate 1 of row i-2).
#pragma once
#include "engine_rowscan.cuh"

namespace wb {

template <class M>
inline bool band_supported(const Geom& g, int HB) {
  if (g.Tx != g.Ty || g.Tx < 2) return false;
  if (g.H > HB || g.H != 2 * g.max_len - 1) return false;
  if (M::kMsmBand && g.H < 3) return false;  // the stale-left rule needs band coordinate 1
  return true;
}

// min_dist: abandon when a checked row's minimum exceeds it (raw dp domain); +inf disables.
// row_min_max (optional): max over checked rows of the row minimum (for the exact replay).
// YS: element stride of y (1: a plain series; 32: interleaved in groups of 32 series, kernels.cuh `KArgs::yil`).
template <class M, int HB, int YS = 1, bool BLK = false>
WB_HD typename M::real band_pair(const Geom& g, const M& m, const typename M::real* __restrict__ x,
                                 const typename M::real* __restrict__ yp, typename M::real min_dist,
                                 typename M::real* row_min_max) {
  using F = typename M::real;
  struct YView {
    const F* p;
    WB_HD F operator[](int t) const { return p[(long long)t * YS]; }
    WB_HD YView operator+(int t) const { return YView{p + (long long)t * YS}; }
  };
  const YView y{yp};
  const int T = g.Tx, a = g.a, H = g.H, R = g.max_len;
  F P[HB];
#pragma unroll
  for (int k = 0; k < HB; ++k) P[k] = m.prev_init();
  F mmax = -Num<F>::inf();
  F cy = F(0), stale = m.lsent(), beyond = m.usent();
  int i_first = 0;

  if (M::kMsmBand) {
    // explicit first row incl. the one cell beyond the band (EL:1611-1617): D[0][j] at band coordinate j + a
    const typename M::Row r0 = m.row(0, x[0], F(0));
    F v = F(0);
#pragma unroll
    for (int k = 0; k < HB; ++k) {
      const int j = k - a;
      if (j >= 0 && j < imin2(T, R)) {
        const typename M::Col cj = m.col(j, y[j], j > 0 ? y[j - 1] : F(0));
        v = (j == 0) ? m.cell(Num<F>::inf(), Num<F>::inf(), F(0), r0, cj, m.dv(0, 0))
                     : m.cell(Num<F>::inf(), v, Num<F>::inf(), r0, cj, m.dv(0, j));
        P[k] = v;
        if (j == 0) cy = v;
      }
    }
    if (R < T) {
      const typename M::Col cj = m.col(R, y[R], y[R - 1]);
      beyond = m.cell(Num<F>::inf(), v, Num<F>::inf(), r0, cj, m.dv(0, R));
    }
    i_first = 1;
  }

  // ---- blocked interior rows (HB <= 16) ----
  // Rows a + 1 (MSM: a + 2) .. T - R have their whole band inside the matrix and no row-0 / column-0 rule: NRB of them run
  // per iteration without band-edge predicates, with the column contexts of the HB + NRB - 1 columns they touch held in
  // registers (cell k of row r reads cols[1 + k + r]; after the block the array moves down by NRB and NRB new columns
  // enter) and the per-diagonal values as loop invariants -- the generic row below re-derives a column context from two
  // loads per cell and tests four predicates per cell (ncu, profiles/r01m_ncu_band.md: 50 instructions per msm cell).
  // (a separate instantiation: the blocked rows need ~2x the registers.  HB = 32 was tried for the policies whose column
  // context is a single value (lcss, edr), two rows per block: 1.6x faster on 400 x 400 x 128, 1.4x SLOWER on
  // 1000 x 5000 x 140 -- profiles/r02v_band_blocked_hb32_light.jsonl -- and is not used)
  constexpr bool kBlocked = BLK && HB <= 16;
  constexpr int NRB = 4;
  constexpr int NCB = kBlocked ? HB + NRB : 1;
  typename M::Col cols[NCB];
  typename M::Dv dvs[kBlocked ? HB : 1];
  const int r_lo = M::kMsmBand ? a + 2 : a + 1, r_hi = T - R;
  bool cols_ready = false;
  const typename M::Col c0col = m.col(0, y[0], F(0));

  for (int i = i_first; i < T; ++i) {
    if (kBlocked && i >= r_lo && i + NRB - 1 <= r_hi) {
      if (!cols_ready) {
#pragma unroll
        for (int q = 0; q < NCB; ++q) {
          const int j = imin2(imax2(i - a - 1 + q, 0), T - 1);
          cols[q] = m.col(j, y[j], j > 0 ? y[j - 1] : F(0));
        }
        if (M::kHasDv) {
#pragma unroll
          for (int k = 0; k < HB; ++k) dvs[k < (kBlocked ? HB : 1) ? k : 0] = m.dv_diag(a - k);
        }
        cols_ready = true;
      }
      // samples of the NRB columns that enter after this block
      F ynew[NRB];
#pragma unroll
      for (int u = 0; u < NRB; ++u) ynew[u] = y[imin2(i + NRB - a - 1 + HB + u, T - 1)];
      F xprev = x[i - 1];
#pragma unroll
      for (int r = 0; r < NRB; ++r) {
        const F xr = x[i + r];
        const typename M::Row rw = m.row(i + r, xr, xprev);
        xprev = xr;
        F rowmin = Num<F>::inf();
        F left = m.lsent();
        F stale_next = stale;
        if (M::kMsmBand) {
          cy = m.cell(cy, Num<F>::inf(), Num<F>::inf(), rw, c0col, m.dv(i + r, 0));  // column 0: always evaluated, part of the row minimum
          rowmin = cy;
          stale_next = P[1];
          left = stale;
        }
#pragma unroll
        for (int k = 0; k < HB; ++k) {
          if (k < H) {
            const F up = (k + 1 < H) ? P[(k + 1 < HB) ? (k + 1) : (HB - 1)] : m.usent();
            typename M::Col cj = cols[(1 + k + r) < NCB ? (1 + k + r) : 0];
            ColLink<M>::apply(cj, cols[(k + r) < NCB ? (k + r) : 0], true);
            const F d = m.cell(up, left, P[k], rw, cj, dvs[k < (kBlocked ? HB : 1) ? k : 0]);
            P[k] = d;
            rowmin = dmin2(rowmin, d);
            left = d;
          }
        }
        stale = stale_next;
        mmax = dmax2(mmax, rowmin);
        if (rowmin > min_dist) { if (row_min_max) *row_min_max = mmax; return Num<F>::inf(); }
      }
#pragma unroll
      for (int q = 0; q + NRB < NCB; ++q) cols[q] = cols[q + NRB];
#pragma unroll
      for (int u = 0; u < NRB; ++u) {
        const int q = HB + u;  // NCB - NRB + u
        const int j = imin2(i + NRB - a - 1 + q, T - 1);
        cols[q < NCB ? q : 0] = m.col(j, ynew[u], cols[(q - 1) < NCB ? (q - 1) : 0].yj);
      }
      i += NRB - 1;
      continue;
    }
    cols_ready = false;
    const F xi = x[i];
    const F xim = (i > 0) ? x[i - 1] : F(0);
    const typename M::Row rw = m.row(i, xi, xim);
    const YView yb = y + (i - a);              // y[j] = yb[k]
    int klo = imax2(0, a - i);                  // first band coordinate inside the matrix (column 0 when i <= a)
    const int khi = imin2(H, T - i + a);        // one past the last
    // what the previous row left one cell past its band: prev_init above row 0, the sentinel otherwise (MSM row 1: the
    // explicit row's extra cell)
    const F uplast = (i == 0) ? m.prev_init() : ((M::kMsmBand && i == 1) ? beyond : m.usent());
    F rowmin = Num<F>::inf();
    F left;
    F stale_next = stale;
    if (M::kMsmBand) {
      const typename M::Col c0 = m.col(0, y[0], F(0));
      cy = m.cell(cy, Num<F>::inf(), Num<F>::inf(), rw, c0, m.dv(i, 0));  // column 0: up-branch only (EL:1620-1628)
      rowmin = cy;
      stale_next = P[1];                        // band coordinate 1 of row i-1: the stale left edge of row i+1
      left = (i <= a + 1) ? cy : stale;         // cost[js-1]: column 0 (= cy) or never reset by the reference
      if (i <= a) {
        // column 0 sits at coordinate klo: it receives cy (next row's diagonal / this row's left) and is not a DP cell
#pragma unroll
        for (int k = 0; k < HB; ++k) if (k == klo) P[k] = cy;
        klo += 1;
      }
    } else {
      left = (i <= a) ? m.left0(i) : m.lsent();
    }
    const bool col0_first = !M::kMsmBand && i <= a;  // the first cell is column 0: its diagonal is diag0(i)
#pragma unroll
    for (int k = 0; k < HB; ++k) {
      if (k >= klo && k < khi) {
        const int j = i - a + k;
        const F up = (k + 1 < H) ? P[(k + 1 < HB) ? (k + 1) : (HB - 1)] : uplast;  // H <= HB: the clamp is never taken
        const F diag = (col0_first && k == klo) ? m.diag0(i) : P[k];
        const F yj = yb[k];
        const F yjm = (j > 0) ? yb[k - 1] : F(0);
        const typename M::Col cj = m.col(j, yj, yjm);
        const F d = m.cell(up, left, diag, rw, cj, m.dv(i, j));
        P[k] = d;
        rowmin = dmin2(rowmin, d);
        left = d;
      }
    }
    stale = stale_next;
    if (!(EaFromRow1<M>::value && i == 0)) {
      mmax = dmax2(mmax, rowmin);
      if (rowmin > min_dist) { if (row_min_max) *row_min_max = mmax; return Num<F>::inf(); }
    }
  }
  if (row_min_max) *row_min_max = mmax;
  // D[T-1][T-1] sits at band coordinate a of the last row
  F result = F(0);
#pragma unroll
  for (int k = 0; k < HB; ++k) if (k == a) result = P[k];
  return m.finish(result, g);
}

}  // namespace wb
