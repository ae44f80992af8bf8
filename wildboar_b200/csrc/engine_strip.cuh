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
// left-sentinel slot that the producer strip appends below it; never more than the Tx rows.
WB_HD int strip_ring_slots(const Geom& g) { return imax2(2, imin2(g.H + 1, g.Tx)); }

// Opaque select: keeps the compiler from turning a chain of register selects into a
// dynamically indexed (local-memory) array access.
WB_HD double sel_f64(bool p, double a, double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
  return r;
#else
  return p ? a : b;
#endif
}

// bnd: ring base for this pair; slot s lives at bnd[s * bs].
// abandon: early-abandon threshold on the RAW dp value (pre-finish); only honoured for
//          policies whose column minima lower-bound the result (DTW family).  Returns +INF
//          when abandoned.
//
// Rows of a strip fall in three phases: a top triangle (the band's upper edge crosses the
// strip), FULL rows (all W cells in band) and a bottom triangle.  Triangle rows go through
// `generic_row` (per-cell, warp-uniform predicates).  Full rows -- (2R-1-W+1)/(2R-1+W-1) of
// all rows, 87 % for the headline shape -- take the predicate-free `fast path`, two rows per
// iteration so that two independent left->right dependency chains are in flight per thread.
template <class M, int W, bool EA, int NR = 2>
WB_HD double strip_pair(const Geom& g, const M& m, const double* __restrict__ x,
                        const double* __restrict__ y, double* bnd, int bs, int NS, double abandon) {
  const int Tx = g.Tx, Ty = g.Ty;
  double result = 0.0;
  const double left0c = m.left0(1);

  for (int j0 = 0; j0 < Ty; j0 += W) {
    const int wv = imin2(W, Ty - j0);
    const bool last_strip = (j0 + W >= Ty);
    const bool first_strip = (j0 == 0);
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
    double Dg;
    if (first_strip) {
      // virtual column -1: left0 is the same constant for every row and diag0(i) == left0(i-1)
      // for i >= 1, so the first strip simply finds left0 in every ring slot it will read.
      const int nfill = imin2(NS, i_hi + 2);
      for (int s = 0; s < nfill; ++s) bnd[s * bs] = left0c;
      Dg = m.diag0(0);
    } else if (i_lo == 0) Dg = m.prev_init();
    else { int sp = (sl == 0) ? NS - 1 : sl - 1; Dg = bnd[sp * bs]; }
    double xi = x[i_lo];
    double xim = (M::kNeedPrevX && i_lo > 0) ? x[i_lo - 1] : 0.0;
    double stale = m.lsent();
    double colmin = WB_INF;
    int i = i_lo;
    double pf[NR];
    bool have_pf = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) pf[r] = 0.0;

    auto generic_row = [&]() {
      const double xnext = x[imin2(i + 1, Tx - 1)];  // prefetch next row's sample
      const int js = row_js<M>(g, i), je = row_je<M>(g, i);
      const int clo = imax2(js - j0, 0), chi = imin2(je - j0, wv);
      double left = bnd[sl * bs];
      double diag = Dg;
      Dg = left;
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
      if (chi == W) {
        const double b = prev[W - 1];
        bnd[sl * bs] = b;
        if (EA && M::kColumnMinBound) colmin = dmin2(colmin, b);
      }
      sl = (sl + 1 == NS) ? 0 : sl + 1;
      xim = xi;
      xi = xnext;
      ++i;
    };

    // full rows: js(i) <= j0 and je(i) >= jl + 1, i >= 1 (row 0 has its own init rules)
    const int fa = imax2(imax2(i_lo, 1), jl + 1 - g.max_len);
    int fb = imin2(i_hi, j0 + g.a);
    if (M::kMsmBand) fb -= 1;  // the generic row captures the stale-left value for the bottom triangle

    // Regular triangles (band edges cross the strip away from the matrix borders) are run as
    // straight-line code with exactly the in-band cells of every row: no per-cell predicates.
    const bool wide = !M::kMsmBand && wv == W && g.H >= 2 * W;
    const bool reg_top = wide && i_lo >= 1 && i_lo == j0 - g.max_len + 1;
    const bool reg_bot = wide && (j0 + g.a + W - 1 <= Tx - 1);

    // One copy of every row routine; the hot two-row loop stays a tight inner loop.
    for (;;) {
      while (i >= fa && i + NR - 1 <= fb) {
        // ---- fast path: rows i .. i+NR-1, all W cells in band, no predicates; NR independent
        // left->right dependency chains (an in-thread anti-diagonal wavefront) ----
        double xr[NR];
        xr[0] = xi;
#pragma unroll
        for (int r = 1; r < NR; ++r) xr[r] = x[i + r];
        const double xnext = x[imin2(i + NR, Tx - 1)];
        int slr[NR];
        slr[0] = sl;
#pragma unroll
        for (int r = 1; r < NR; ++r) slr[r] = (slr[r - 1] + 1 == NS) ? 0 : slr[r - 1] + 1;
        double lft[NR], dg[NR];
        typename M::Row rws[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          lft[r] = have_pf ? pf[r] : bnd[slr[r] * bs];
          rws[r] = m.row(i + r, xr[r], r == 0 ? xim : xr[r - 1]);
        }
        // prefetch the boundary values of the NEXT NR rows (hides the ring's load latency when
        // it lives in global memory; any slot is valid memory, unused values are discarded)
        {
          int sn = (slr[NR - 1] + 1 == NS) ? 0 : slr[NR - 1] + 1;
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            pf[r] = bnd[sn * bs];
            sn = (sn + 1 == NS) ? 0 : sn + 1;
          }
          have_pf = true;
        }
        dg[0] = Dg;
#pragma unroll
        for (int r = 1; r < NR; ++r) dg[r] = lft[r - 1];
        Dg = lft[NR - 1];
#pragma unroll
        for (int c = 0; c < W; ++c) {
          double v = prev[c];
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const double d = m.cell(v, lft[r], dg[r], rws[r], cols[c], i + r, j0 + c);
            dg[r] = v;
            lft[r] = d;
            v = d;
          }
          prev[c] = v;
        }
        // lft[r] now holds column W-1 of row i+r (the last strip's writes are never read; they
        // are kept so the hot loop has no strip-dependent branch)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          bnd[slr[r] * bs] = lft[r];
          if (EA && M::kColumnMinBound) colmin = dmin2(colmin, lft[r]);
        }
        sl = (slr[NR - 1] + 1 == NS) ? 0 : slr[NR - 1] + 1;
        xim = xr[NR - 1];
        xi = xnext;
        i += NR;
      }
      have_pf = false;
      if (i > i_hi) break;
      if (reg_top && i == i_lo) {
#pragma unroll
        for (int t = 0; t < W - 1; ++t) {
          const double xnext = x[imin2(i + 1, Tx - 1)];
          double left = bnd[sl * bs];
          double diag = Dg;
          Dg = left;
          const typename M::Row rw = m.row(i, xi, xim);
#pragma unroll
          for (int c = 0; c <= t; ++c) {
            const double up = prev[c];
            const double d = m.cell(up, left, diag, rw, cols[c], i, j0 + c);
            prev[c] = d;
            left = d;
            diag = up;
          }
          sl = (sl + 1 == NS) ? 0 : sl + 1;
          xim = xi;
          xi = xnext;
          ++i;
        }
      } else if (reg_bot && i == j0 + g.a + 1) {
#pragma unroll
        for (int t = 1; t < W; ++t) {
          // row i = j0 + a + t: cells t..W-1; the cell left of the band reads the sentinel
          const double xnext = x[imin2(i + 1, Tx - 1)];
          const typename M::Row rw = m.row(i, xi, xim);
          double left = m.lsent();
          double diag = prev[t - 1];
#pragma unroll
          for (int c = t; c < W; ++c) {
            const double up = prev[c];
            const double d = m.cell(up, left, diag, rw, cols[c], i, j0 + c);
            prev[c] = d;
            left = d;
            diag = up;
          }
          bnd[sl * bs] = left;
          if (EA && M::kColumnMinBound) colmin = dmin2(colmin, left);
          sl = (sl + 1 == NS) ? 0 : sl + 1;
          xim = xi;
          xi = xnext;
          ++i;
        }
      } else {
        generic_row();
      }
    }

    if (!last_strip) {
      if (jl + g.a + 1 <= Tx - 1) bnd[sl * bs] = M::kMsmBand ? stale : m.lsent();
      if (EA && M::kColumnMinBound && colmin > abandon) return WB_INF;
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c) result = sel_f64(c == wv - 1, prev[c], result);
    }
  }
  return m.finish(result, g);
}

}  // namespace wb
