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

// Boundary-buffer slots needed per pair.  The buffer between strip s and strip s+1 holds the
// last column of strip s, D[i][jl], LINEARLY: slot(i) = i - (i_lo(s+1) - 1).  Strip s therefore
// reads slot i - B_in and writes slot i - B_out with B_out - B_in = i_lo(s+1) - i_lo(s) >= 0:
// writes trail reads inside the same buffer, no modulo arithmetic, and in the hot loop every
// access is pointer + immediate.  Reads reach at most min(Tx, H + W - 1) slots past slot 0; the
// fast path prefetches one NR-row group ahead (padding: NR <= 8 slots).
WB_HD int strip_ring_slots(const Geom& g, int W) { return imin2(g.Tx, g.H + W - 1) + 1 + 8; }

// first row of the strip that starts at column j0
template <class M>
WB_HD int strip_ilo(const Geom& g, int j0) {
  int v = imax2(0, j0 - g.max_len + 1);
  if (M::kMsmBand && j0 == g.max_len) v = 0;  // row 0 reaches one cell beyond the band (EL:1615-1617)
  return v;
}

// Opaque select: keeps the compiler from turning a chain of register selects into a
// dynamically indexed (local-memory) array access.
WB_HD double sel_real(bool p, double a, double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
  return r;
#else
  return p ? a : b;
#endif
}
WB_HD float sel_real(bool p, float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f32 %0, %1, %2, q; }" : "=f"(r) : "f"(a), "f"(b), "r"((int)p));
  return r;
#else
  return p ? a : b;
#endif
}

// bnd: boundary buffer of this pair; slot s lives at bnd[s * bs] (BS > 0: compile-time stride).
// abandon: early-abandon threshold on the RAW dp value (pre-finish); only honoured for
//          policies whose column minima lower-bound the result (DTW family).  Returns +INF
//          when abandoned.
//
// Rows of a strip fall in three phases: a top triangle (the band's upper edge crosses the
// strip), FULL rows (all W cells in band) and a bottom triangle.  Triangle rows go through
// `generic_row` (per-cell, warp-uniform predicates) or, for regular triangles, straight-line
// code.  Full rows -- (2R-1-W+1)/(2R-1+W-1) of all rows, 87 % for the headline shape -- take
// the predicate-free `fast path`, NR rows per iteration so that NR independent left->right
// dependency chains are in flight per thread.
template <class M, int W, bool EA, int NR = 2, int BS = 0>
WB_HD typename M::real strip_pair(const Geom& g, const M& m, const typename M::real* __restrict__ x,
                                  const typename M::real* __restrict__ y, typename M::real* bnd, int bs_rt,
                                  typename M::real abandon) {
  using F = typename M::real;  // F: bit-exact mode; float: fp32 mode
  const int Tx = g.Tx, Ty = g.Ty;
  const long long bs = BS > 0 ? BS : bs_rt;
  F result = F(0);
  const F left0c = m.left0(1);

  for (int j0 = 0; j0 < Ty; j0 += W) {
    const int wv = imin2(W, Ty - j0);
    const bool last_strip = (j0 + W >= Ty);
    const bool first_strip = (j0 == 0);
    const int jl = j0 + wv - 1;
    const int i_lo = strip_ilo<M>(g, j0);
    const int i_hi = imin2(Tx - 1, jl + g.a);
    // read slot of row i: i - B_in (slot 0 = row i_lo - 1, the first diagonal); write slot: i - B_out
    const int B_in = i_lo - 1;
    const int B_out = last_strip ? B_in : strip_ilo<M>(g, j0 + W) - 1;
    F* const rbase = bnd - (long long)B_in * bs;   // rbase[i * bs] = read slot of row i
    F* const wbase = bnd - (long long)B_out * bs;  // wbase[i * bs] = write slot of row i

    // y[j-1] of column c is the y[j] of column c-1 (one register, and the metrics' |x[i-1] - y[j-1]| terms become
    // common subexpressions of the neighbouring cell's |x[i] - y[j]|); columns clipped by the matrix border repeat
    // y[Ty-1] -- their cells are never used
    typename M::Col cols[W];
    {
      F yprev = (j0 > 0) ? y[j0 - 1] : F(0);
#pragma unroll
      for (int c = 0; c < W; ++c) {
        const int jj = imin2(j0 + c, Ty - 1);
        const F yj = y[jj];
        cols[c] = m.col(jj, yj, yprev);
        yprev = yj;
      }
    }
    F prev[W];
#pragma unroll
    for (int c = 0; c < W; ++c) prev[c] = m.usent();

    F Dg;
    if (first_strip) {
      // virtual column -1: left0 is the same constant for every row and diag0(i) == left0(i-1)
      // for i >= 1, so the first strip simply finds left0 in every slot it will read.
      for (int s = 0; s <= i_hi + 1; ++s) bnd[s * bs] = left0c;
      Dg = m.diag0(0);
    } else if (i_lo == 0) Dg = m.prev_init();
    else Dg = bnd[0];
    F xi = x[i_lo];
    F xim = (M::kNeedPrevX && i_lo > 0) ? x[i_lo - 1] : F(0);
    F stale = m.lsent();
    F colmin = Num<F>::inf();
    int i = i_lo;

    auto generic_row = [&]() {
      const F xnext = x[imin2(i + 1, Tx - 1)];  // prefetch next row's sample
      const int js = row_js<M>(g, i), je = row_je<M>(g, i);
      const int clo = imax2(js - j0, 0), chi = imin2(je - j0, wv);
      F left = rbase[i * bs];
      F diag = Dg;
      Dg = left;
      if (i == 0) {
#pragma unroll
        for (int c = 0; c < W; ++c) if (c < chi) prev[c] = m.prev_init();
      }
      const typename M::Row rw = m.row(i, xi, xim);
      const int cst = M::kMsmBand ? (row_js<M>(g, i + 1) - 1 - j0) : -1;
      F stale_next = stale;
#pragma unroll
      for (int c = 0; c < W; ++c) {
        const F up = prev[c];
        if (M::kMsmBand && c == cst) stale_next = up;
        if (c >= clo && c < chi) {
          if (c > 0 && c == clo) left = M::kMsmBand ? stale : m.lsent();
          const F d = m.cell(up, left, diag, rw, cols[c], m.dv(i, j0 + c));
          prev[c] = d;
          left = d;
        }
        diag = up;
      }
      stale = stale_next;
      // (MSM's row 0 reaches one cell beyond the band; no later strip reads that value)
      if (chi == W && (!M::kMsmBand || i >= B_out)) {
        const F b = prev[W - 1];
        wbase[i * bs] = b;
        if (EA && M::kColumnMinBound) colmin = dmin2(colmin, b);
      }
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

    // One copy of every row routine; the hot NR-row loop stays a tight inner loop.
    for (;;) {
      if (i >= fa && i + NR - 1 <= fb) {
        // ---- fast path: rows i .. i+NR-1, all W cells in band, no predicates; NR independent
        // left->right dependency chains (an in-thread anti-diagonal wavefront).  All boundary
        // and x accesses are pointer + immediate; the boundary values and x samples of the NEXT
        // group are loaded one group ahead (hides the L1/L2 latency of the global rings).
        F* rp = rbase + (long long)i * bs;
        F* wp = wbase + (long long)i * bs;
        const F* xp = x + i;
        F pf[NR], xn[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) pf[r] = rp[r * bs];
        xn[0] = xi;
#pragma unroll
        for (int r = 1; r < NR; ++r) xn[r] = xp[r];
        do {
          F xr[NR], lft[NR], dg[NR];
          typename M::Row rws[NR];
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            xr[r] = xn[r];
            lft[r] = pf[r];
            rws[r] = m.row(i + r, xr[r], r == 0 ? xim : xr[r - 1]);
          }
          // next group's operands (the rows exist whenever the loop continues; otherwise the
          // clamped x pointer / the padded buffer keep the loads in bounds and the values unused)
          const F* xq = (i + 2 * NR <= Tx) ? xp + NR : xp;
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            pf[r] = rp[(NR + r) * bs];
            xn[r] = xq[r];
          }
          // per-diagonal values (weights / stiffness terms) of the group's W + NR - 1 diagonals:
          // pointer + immediate loads from the signed tables, nothing for the other metrics
          typename M::Dv dvs[W + NR - 1];
          if (M::kHasDv) {
#pragma unroll
            for (int d = 0; d < W + NR - 1; ++d) dvs[d] = m.dv_diag(i - j0 + d - (W - 1));
          }
          dg[0] = Dg;
#pragma unroll
          for (int r = 1; r < NR; ++r) dg[r] = lft[r - 1];
          Dg = lft[NR - 1];
#pragma unroll
          for (int c = 0; c < W; ++c) {
            F v = prev[c];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              const F d = m.cell(v, lft[r], dg[r], rws[r], cols[c], dvs[r - c + W - 1]);
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
            wp[r * bs] = lft[r];
            if (EA && M::kColumnMinBound) colmin = dmin2(colmin, lft[r]);
          }
          xim = xr[NR - 1];
          rp += NR * bs;
          wp += NR * bs;
          xp += NR;
          i += NR;
        } while (i + NR - 1 <= fb);
        xi = x[imin2(i, Tx - 1)];
      }
      if (i > i_hi) break;
      if (reg_top && i == i_lo) {
#pragma unroll
        for (int t = 0; t < W - 1; ++t) {
          const F xnext = x[imin2(i + 1, Tx - 1)];
          F left = rbase[i * bs];
          F diag = Dg;
          Dg = left;
          const typename M::Row rw = m.row(i, xi, xim);
#pragma unroll
          for (int c = 0; c <= t; ++c) {
            const F up = prev[c];
            const F d = m.cell(up, left, diag, rw, cols[c], m.dv(i, j0 + c));
            prev[c] = d;
            left = d;
            diag = up;
          }
          xim = xi;
          xi = xnext;
          ++i;
        }
      } else if (reg_bot && i == j0 + g.a + 1) {
#pragma unroll
        for (int t = 1; t < W; ++t) {
          // row i = j0 + a + t: cells t..W-1; the cell left of the band reads the sentinel
          const F xnext = x[imin2(i + 1, Tx - 1)];
          const typename M::Row rw = m.row(i, xi, xim);
          F left = m.lsent();
          F diag = prev[t - 1];
#pragma unroll
          for (int c = t; c < W; ++c) {
            const F up = prev[c];
            const F d = m.cell(up, left, diag, rw, cols[c], m.dv(i, j0 + c));
            prev[c] = d;
            left = d;
            diag = up;
          }
          wbase[i * bs] = left;
          if (EA && M::kColumnMinBound) colmin = dmin2(colmin, left);
          xim = xi;
          xi = xnext;
          ++i;
        }
      } else {
        generic_row();
      }
    }

    if (!last_strip) {
      if (jl + g.a + 1 <= Tx - 1) wbase[(jl + g.a + 1) * bs] = M::kMsmBand ? stale : m.lsent();
      if (EA && M::kColumnMinBound && colmin > abandon) return Num<F>::inf();
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c) result = sel_real(c == wv - 1, prev[c], result);
    }
  }
  return m.finish(result, g);
}

}  // namespace wb
