// Cooperative banded-DP engine: a GROUP OF LANES PER (x_i, y_j) PAIR, the band row distributed over the lanes' REGISTERS,
// neighbours exchanged with __shfl_sync -- no per-pair state in shared or global memory at all.
//
// Why (B200): the strip engine (engine_strip.cuh) keeps one boundary column of H + W values per PAIR; with a thread per
// pair that is 108 KB per warp for T = 4096, r = 0.05 (H = 407), i.e. 192-256 MB for the resident grid -- past the 126 MB
// L2, so ~1 TB/s of the traffic spills to HBM.  It also needs >= 2400 warps of 32 pairs to fill the chip, which a 200 x 200
// matrix (1250 warps) does not have.  Here the state of a pair is ONE band row of H values spread over G lanes (13 doubles
// per lane for H = 407, G = 32): it never leaves the register file, and a pair occupies G lanes instead of one, so small
// problems fill the machine (G = 4 for the 200 x 200, H = 29 case: 8 pairs per warp, 5000 warps).
//
// Geometry.  Band coordinate k = j - i + a, k in [0, H): the neighbours of cell (i, k) are
//     left = (i, k-1)      diag = (i-1, k)      up = (i-1, k+1).
// Lane l of a group owns the band coordinates [k0(l), k0(l) + wc(l)), wc = W ("wide") for the first n1 lanes and W - 1
// ("narrow") for the rest, n1 * W + n2 * (W - 1) = H exactly, so no lane carries padding cells (every lane executes the same
// unrolled W cells; the narrow lanes skip the last one -- one guarded cell per row instead of a predicate per cell).
// Systolic schedule with a skew of ONE row per lane: at step t lane l works on row i = t - l.
//     left of its first cell  = the last cell lane l-1 produced one step earlier        (__shfl_up, start of the step)
//     up   of its last cell   = the first cell lane l+1 produces IN THIS STEP (row i-1)  (__shfl_down, after cell 0)
// so a step is: shuffle, cell 0, shuffle, cells 1 .. W-1.  Two 64-bit shuffles per W cells.
//
// Column contexts (y values and what the metric derives from them) move one band coordinate to the left per row.  A lane
// keeps the contexts of the W columns of its current row plus the U - 1 columns that will enter during the next U - 1 rows
// (and the one column to the left of its first cell, whose y value some metrics need as "y[j-1]"): cols[q] = column
// j0 - 1 + q, j0 = column of cell 0.  In the FAST block (interior rows: the whole band inside the matrix, every lane
// active) U steps are unrolled and cell c of unrolled step r reads cols[1 + c + r] -- nothing moves inside the block; after
// it the array is shifted down by U registers and U new columns (consecutive addresses, loaded at the START of the block)
// enter at the top.  U * W cells per block keep the loop body inside the 32 KB L1.5 instruction cache (the first version
// unrolled W steps with the registers in rotation: 169 cells of msm = 85 KB of code, and ncu showed one "no instruction"
// stall per issued instruction).  Rows where the band sticks out of the matrix (i < a or i > Ty - max_len: 2 a of the Tx
// rows), the start-up / drain of the skew and every reference quirk (row 0, MSM's column 0 / stale left edge / extra cell)
// run in the MASKED step: one row per step, per-cell predicates, the array shifted by one per row.  The masked step is the
// band-register engine's row (engine_band.cuh) distributed over lanes; engine_rowscan.cuh remains the executable
// specification and tests/hostsim runs this file on the host (all lanes of a group in lockstep) against the oracle.
//
// Per-diagonal values (wdtw weights, twe's stiffness term) are constant per band coordinate (i - j = a - k): they are loop
// invariants held in registers.
//
// Scope: plain distances (pairwise / paired / pair lists) -- no row minima, no early abandoning (those callers keep the
// thread-per-pair engines, which visit whole rows).
#pragma once
#include "metrics.cuh"

namespace wb {

struct CoopLayout {
  int n_act;  // lanes of a group that hold band coordinates
  int n1;     // the first n1 of them hold W coordinates, the others W - 1
};

// n lanes with W or W-1 cells each must tile H exactly: n (W - 1) <= H <= n W, n <= G
WB_HD bool coop_layout(const Geom& g, int W, int G, CoopLayout* out) {
  if (W < 3) return false;
  const int n = (g.H + W - 1) / W;
  if (n < 1 || n > G || n * (W - 1) > g.H) return false;
  out->n_act = n;
  out->n1 = g.H - n * (W - 1);
  return true;
}

template <class M>
WB_HD bool coop_supported(const Geom& g, int W, int G) {
  CoopLayout l;
  if (g.Tx < 2 || g.Ty < 2) return false;
  if (M::kMsmBand && g.H < 3) return false;  // the stale-left rule needs band coordinate 1
  return coop_layout(g, W, G, &l);
}

template <class M, int W, int U>
struct CoopLane {
  using F = typename M::real;
  static constexpr int NC = W + U;  // contexts held: one left of cell 0, the W of the row, U - 1 that enter next
  F P[W];                     // previous row at this lane's band coordinates (updated in place, left to right)
  typename M::Col cols[NC];   // cols[q] = context of column j0 - 1 + q (j0 = column of cell 0 in the lane's current row)
  typename M::Dv dvs[W];      // per-diagonal values of the lane's band coordinates (rows >= 1)
  typename M::Col col0;       // MSM: context of column 0 (always evaluated, EL:1620-1628)
  F xi, xim;                  // sample of the lane's current row / of the row before
  F left_in, up_in, c0, last_out;  // exchanged values
  F cy, stale, stale_next, beyond;  // MSM: column-0 running value, never-reset left edge, row 0's extra cell
  F ynew[U];                  // fast block: samples of the U columns that enter after the block (loaded at its start)
  int gl, k0, wc;             // lane within the group, first band coordinate, cells held (W, W - 1, or 0: idle lane)
  bool top;                   // holds band coordinate H - 1

  WB_HD static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

  WB_HD typename M::Col load_col(const M& m, const F* __restrict__ y, int Ty, int j) const {
    const int jc = clampi(j, 0, Ty - 1);  // columns outside the matrix are never used (their cells are masked)
    return m.col(jc, y[jc], jc > 0 ? y[jc - 1] : F(0));
  }
  // context of cell c at unrolled step r (r = 0 in masked steps), linked to its left neighbour's
  // (column 0 keeps the "y[-1] = 0" its own context was built with: linked = false)
  template <int Q>
  WB_HD typename M::Col ctx(bool linked = true) const {
    typename M::Col c = cols[Q];
    ColLink<M>::apply(c, cols[Q > 0 ? Q - 1 : 0], linked);
    return c;
  }

  WB_HD void init(const Geom& g, const M& m, const CoopLayout& lay, int lane_in_group, const F* __restrict__ x,
                  const F* __restrict__ y) {
    gl = lane_in_group;
    if (gl < lay.n_act) {
      wc = gl < lay.n1 ? W : W - 1;
      k0 = gl < lay.n1 ? gl * W : lay.n1 * W + (gl - lay.n1) * (W - 1);
    } else { wc = 0; k0 = 0; }
    top = gl == lay.n_act - 1;
#pragma unroll
    for (int c = 0; c < W; ++c) {
      P[c] = m.prev_init();
      if (M::kHasDv) dvs[c] = m.dv_diag(g.a - (k0 + c));
    }
#pragma unroll
    for (int q = 0; q < NC; ++q) cols[q] = load_col(m, y, g.Ty, k0 - g.a - 1 + q);  // row 0: cell 0 is column k0 - a
    col0 = m.col(0, y[0], F(0));
    xi = x[0]; xim = F(0);
    left_in = up_in = c0 = last_out = F(0);
    cy = F(0); stale = m.lsent(); stale_next = stale; beyond = m.usent();
#pragma unroll
    for (int u = 0; u < U; ++u) ynew[u] = F(0);
  }

  // ---------------------------------------------------------------- masked step (one row, any row)
  // phase A: everything that does not need the right neighbour's value of THIS step: row set-up and cell 0
  struct RowInfo { int i, klo, kfirst, khi; bool act, msm_col0, col0_first; F left_first; typename M::Row rw; };
  RowInfo ri;
  F run_left;  // running `left` inside the row

  template <int C>
  WB_HD void masked_cell(const Geom& g, const M& m, F up) {
    const int k = k0 + C;
    if (ri.msm_col0 && k == ri.klo) {  // MSM: column 0 is parked here (next row's diagonal, this row's left), not a DP cell
      P[C] = cy;
      run_left = cy;
    } else if (C < wc && k >= ri.kfirst && k < ri.khi) {
      const F diag = (ri.col0_first && k == ri.klo) ? m.diag0(ri.i) : P[C];
      const F lf = (k == ri.kfirst) ? ri.left_first : run_left;
      const int j = ri.i - g.a + k;
      const F d = m.cell(up, lf, diag, ri.rw, ctx<1 + C>(j > 0), m.dv(ri.i, j));
      P[C] = d;
      run_left = d;
    }
  }
  template <int C>
  WB_HD void masked_cells_from(const Geom& g, const M& m, F upv) {
    if constexpr (C < W) {
      const F up = (C + 1 < W) ? ((C + 1 < wc) ? P[C + 1 < W ? C + 1 : W - 1] : upv) : upv;
      masked_cell<C>(g, m, up);
      masked_cells_from<C + 1>(g, m, upv);
    }
  }

  WB_HD void masked_a(const Geom& g, const M& m, int t) {
    ri.i = t - gl;
    ri.act = wc > 0 && ri.i >= 0 && ri.i < g.Tx;
    if (!ri.act) { c0 = P[0]; return; }
    const int i = ri.i;
    ri.klo = imax2(0, g.a - i);
    ri.khi = imin2(g.H, g.Ty - i + g.a);
    ri.msm_col0 = M::kMsmBand && i <= g.a;
    ri.kfirst = ri.klo + (ri.msm_col0 ? 1 : 0);
    ri.col0_first = !M::kMsmBand && i <= g.a;
    ri.rw = m.row(i, xi, xim);
    if (M::kMsmBand) {
      // column 0: cy[i] = cy[i-1] + cost(X[i], X[i-1], Y[0]) (up-branch only, EL:1620-1628); D[0][0] in row 0
      cy = (i == 0) ? m.cell(Num<F>::inf(), Num<F>::inf(), F(0), ri.rw, col0, m.dv(0, 0))
                    : m.cell(cy, Num<F>::inf(), Num<F>::inf(), ri.rw, col0, m.dv(i, 0));
      stale_next = P[1];  // band coordinate 1 of row i-1 (lane 0): the stale left edge of row i+1
      ri.left_first = (i <= g.a + 1) ? cy : stale;
    } else {
      ri.left_first = (i <= g.a) ? m.left0(i) : m.lsent();
    }
    run_left = left_in;
    // cell 0 reads its upper neighbour from the lane itself (wc >= 2)
    masked_cell<0>(g, m, P[1]);
    c0 = P[0];
  }

  // phase B: up_in = first cell of the right neighbour (row i-1) or what lies above the band
  WB_HD void masked_b(const Geom& g, const M& m, const F* __restrict__ x, const F* __restrict__ y) {
    if (!ri.act) return;
    const int i = ri.i;
    F upv = up_in;
    if (top) upv = (M::kMsmBand && i == 1) ? beyond : m.usent();
    if (i == 0) upv = m.prev_init();
    masked_cells_from<1>(g, m, upv);
    last_out = run_left;
    if (M::kMsmBand) {
      if (i == 0 && top && g.max_len < g.Ty) {
        // row 0 fills one cell beyond the band (EL:1615-1617): D[0][R], the upper neighbour of row 1's last cell
        const typename M::Col cj = m.col(g.max_len, y[g.max_len], y[g.max_len - 1]);
        beyond = m.cell(Num<F>::inf(), run_left, Num<F>::inf(), ri.rw, cj, m.dv(0, g.max_len));
      }
      stale = stale_next;
    }
    // next row: every context moves one position down, a new last column enters
#pragma unroll
    for (int q = 0; q + 1 < NC; ++q) cols[q] = cols[q + 1];
    cols[NC - 1] = load_col(m, y, g.Ty, (i + 1) - g.a + k0 - 1 + (NC - 1));
    xim = xi;
    xi = x[imin2(i + 1, g.Tx - 1)];
  }

  // ---------------------------------------------------------------- fast block: U unrolled steps
  // Preconditions (checked by the driver for the whole block): every holding lane is active (0 <= t - gl < Tx) and past
  // the rows with row-0 rules (i >= 2).  TOP = false: additionally every lane is past the rows where column 0 is inside the
  // band (i >= a + 1, MSM: i >= a + 2).  Cells to the RIGHT of the matrix (rows i > Ty - max_len) are simply computed on
  // clamped samples: nothing flows from a cell to its left, and the upper neighbour of a cell inside the matrix is inside
  // the matrix too (the valid range shrinks by one coordinate per row), so that garbage never reaches a valid cell.
  // TOP = true (rows i <= a + 1): column 0 sits at band coordinate klo = a - i; the cells left of it are garbage, and the
  // one cell AT klo takes the column-0 rules (left0 / diag0; MSM: the running value cy instead of a DP cell) -- one
  // compare + select per cell instead of the masked step's full predicate set.  Idle lanes (wc == 0) run along on garbage;
  // nothing they produce reaches a holding lane.
  WB_HD void fast_begin(const Geom& g, int t0, const F* __restrict__ y) {
    // samples of the U columns that enter after this block: cols[W .. W + U - 1] of row i + U
    const int jb = (t0 - gl) + U - g.a + k0 - 1 + W;
#pragma unroll
    for (int u = 0; u < U; ++u) ynew[u] = y[clampi(jb + u, 0, g.Ty - 1)];
  }
  // the cell of band coordinate k0 + C in a TOP row: column-0 rules at k == klo
  template <bool TOP, int Q, int C>
  WB_HD F fast_cell(const M& m, const typename M::Row& rw, F up, F left, int i, int klo) {
    F diag = P[C];
    bool at0 = false;
    if (TOP) {
      at0 = k0 + C == klo;  // klo = a - i: negative (never matched) once column 0 has left the band
      if (!M::kMsmBand) {
        left = at0 ? m.left0(i) : left;
        diag = at0 ? m.diag0(i) : diag;
      }
    }
    F d = m.cell(up, left, diag, rw, ctx<Q>(!TOP || !at0), dvs[C]);
    if (TOP && M::kMsmBand) d = at0 ? cy : d;  // column 0 is not a DP cell: parked running value (EL:1620-1628)
    P[C] = d;
    return d;
  }
  template <int R, bool TOP>
  WB_HD void fast_a(const Geom& g, const M& m, int t0, const F* __restrict__ x, F& xnext, typename M::Row& rw, int& klo) {
    const int i = t0 + R - gl;
    rw = m.row(i, xi, xim);
    klo = TOP ? g.a - i : -1;  // band coordinate of column 0 (negative: outside the band)
    F lf = left_in;
    if (M::kMsmBand) {
      if (TOP) cy = m.cell(cy, Num<F>::inf(), Num<F>::inf(), rw, col0, dvs[0]);  // msm has no per-diagonal values
      stale_next = P[1];
      if (gl == 0) lf = (TOP && i <= g.a + 1) ? cy : stale;
    } else if (gl == 0) lf = m.lsent();
    const F d = fast_cell<TOP, 1 + R, 0>(m, rw, P[1], lf, i, klo);
    c0 = d;
    run_left = d;
    xnext = x[clampi(i + 1, 0, g.Tx - 1)];  // next row's sample, in flight while cells 1 .. W-1 run
  }
  template <int R, bool TOP, int C>
  WB_HD void fast_cells_from(const M& m, const typename M::Row& rw, F upv, F& left, int i, int klo) {
    if constexpr (C + 1 < W) {
      const F up = (C == W - 2) ? (wc < W ? upv : P[W - 1]) : P[C + 1 < W ? C + 1 : W - 1];
      left = fast_cell<TOP, 1 + C + R, C>(m, rw, up, left, i, klo);
      fast_cells_from<R, TOP, C + 1>(m, rw, upv, left, i, klo);
    }
  }
  template <int R, bool TOP>
  WB_HD void fast_b(const Geom&, const M& m, int t0, const F& xnext, const typename M::Row& rw, int klo) {
    const int i = t0 + R - gl;
    const F upv = top ? m.usent() : up_in;
    F left = run_left;
    fast_cells_from<R, TOP, 1>(m, rw, upv, left, i, klo);
    if (wc == W) left = fast_cell<TOP, W + R, W - 1>(m, rw, upv, left, i, klo);
    last_out = left;
    xim = xi;
    xi = xnext;
    if (M::kMsmBand) stale = stale_next;
  }
  WB_HD void fast_end(const Geom& g, const M& m, int t0) {
    // the row advanced by U: contexts move down by U, the U prefetched columns enter at the top
#pragma unroll
    for (int q = 0; q < W; ++q) cols[q] = cols[q + U];
    const int jb = (t0 - gl) + U - g.a + k0 - 1 + W;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int jc = clampi(jb + u, 0, g.Ty - 1);
      cols[W + u] = m.col(jc, ynew[u], jc > 0 ? cols[W + u - 1].yj : F(0));
    }
  }

  // D[Tx-1][Ty-1] sits at band coordinate a + Ty - Tx of the last row
  WB_HD bool holds_result(const Geom& g) const { const int k = g.a + g.Ty - g.Tx; return wc > 0 && k >= k0 && k < k0 + wc; }
  WB_HD F result(const Geom& g) const {
    const int c = g.a + g.Ty - g.Tx - k0;
    F r = F(0);
#pragma unroll
    for (int q = 0; q < W; ++q) if (q == c) r = P[q];
    return r;
  }
};

// ---------------------------------------------------------------- driver, shared by the device kernel and tests/hostsim
// Grp provides: each(f) -- apply f to the lane(s) this executor owns (device: the thread's lane; host: all lanes of the
// group in turn); xchg_left() -- left_in <- last_out of the lane below; xchg_up() -- up_in <- c0 of the lane above.
template <class M, int W, int U>
struct CoopTmp { typename M::real xnext; typename M::Row rw; int klo; };

template <class M, int W, int U, int R, bool TOP, class Grp>
struct CoopFastSteps {
  WB_HD static void run(Grp& grp, const Geom& g, const M& m, int t0) {
    grp.xchg_left();
    grp.each([&](CoopLane<M, W, U>& L, const typename M::real* x, const typename M::real*, CoopTmp<M, W, U>& tmp) {
      L.template fast_a<R, TOP>(g, m, t0, x, tmp.xnext, tmp.rw, tmp.klo);
    });
    grp.xchg_up();
    grp.each([&](CoopLane<M, W, U>& L, const typename M::real*, const typename M::real*, CoopTmp<M, W, U>& tmp) {
      L.template fast_b<R, TOP>(g, m, t0, tmp.xnext, tmp.rw, tmp.klo);
    });
    CoopFastSteps<M, W, U, R + 1, TOP, Grp>::run(grp, g, m, t0);
  }
};
template <class M, int W, int U, bool TOP, class Grp>
struct CoopFastSteps<M, W, U, U, TOP, Grp> {
  WB_HD static void run(Grp&, const Geom&, const M&, int) {}
};

// first row from which no cell of the band is column 0 and no column-0 rule applies
template <class M>
WB_HD int coop_interior_row(const Geom& g) { return M::kMsmBand ? g.a + 2 : g.a + 1; }

template <class M, int W, int U, class Grp>
WB_HD void coop_run(Grp& grp, const Geom& g, const M& m, const CoopLayout& lay) {
  using L_t = CoopLane<M, W, U>;
  using T_t = CoopTmp<M, W, U>;
  using F = typename M::real;
  const int total = g.Tx + lay.n_act - 1;
  // steps at which EVERY holding lane is active and past row 1 (row 0 / row 1 have their own rules): [t_act, t_end]
  const int t_act = lay.n_act - 1 + 2;
  const int t_end = g.Tx - 1;                                       // lane 0 still has a row
  const int t_int = coop_interior_row<M>(g) + lay.n_act - 1;        // the highest lane has left the column-0 rows
  int t = 0;
  while (t < total) {
    if (t >= t_int && t + U - 1 <= t_end) {
      grp.each([&](L_t& L, const F*, const F* y, T_t&) { L.fast_begin(g, t, y); });
      CoopFastSteps<M, W, U, 0, false, Grp>::run(grp, g, m, t);
      grp.each([&](L_t& L, const F*, const F*, T_t&) { L.fast_end(g, m, t); });
      t += U;
    } else if (t >= t_act && t + U - 1 <= t_end && t < t_int) {
      grp.each([&](L_t& L, const F*, const F* y, T_t&) { L.fast_begin(g, t, y); });
      CoopFastSteps<M, W, U, 0, true, Grp>::run(grp, g, m, t);
      grp.each([&](L_t& L, const F*, const F*, T_t&) { L.fast_end(g, m, t); });
      t += U;
    } else {
      grp.xchg_left();
      grp.each([&](L_t& L, const F*, const F*, T_t&) { L.masked_a(g, m, t); });
      grp.xchg_up();
      grp.each([&](L_t& L, const F* x, const F* y, T_t&) { L.masked_b(g, m, x, y); });
      t += 1;
    }
  }
}

}  // namespace wb
