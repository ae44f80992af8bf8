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
// Column contexts (y values and what the metric derives from them) move one band coordinate to the left per row.  In the
// FAST block (interior rows: the whole band inside the matrix, every lane active) W steps are unrolled and the W context
// registers are used in rotation -- cell c of unrolled step r reads cols[(c + r) % W] -- so nothing is moved: the register
// of the column that leaves the lane is refilled with the column that enters it, loaded one step ahead.  Rows where the band
// sticks out of the matrix (i < a or i > Ty - max_len: 2 a of the Tx rows), the start-up / drain of the skew and every
// reference quirk (row 0, MSM's column 0 / stale left edge / extra cell) run in the MASKED step: one row per step, per-cell
// predicates, contexts shifted by register moves.  The masked step is the band-register engine's row (engine_band.cuh)
// distributed over lanes; engine_rowscan.cuh remains the executable specification and tests/hostsim runs this file on the
// host (all lanes of a group in lockstep) against the oracle.
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

template <class M, int W>
struct CoopLane {
  using F = typename M::real;
  F P[W];                     // previous row at this lane's band coordinates (updated in place, left to right)
  typename M::Col cols[W];    // column contexts; masked steps: cell c <-> cols[c]; fast step r: cell c <-> cols[(c + r) % W]
  typename M::Dv dvs[W];      // per-diagonal values of the lane's band coordinates (rows >= 1)
  typename M::Col col0;       // MSM: context of column 0 (always evaluated, EL:1620-1628)
  F xi, xim;                  // sample of the lane's current row / of the row before
  F left_in, up_in, c0, last_out;  // exchanged values
  F cy, stale, stale_next, beyond;  // MSM: column-0 running value, never-reset left edge, row 0's extra cell
  int gl, k0, wc;             // lane within the group, first band coordinate, cells held (W, W - 1, or 0: idle lane)
  bool top;                   // holds band coordinate H - 1

  WB_HD static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

  WB_HD typename M::Col load_col(const M& m, const F* __restrict__ y, int Ty, int j) const {
    const int jc = clampi(j, 0, Ty - 1);  // columns outside the matrix are never used (their cells are masked)
    return m.col(jc, y[jc], jc > 0 ? y[jc - 1] : F(0));
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
      cols[c] = load_col(m, y, g.Ty, k0 + c - g.a);  // row 0: j = k - a
      if (M::kHasDv) dvs[c] = m.dv_diag(g.a - (k0 + c));
    }
    col0 = m.col(0, y[0], F(0));
    xi = x[0]; xim = F(0);
    left_in = up_in = c0 = last_out = F(0);
    cy = F(0); stale = m.lsent(); stale_next = stale; beyond = m.usent();
  }

  // ---------------------------------------------------------------- masked step (one row, any row)
  // phase A: everything that does not need the right neighbour's value of THIS step: row set-up and cell 0
  struct RowInfo { int i, klo, kfirst, khi; bool act, msm_col0, col0_first; F left_first; typename M::Row rw; };
  RowInfo ri;
  F run_left;  // running `left` inside the row

  WB_HD void masked_cell(const Geom& g, const M& m, int c, F up) {
    const int k = k0 + c;
    if (ri.msm_col0 && k == ri.klo) {  // MSM: column 0 is parked here (next row's diagonal, this row's left), not a DP cell
      P[c] = cy;
      run_left = cy;
    } else if (c < wc && k >= ri.kfirst && k < ri.khi) {
      const F diag = (ri.col0_first && k == ri.klo) ? m.diag0(ri.i) : P[c];
      const F lf = (k == ri.kfirst) ? ri.left_first : run_left;
      const F d = m.cell(up, lf, diag, ri.rw, cols[c], m.dv(ri.i, ri.i - g.a + k));
      P[c] = d;
      run_left = d;
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
    masked_cell(g, m, 0, P[1]);
    c0 = P[0];
  }

  // phase B: up_in = first cell of the right neighbour (row i-1) or what lies above the band
  WB_HD void masked_b(const Geom& g, const M& m, const F* __restrict__ x, const F* __restrict__ y) {
    if (!ri.act) return;
    const int i = ri.i;
    F upv = up_in;
    if (top) upv = (M::kMsmBand && i == 1) ? beyond : m.usent();
    if (i == 0) upv = m.prev_init();
#pragma unroll
    for (int c = 1; c < W; ++c) {
      const F up = (c + 1 < W) ? ((c + 1 < wc) ? P[c + 1 < W ? c + 1 : W - 1] : upv) : upv;
      masked_cell(g, m, c, up);
    }
    last_out = run_left;
    if (M::kMsmBand) {
      if (i == 0 && top && g.max_len < g.Ty) {
        // row 0 fills one cell beyond the band (EL:1615-1617): D[0][R], the upper neighbour of row 1's last cell
        const typename M::Col cj = m.col(g.max_len, y[g.max_len], y[g.max_len - 1]);
        beyond = m.cell(Num<F>::inf(), run_left, Num<F>::inf(), ri.rw, cj, m.dv(0, g.max_len));
      }
      stale = stale_next;
    }
    // next row: every context moves one band coordinate to the left, the new last column enters
#pragma unroll
    for (int c = 0; c + 1 < W; ++c) cols[c] = cols[c + 1];
    cols[W - 1] = load_col(m, y, g.Ty, (i + 1) - g.a + k0 + W - 1);
    xim = xi;
    xi = x[imin2(i + 1, g.Tx - 1)];
  }

  // ---------------------------------------------------------------- fast step R of an unrolled block of W steps
  // Preconditions (checked by the driver for the whole block): every holding lane's row i = t - gl is an interior row
  // (a <= i <= Ty - max_len, i >= 1, MSM: i >= a + 2), so all wc cells are inside the matrix and no row-0 / column-0 rule
  // applies.  Idle lanes (wc == 0) run along on garbage; nothing they produce reaches a holding lane.
  template <int R>
  WB_HD void fast_a(const Geom& g, const M& m, int t, const F* __restrict__ x, const F* __restrict__ y, F& xnext, F& ynew,
                    int& jn, typename M::Row& rw) {
    const int i = t - gl;
    rw = m.row(i, xi, xim);
    if (M::kMsmBand) stale_next = P[1];
    const F lf = (gl == 0) ? (M::kMsmBand ? stale : m.lsent()) : left_in;
    const F d = m.cell(P[1], lf, P[0], rw, cols[R % W], dvs[0]);
    P[0] = d;
    c0 = d;
    run_left = d;
    // operands of the NEXT step, in flight while cells 1 .. W-1 run
    jn = imin2(imax2(i + 1 - g.a + k0 + W - 1, 0), g.Ty - 1);
    ynew = y[jn];
    xnext = x[imin2(imax2(i + 1, 0), g.Tx - 1)];
  }
  template <int R>
  WB_HD void fast_b(const Geom& g, const M& m, const F& xnext, const F& ynew, int jn, const typename M::Row& rw) {
    const F upv = top ? m.usent() : up_in;
    F left = run_left;
#pragma unroll
    for (int c = 1; c + 1 < W; ++c) {
      const F up = (c == W - 2) ? (wc < W ? upv : P[W - 1]) : P[c + 1 < W ? c + 1 : W - 1];
      const F d = m.cell(up, left, P[c], rw, cols[(c + R) % W], dvs[c]);
      P[c] = d;
      left = d;
    }
    if (wc == W) {
      const F d = m.cell(upv, left, P[W - 1], rw, cols[(W - 1 + R) % W], dvs[W - 1]);
      P[W - 1] = d;
      left = d;
    }
    last_out = left;
    // the register of the column that left the lane (cell 0's) takes the column that enters it (next step's cell W-1)
    cols[R % W] = m.col(jn, ynew, cols[(W - 1 + R) % W].yj);
    xim = xi;
    xi = xnext;
    if (M::kMsmBand) stale = stale_next;
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
template <class M, int W, int R, class Grp>
struct CoopFastSteps {
  WB_HD static void run(Grp& grp, const Geom& g, const M& m, int t0) {
    grp.xchg_left();
    grp.each([&](CoopLane<M, W>& L, const typename M::real* x, const typename M::real* y, auto& tmp) {
      L.template fast_a<R>(g, m, t0 + R, x, y, tmp.xnext, tmp.ynew, tmp.jn, tmp.rw);
    });
    grp.xchg_up();
    grp.each([&](CoopLane<M, W>& L, const typename M::real*, const typename M::real*, auto& tmp) {
      L.template fast_b<R>(g, m, tmp.xnext, tmp.ynew, tmp.jn, tmp.rw);
    });
    CoopFastSteps<M, W, R + 1, Grp>::run(grp, g, m, t0);
  }
};
template <class M, int W, class Grp>
struct CoopFastSteps<M, W, W, Grp> {
  WB_HD static void run(Grp&, const Geom&, const M&, int) {}
};

template <class M, int W>
struct CoopTmp { typename M::real xnext, ynew; int jn; typename M::Row rw; };

template <class M, int W, class Grp>
WB_HD void coop_run(Grp& grp, const Geom& g, const M& m, const CoopLayout& lay) {
  const int total = g.Tx + lay.n_act - 1;
  // rows every cell of which is inside the matrix and free of the row-0 / column-0 rules
  const int row_lo = M::kMsmBand ? g.a + 2 : imax2(g.a, 1);
  const int row_hi = imin2(g.Tx - 1, g.Ty - g.max_len);
  const int t_lo = row_lo + lay.n_act - 1;  // the highest lane has reached row_lo
  const int t_hi = row_hi;                  // lane 0 has not passed row_hi
  int t = 0;
  while (t < total) {
    if (t >= t_lo && t + W - 1 <= t_hi) {
      CoopFastSteps<M, W, 0, Grp>::run(grp, g, m, t);
      t += W;
    } else {
      grp.xchg_left();
      grp.each([&](CoopLane<M, W>& L, const typename M::real*, const typename M::real*, auto&) { L.masked_a(g, m, t); });
      grp.xchg_up();
      grp.each([&](CoopLane<M, W>& L, const typename M::real* x, const typename M::real* y, auto&) { L.masked_b(g, m, x, y); });
      t += 1;
    }
  }
}

}  // namespace wb
