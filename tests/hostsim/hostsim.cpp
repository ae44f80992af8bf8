// TEST INFRASTRUCTURE: compiles the device engines (wildboar_b200/csrc/engine_*.cuh) for the
// HOST with g++ and runs them one emulated thread at a time, so the band/strip logic can be
// checked bit-for-bit against the oracle without a GPU.  Not linked into the product library.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../wildboar_b200/csrc/dispatch.cuh"
#include "../../wildboar_b200/csrc/engine_rowscan.cuh"
#include "../../wildboar_b200/csrc/engine_strip.cuh"
#include "../../wildboar_b200/csrc/engine_band.cuh"
#include "../../wildboar_b200/csrc/prep.hpp"

using namespace wb;

template <int W>
struct StripRunner {
  template <class M>
  static double run(const Geom& g, const M& m, const double* x, const double* y, int NS, int bs, double abandon, int* ok) {
    if (!strip_supported<M>(g, W)) { *ok = 0; return 0; }
    *ok = 1;
    std::vector<double> bnd((size_t)NS * bs, -12345.0);
    if (abandon < INFINITY) return strip_pair<M, W, true>(g, m, x, y, bnd.data(), bs, abandon);
    // exercise several in-thread wavefront depths
    const double r2 = strip_pair<M, W, false, 2>(g, m, x, y, bnd.data(), bs, abandon);
    const double r1 = strip_pair<M, W, false, 1>(g, m, x, y, bnd.data(), bs, abandon);
    const double r4 = strip_pair<M, W, false, 4>(g, m, x, y, bnd.data(), bs, abandon);
    const double r3 = strip_pair<M, W, false, 3>(g, m, x, y, bnd.data(), bs, abandon);
    if (!(r1 == r2 && r2 == r4 && r3 == r2) && !(r1 != r1 && r2 != r2)) return -1e300;  // NR variants disagree
    return r2;
  }
};

// band-register engine, both instantiations (plain rows / blocked interior rows): value AND row-minimum maximum must agree
template <class M, int HB, int YS>
static double band_both(const Geom& g, const M& m, const double* x, const double* yp, double min_dist, double* mm) {
  double m0 = 0, m1 = 0;
  const double d0 = band_pair<M, HB, YS, false>(g, m, x, yp, min_dist, &m0);
  const double d1 = band_pair<M, HB, YS, true>(g, m, x, yp, min_dist, &m1);
  *mm = m0;
  const bool same = (d0 == d1 || (d0 != d0 && d1 != d1)) && (m0 == m1 || (m0 != m0 && m1 != m1));
  return same ? d0 : -1e300;
}

// engine: 1 row-scan, 2 strip.  ea: 0 distance() / 1 eadistance() semantics for R (ddtw, edr).
// Returns 0 ok, 1 unsupported by this engine, 2 bad args.
extern "C" int hostsim_pair(int engine, int W, int metric, const wb_params* p, const double* x, int64_t Tx,
                            const double* y, int64_t Ty, int ea, double min_dist_raw, int ns_extra, int bs,
                            double* out, double* out_rowminmax) {
  std::vector<double> dx, dy;
  int64_t tx = Tx, ty = Ty;
  int64_t R = compute_r(Tx < Ty ? Tx : Ty, p->r);
  if (is_derivative(metric)) {
    if ((Tx < Ty ? Tx : Ty) < 3) { *out = 0.0; return 0; }
    dx.resize(Tx - 2); dy.resize(Ty - 2);
    average_slope(x, Tx, dx.data()); average_slope(y, Ty, dy.data());
    x = dx.data(); y = dy.data(); tx = Tx - 2; ty = Ty - 2;
    if (ea) R = compute_r(tx < ty ? tx : ty, p->r);
  }
  if (metric == M_EDR && ea) R = compute_r(Tx, p->r);
  int64_t nmax = tx > ty ? tx : ty;
  std::vector<double> w, tw;
  if (metric == M_WDTW || metric == M_WLCSS) w = make_weights(p->g, nmax);
  if (metric == M_WDDTW) w = make_weights(p->g, nmax);
  if (metric == M_TWE) tw = make_tw(p->stiffness, nmax + 1);
  Tables t{w.empty() ? nullptr : w.data() + table_center(nmax), tw.empty() ? nullptr : tw.data() + table_center(nmax + 1)};
  PairCtx pc{0, 0};
  if (metric == M_ERP) { pc.sx = seq_gap_sum(x, tx, p->g); pc.sy = seq_gap_sum(y, ty, p->g); }
  if (metric == M_EDR) { pc.sx = seq_std(x, tx); pc.sy = seq_std(y, ty); }
  Geom g = make_geom((int)tx, (int)ty, (int)R);
  int rc = 0;
  bool known = with_policy(metric, *p, t, [&](auto m) {
    m.begin_pair(pc);
    if (engine == 4 || engine == 5) {
      // the YS = 32 instantiations: y sits in lane 5 of an interleaved group (metrics.cuh interleave32_*), every other
      // lane holds garbage; engine 4 = band (HB = W), 5 = row-scan
      using MM = decltype(m);
      std::vector<double> yi((size_t)interleave32_size(32, (int)ty), -4242.0);
      for (int64_t t = 0; t < ty; ++t) yi[(size_t)interleave32_index(5, (int)t, (int)ty)] = y[t];
      const double* yp = yi.data() + interleave32_base(5, (int)ty);
      double mm = 0;
      bool ok = true;
      if (engine == 5) {
        std::vector<double> b0((size_t)(nmax + 1) * bs, -777.0), b1((size_t)(nmax + 1) * bs, -888.0);
        *out = rowscan_pair<MM, 32>(g, m, x, yp, b0.data(), b1.data(), (long long)bs, min_dist_raw, &mm);
      } else {
        switch (W) {
          case 8: if ((ok = band_supported<MM>(g, 8))) *out = band_both<MM, 8, 32>(g, m, x, yp, min_dist_raw, &mm); break;
          case 16: if ((ok = band_supported<MM>(g, 16))) *out = band_both<MM, 16, 32>(g, m, x, yp, min_dist_raw, &mm); break;
          case 32: if ((ok = band_supported<MM>(g, 32))) *out = band_pair<MM, 32, 32>(g, m, x, yp, min_dist_raw, &mm); break;
          default: ok = false;
        }
      }
      if (!ok) rc = 1; else if (out_rowminmax) *out_rowminmax = mm;
    } else if (engine == 3) {
      // band-register engine, HB = W
      using MM = decltype(m);
      double mm = 0;
      bool ok = true;
      switch (W) {
        case 8: if ((ok = band_supported<MM>(g, 8))) *out = band_both<MM, 8, 1>(g, m, x, y, min_dist_raw, &mm); break;
        case 16: if ((ok = band_supported<MM>(g, 16))) *out = band_both<MM, 16, 1>(g, m, x, y, min_dist_raw, &mm); break;
        case 32: if ((ok = band_supported<MM>(g, 32))) *out = band_pair<MM, 32>(g, m, x, y, min_dist_raw, &mm); break;
        default: ok = false;
      }
      if (!ok) rc = 1; else if (out_rowminmax) *out_rowminmax = mm;
    } else if (engine == 1) {
      std::vector<double> b0((size_t)(nmax + 1) * bs, -777.0), b1((size_t)(nmax + 1) * bs, -888.0);
      double mm = 0;
      *out = rowscan_pair(g, m, x, y, b0.data(), b1.data(), (long long)bs, min_dist_raw, &mm);
      if (out_rowminmax) *out_rowminmax = mm;
    } else {
      int NS = strip_ring_slots(g, W) + ns_extra;
      int ok = 0;
      double r = 0;
      switch (W) {
        case 2: r = StripRunner<2>::run(g, m, x, y, NS, bs, min_dist_raw, &ok); break;
        case 3: r = StripRunner<3>::run(g, m, x, y, NS, bs, min_dist_raw, &ok); break;
        case 4: r = StripRunner<4>::run(g, m, x, y, NS, bs, min_dist_raw, &ok); break;
        case 8: r = StripRunner<8>::run(g, m, x, y, NS, bs, min_dist_raw, &ok); break;
        case 12: r = StripRunner<12>::run(g, m, x, y, NS, bs, min_dist_raw, &ok); break;
        case 16: r = StripRunner<16>::run(g, m, x, y, NS, bs, min_dist_raw, &ok); break;
        default: ok = 0;
      }
      if (!ok) rc = 1; else *out = r;
    }
  });
  if (!known) return 2;
  return rc;
}

// window statistics of the generic scaled subsequence metrics (metrics.cuh inc_window_stats_one == the body of k_inc_window_stats)
extern "C" void hostsim_inc_window_stats(const double* x, int64_t T, int64_t m, double* mean, double* stdv) {
  inc_window_stats_one(x, (int)T, (int)m, mean, stdv);
}

// interleaved series layout (metrics.cuh): emulate k_interleave32 on the host and read every series back the way the
// YS = 32 kernels do (base + 32 * t).  Returns the number of mismatches.
extern "C" long long hostsim_interleave_roundtrip(const double* src, long long n, int T) {
  std::vector<double> dst((size_t)interleave32_size(n, T), -1.0);
  for (long long o = 0; o < (long long)dst.size(); ++o) {
    long long e; int t;
    interleave32_source(o, T, &e, &t);
    if (e < n) dst[(size_t)o] = src[e * T + t];
  }
  long long bad = 0;
  for (long long e = 0; e < n; ++e) {
    const double* yp = dst.data() + interleave32_base(e, T);
    for (int t = 0; t < T; ++t) bad += (yp[32LL * t] != src[e * T + t]) + (dst[(size_t)interleave32_index(e, t, T)] != src[e * T + t]);
  }
  return bad;
}
