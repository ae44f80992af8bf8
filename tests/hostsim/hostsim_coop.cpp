// TEST INFRASTRUCTURE: the cooperative engine (wildboar_b200/csrc/engine_coop.cuh) compiled for the HOST with g++; all G
// lanes of one group are emulated in lockstep (the shuffles become array shifts), so the systolic schedule, the masked /
// top / fast steps and every reference quirk can be checked bit for bit against the oracle without a GPU.  A separate
// translation unit (and shared object) from hostsim.cpp so that the two compile in parallel.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../wildboar_b200/csrc/dispatch.cuh"
#include "../../wildboar_b200/csrc/engine_coop.cuh"
#include "../../wildboar_b200/csrc/prep.hpp"

using namespace wb;

// ---- cooperative engine (engine_coop.cuh): all G lanes of one group emulated in lockstep ----
template <class M, int W, int U>
struct CoopHostGroup {
  std::vector<CoopLane<M, W, U>> L;
  std::vector<CoopTmp<M, W, U>> tmp;
  const double* x; const double* y;
  template <class Fn> void each(Fn f) { for (size_t l = 0; l < L.size(); ++l) f(L[l], x, y, tmp[l]); }
  void xchg_left() {  // __shfl_up_sync(.., 1, G): lane 0 keeps its own value
    for (size_t l = L.size(); l-- > 1;) L[l].left_in = L[l - 1].last_out;
    L[0].left_in = L[0].last_out;
  }
  void xchg_up() {  // __shfl_down_sync(.., 1, G): the last lane keeps its own value
    for (size_t l = 0; l + 1 < L.size(); ++l) L[l].up_in = L[l + 1].c0;
    L.back().up_in = L.back().c0;
  }
};

template <class M, int W, int U>
static int coop_host_pair(const Geom& g, const M& m, int G, const double* x, const double* y, double* out) {
  CoopLayout lay;
  if (!coop_supported<M>(g, W, G) || !coop_layout(g, W, G, &lay)) return 1;
  CoopHostGroup<M, W, U> grp;
  grp.L.resize((size_t)G); grp.tmp.resize((size_t)G); grp.x = x; grp.y = y;
  for (int l = 0; l < G; ++l) grp.L[(size_t)l].init(g, m, lay, l, x, y);
  coop_run<M, W, U>(grp, g, m, lay);
  int holders = 0;
  for (int l = 0; l < G; ++l) if (grp.L[(size_t)l].holds_result(g)) { *out = m.finish(grp.L[(size_t)l].result(g), g); ++holders; }
  return holders == 1 ? 0 : 3;
}

// Returns 0 ok, 1 geometry not supported by this (W, G), 2 bad args, 3 internal error.
extern "C" int hostsim_coop_pair(int W, int G, int metric, const wb_params* p, const double* x, int64_t Tx, const double* y,
                                 int64_t Ty, double* out) {
  std::vector<double> dx, dy;
  int64_t tx = Tx, ty = Ty;
  int64_t R = compute_r(Tx < Ty ? Tx : Ty, p->r);
  if (is_derivative(metric)) {
    if ((Tx < Ty ? Tx : Ty) < 3) { *out = 0.0; return 0; }
    dx.resize(Tx - 2); dy.resize(Ty - 2);
    average_slope(x, Tx, dx.data()); average_slope(y, Ty, dy.data());
    x = dx.data(); y = dy.data(); tx = Tx - 2; ty = Ty - 2;
  }
  int64_t nmax = tx > ty ? tx : ty;
  std::vector<double> w, tw;
  if (metric == M_WDTW || metric == M_WLCSS || metric == M_WDDTW) w = make_weights(p->g, nmax);
  if (metric == M_TWE) tw = make_tw(p->stiffness, nmax + 1);
  Tables t{w.empty() ? nullptr : w.data() + table_center(nmax), tw.empty() ? nullptr : tw.data() + table_center(nmax + 1)};
  PairCtx pc{0, 0};
  if (metric == M_ERP) { pc.sx = seq_gap_sum(x, tx, p->g); pc.sy = seq_gap_sum(y, ty, p->g); }
  if (metric == M_EDR) { pc.sx = seq_std(x, tx); pc.sy = seq_std(y, ty); }
  Geom g = make_geom((int)tx, (int)ty, (int)R);
  int rc = 2;
  bool known = with_policy(metric, *p, t, [&](auto m) {
    using MM = decltype(m);
    m.begin_pair(pc);
    switch (W) {
      // (W, U): the shipped layouts (8, 8), (13, 4) and (4, 8) plus odd ones that stress the shifting
      case 3: rc = coop_host_pair<MM, 3, 1>(g, m, G, x, y, out); break;
      case 4: rc = coop_host_pair<MM, 4, 8>(g, m, G, x, y, out); break;
      case 104: rc = coop_host_pair<MM, 4, 5>(g, m, G, x, y, out); break;
      case 8: rc = coop_host_pair<MM, 8, 8>(g, m, G, x, y, out); break;
      case 13: rc = coop_host_pair<MM, 13, 4>(g, m, G, x, y, out); break;
      case 108: rc = coop_host_pair<MM, 8, 3>(g, m, G, x, y, out); break;
      default: rc = 2;
    }
  });
  return known ? rc : 2;
}

