// Runtime metric id -> compile-time policy.
#pragma once
#include "metrics.cuh"
#include "../../include/wb_cuda.h"

namespace wb {

struct Tables {
  const double* weights;  // wdtw / wddtw / wlcss
  const double* tw;       // twe
};

// Calls f(policy) with the policy object for `metric`; returns false for an unknown id.
template <class F>
inline bool with_policy(int metric, const wb_params& p, const Tables& t, F&& f) {
  switch (metric) {
    case M_DTW: case M_DDTW: { DtwPolicy<false, false> m; m.w = nullptr; m.p = 0; f(m); return true; }
    case M_WDTW: case M_WDDTW: { DtwPolicy<true, false> m; m.w = t.weights; m.p = 0; f(m); return true; }
    case M_ADTW: { DtwPolicy<false, true> m; m.w = nullptr; m.p = p.p; f(m); return true; }
    case M_LCSS: { LcssPolicy<false> m; m.w = nullptr; m.eps = p.epsilon; f(m); return true; }
    case M_WLCSS: { LcssPolicy<true> m; m.w = t.weights; m.eps = p.epsilon; f(m); return true; }
    case M_ERP: { ErpPolicy m; m.g = p.g; m.gx_sum = 0; m.gy_sum = 0; f(m); return true; }
    case M_EDR: { EdrPolicy m; m.eps_param = p.epsilon; m.eps = p.epsilon; f(m); return true; }
    case M_MSM: { MsmPolicy m; m.cf = (float)p.c; m.c = (double)m.cf; f(m); return true; }
    case M_TWE: { TwePolicy m; m.pen = p.penalty + p.stiffness; m.tw = t.tw; f(m); return true; }
  }
  return false;
}

inline bool is_derivative(int metric) { return metric == M_DDTW || metric == M_WDDTW; }
inline bool is_dtw_family(int metric) {
  return metric == M_DTW || metric == M_WDTW || metric == M_DDTW || metric == M_ADTW || metric == M_WDDTW;
}

}  // namespace wb
