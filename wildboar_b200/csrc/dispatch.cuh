// Runtime metric id -> compile-time policy.
#pragma once
#include "metrics.cuh"
#include "../../include/wb_cuda.h"

namespace wb {

template <class F>
struct TablesT {
  const F* weights;  // wdtw / wddtw / wlcss (centre of the signed table)
  const F* tw;       // twe (centre of the signed table)
};
using Tables = TablesT<double>;

// metrics with an fp32 variant (lcss / wlcss / edr are step functions of a threshold test and
// always run in double)
inline bool has_fp32_variant(int metric) { return !(metric == M_LCSS || metric == M_WLCSS || metric == M_EDR); }

// Calls f(policy) with the policy object for `metric`; returns false for an unknown id.
template <class Fn>
inline bool with_policy(int metric, const wb_params& p, const Tables& t, Fn&& f) {
  if (p.precision == 2) {
    // optional fused-multiply-add mode (DTW family only: the other recurrences contain no multiplication)
    switch (metric) {
      case M_DTW: case M_DDTW: { DtwPolicy<false, false, double, true> m; m.w = nullptr; m.p = 0; f(m); return true; }
      case M_WDTW: case M_WDDTW: { DtwPolicy<true, false, double, true> m; m.w = t.weights; m.p = 0; f(m); return true; }
      case M_ADTW: { DtwPolicy<false, true, double, true> m; m.w = nullptr; m.p = p.p; f(m); return true; }
      default: break;
    }
  }
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
    case M_SCALED_DTW: { ScaledDtwPolicy m; m.w = nullptr; m.p = 0; m.mean = 0; m.stdv = 1; f(m); return true; }
  }
  return false;
}

// fp32 mode: same policies instantiated for float
template <class Fn>
inline bool with_policy_f32(int metric, const wb_params& p, const TablesT<float>& t, Fn&& f) {
  switch (metric) {
    case M_DTW: case M_DDTW: { DtwPolicy<false, false, float> m; m.w = nullptr; m.p = 0; f(m); return true; }
    case M_WDTW: case M_WDDTW: { DtwPolicy<true, false, float> m; m.w = t.weights; m.p = 0; f(m); return true; }
    case M_ADTW: { DtwPolicy<false, true, float> m; m.w = nullptr; m.p = (float)p.p; f(m); return true; }
    case M_ERP: { ErpPolicyT<float> m; m.g = (float)p.g; m.gx_sum = 0; m.gy_sum = 0; f(m); return true; }
    case M_MSM: { MsmPolicyT<float> m; m.cf = (float)p.c; m.c = m.cf; f(m); return true; }
    case M_TWE: { TwePolicyT<float> m; m.pen = (float)(p.penalty + p.stiffness); m.tw = t.tw; f(m); return true; }
  }
  return false;
}

inline bool is_derivative(int metric) { return metric == M_DDTW || metric == M_WDDTW; }
inline bool is_dtw_family(int metric) {
  return metric == M_DTW || metric == M_WDTW || metric == M_DDTW || metric == M_ADTW || metric == M_WDDTW;
}

}  // namespace wb
