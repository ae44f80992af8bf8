// Host-side preparation shared by the library and the host simulator: everything that needs
// libm or a fixed sequential summation order is computed here exactly as the reference does.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace wb {

// EL:1924-1925
inline int64_t compute_r(int64_t length, double r) {
  return (int64_t)std::fmax(std::floor((double)length * r), 1.0);
}

// Lookup tables indexed by the SIGNED diagonal offset d = i - j (so that the DP cell needs no
// abs() and the strip engine's hot loop can address them as pointer + immediate):
// t[center + d] = f(|d|) for |d| < n, with kTablePad zero entries on both sides -- the fast
// path also evaluates the clipped columns of a partial last strip (results unused) and indexes up
// to W-1 entries past the last meaningful one.  Kernels receive the pointer to the CENTER.
constexpr int64_t kTablePad = 16;
inline int64_t table_center(int64_t n) { return (n > 0 ? n : 0) + kTablePad; }

// EL:3339-3341 (wdtw/wlcss: n = max(Tx,Ty)), EL:3418-3428 (wddtw: n = max(Tx,Ty) - 2)
inline std::vector<double> make_weights(double g, int64_t n) {
  const int64_t c = table_center(n);
  std::vector<double> w((size_t)(2 * c + 1), 0.0);
  for (int64_t i = 0; i < n; i++) {
    const double v = 1.0 / (1.0 + std::exp(-g * ((double)i - (double)n / 2.0)));
    w[(size_t)(c + i)] = v; w[(size_t)(c - i)] = v;
  }
  return w;
}

// EL:1813: stiffness * 2 * labs(i - j)
inline std::vector<double> make_tw(double stiffness, int64_t n) {
  const int64_t c = table_center(n);
  std::vector<double> t((size_t)(2 * c + 1), 0.0);
  for (int64_t k = 0; k < n; k++) {
    const double v = stiffness * 2 * (double)k;
    t[(size_t)(c + k)] = v; t[(size_t)(c - k)] = v;
  }
  return t;
}

// utils/_stats.pyx:22-42
inline double seq_std(const double* d, int64_t n) {
  double ex = 0, ex2 = 0;
  for (int64_t i = 0; i < n; i++) { double v = d[i]; ex += v; ex2 += std::pow(v, 2.0); }
  double mean = ex / (double)n;
  ex2 = ex2 / (double)n - mean * mean;
  return ex2 > 1e-13 ? std::sqrt(ex2) : 0.0;
}

// EL:1295-1303
inline double seq_gap_sum(const double* d, int64_t n, double g) {
  double s = 0;
  for (int64_t i = 0; i < n; i++) s += std::fabs(d[i] - g);
  return s;
}

// EL:3220-3225
inline void average_slope(const double* q, int64_t len, double* d) {
  int64_t j = 0;
  for (int64_t i = 1; i < len - 1; i++) { d[j] = ((q[i] - q[i - 1]) + ((q[i + 1] - q[i - 1]) / 2)) / 2; j++; }
}

}  // namespace wb
