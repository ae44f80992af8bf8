// Metric policies for the banded-DP engines.
//
// Every elastic metric of the reference (wildboar/distance/_elastic.pyx, "EL") is expressed
// as ONE uniform recurrence over DP cells (i, j):
//
//     D[i][j] = cell(up = D[i-1][j], left = D[i][j-1], diag = D[i-1][j-1], row ctx, col ctx)
//
// plus a small set of boundary constants that reproduce what the reference's two scratch
// rows contain just outside the Sakoe-Chiba band:
//
//     prev_init : value of the virtual row -1 inside the band of row 0
//     usent     : what a cell reads as `up` when (i-1, j) is above the band
//     lsent     : what a cell reads as `left` when (i, j-1) is left of the band
//                 (MSM: never reset by the reference => stale value, handled by the engine)
//     left0(i), diag0(i) : the j == 0 special cases
//
// The reference special-cases row 0 for dtw/adtw/msm (prefix sums); with the constants
// below the general recurrence produces the identical operations, e.g. for DTW row 0
// min(min(INF, left), INF) + c == left + c and 0 + c == c are exact.
//
// All arithmetic is IEEE double with ONE rounding per operation: the library is compiled with
// -fmad=false, so no FMA contraction happens and results are bit-identical to the
// reference's -O2 x86-64 build (which contains no FMA instructions).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define WB_HD __host__ __device__ __forceinline__
#else
#define WB_HD inline
#endif

namespace wb {

enum MetricId : int {
  M_DTW = 0, M_WDTW = 1, M_DDTW = 2, M_ADTW = 3, M_LCSS = 4, M_ERP = 5,
  M_EDR = 6, M_MSM = 7, M_TWE = 8, M_WDDTW = 9, M_WLCSS = 10, M_COUNT = 11,
  M_SCALED_DTW = 100  // internal: scaled_dtw subsequence search (not a pairwise metric of the public enum)
};

#if defined(__CUDA_ARCH__)
#define WB_INF __longlong_as_double(0x7ff0000000000000LL)
#else
#define WB_INF ((double)INFINITY)
#endif

WB_HD double dmin2(double a, double b) { return a < b ? a : b; }
WB_HD double dmax2(double a, double b) { return a > b ? a : b; }
// fp32 mode (optional, <= 1e-4 relative): native FMNMX / FMNMX3
WB_HD float dmin2(float a, float b) { return fminf(a, b); }
WB_HD float dmax2(float a, float b) { return fmaxf(a, b); }
WB_HD int imin2(int a, int b) { return a < b ? a : b; }
WB_HD int imax2(int a, int b) { return a > b ? a : b; }
WB_HD int iabs1(int a) { return a < 0 ? -a : a; }

// Read-only global load (table lookups that are uniform across a warp).
template <class F>
WB_HD F ldg(const F* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// +infinity of the arithmetic type (F = double: the bit-exact mode; F = float: the fp32 mode)
template <class F> struct Num;
template <> struct Num<double> { WB_HD static double inf() { return WB_INF; } };
template <> struct Num<float> {
  WB_HD static float inf() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x7f800000);
#else
    return (float)INFINITY;
#endif
  }
};

// Band geometry shared by every metric (EL:890-891, 909-910 and the same expressions inlined
// in the other kernels).  R = max(floor(min(Tx,Ty) * r), 1) is computed by the caller.
struct Geom {
  int Tx, Ty;   // rows (first operand), columns (second operand)
  int max_len;  // max(0, Ty-Tx) + R
  int a;        // min_len + R - 1, min_len = max(0, Tx-Ty): row i starts at max(0, i - a)
  int H;        // band height of a column: a + max_len
  int raw;      // 1: DTW-family policies return the squared-cost DP value without the final sqrt (subsequence search
                //    compares in that domain, EL:622-660)
};

WB_HD Geom make_geom(int Tx, int Ty, int R) {
  Geom g;
  g.Tx = Tx; g.Ty = Ty;
  g.max_len = imax2(0, Ty - Tx) + R;
  g.a = imax2(0, Tx - Ty) + R - 1;
  g.H = g.a + g.max_len;
  g.raw = 0;
  return g;
}

// Per-pair scalars a metric may need (computed by small prologue kernels / the host).
struct PairCtx {
  double sx;  // erp: sum |x - g|   edr: std(x)
  double sy;  // erp: sum |y - g|   edr: std(y)      scaled dtw: mean of the window
  double sy2; //                                      scaled dtw: std of the window
};

// ------------------------------------------------------------------------------------------
// DTW family.  EL:869-940 (dtw / wdtw / ddtw / wddtw), EL:943-1008 (adtw).
// WEIGHTED: cost is (v*v)*w[widx]; row 0 uses w[max(j-1,0)] (EL:894-903 quirk).
// AMERCING: min(min(up+p, left+p), diag) + v*v, no penalty on row 0 (EL:966-971).
// ------------------------------------------------------------------------------------------
// FMA (optional "fp64_fma" mode, wb_params.precision == 2): the cost is folded into the minimum with ONE fused
// multiply-add, fma(v, v, m) resp. fma(v*v, w, m) -- one rounding instead of two, 4 instead of 5 FP64-pipe instructions
// per cell.  Not bit-equal to the reference (whose build has no FMA) but within the north star's 1e-12 relative.
template <bool WEIGHTED, bool AMERCING, class F = double, bool FMA = false>
struct DtwPolicy {
  using real = F;
  static constexpr bool kMsmBand = false;
  static constexpr bool kNeedPrevX = false;
  static constexpr bool kColumnMinBound = true;  // column minima lower-bound the result
  const F* w;  // WEIGHTED: CENTER of the signed weights table, w[d] = weight(|d|), d = i - j
  F p;         // penalty (AMERCING)

  WB_HD F prev_init() const { return Num<F>::inf(); }
  WB_HD F usent() const { return Num<F>::inf(); }
  WB_HD F lsent() const { return Num<F>::inf(); }
  WB_HD F left0(int) const { return Num<F>::inf(); }
  WB_HD F diag0(int i) const { return i == 0 ? F(0) : Num<F>::inf(); }
  WB_HD void begin_pair(const PairCtx&) {}

  struct Row { F xi; F p; };
  struct Col { F yj; };
  WB_HD Row row(int i, F xi, F) const {
    Row r; r.xi = xi; r.p = (AMERCING && i > 0) ? p : F(0); return r;
  }
  WB_HD Col col(int, F yj, F) const { Col c; c.yj = yj; return c; }
  // per-diagonal value: the weight; row 0 uses w[max(j-1,0)] (EL:894-903 quirk)
  struct Dv { F w; };
  static constexpr bool kHasDv = WEIGHTED;
  WB_HD Dv dv(int i, int j) const { Dv d; d.w = WEIGHTED ? ldg(w + ((i == 0) ? imax2(j - 1, 0) : (i - j))) : F(1); return d; }
  WB_HD Dv dv_diag(int d) const { Dv v; v.w = WEIGHTED ? ldg(w + d) : F(1); return v; }  // rows >= 1

  WB_HD F cell(F up, F left, F diag, const Row& r, const Col& c, const Dv& d) const {
    F v = r.xi - c.yj;
    if (AMERCING) { up = up + r.p; left = left + r.p; }
    if (sizeof(F) == 4) {
      // fp32 mode: FADD, FMNMX3, FFMA (contraction allowed: the mode's contract is 1e-4 relative)
      const F m = dmin2(dmin2(up, left), diag);
      return WEIGHTED ? (F)fmaf((float)(v * v), (float)d.w, (float)m) : (F)fmaf((float)v, (float)v, (float)m);
    }
    if (FMA) {
      const F m = dmin2(dmin2(up, left), diag);
      return WEIGHTED ? (F)fma((double)(v * v), (double)d.w, (double)m) : (F)fma((double)v, (double)v, (double)m);
    }
    F cost = v * v;
    if (WEIGHTED) cost = cost * d.w;
    return dmin2(dmin2(up, left), diag) + cost;
  }
  WB_HD F finish(F d, const Geom& g) const { return g.raw ? d : sqrt(d); }
};

// ------------------------------------------------------------------------------------------
// Scaled (z-normalised) DTW of the UCR-suite subsequence search, `inner_scaled_dtw_subsequence_distance` EL:263-345:
// v = (S[i] - s_mean) / s_std - (X[j] - mean) / std.  The first operand arrives already normalised; the second
// operand's samples are normalised when a strip loads its columns, with the window's running mean / std (EL:390-401).
// ------------------------------------------------------------------------------------------
struct ScaledDtwPolicy : DtwPolicy<false, false, double> {
  double mean, stdv;
  WB_HD void begin_pair(const PairCtx& pc) { mean = pc.sy; stdv = pc.sy2; }
  WB_HD Col col(int, double yj, double) const { Col c; c.yj = (yj - mean) / stdv; return c; }
};

// Window statistics of the generic scaled subsequence metrics (ScaledSubsequenceMetricWrap._distance, CD:505-531) with the
// reference's IncStats (utils/_stats.pyx:45-93): Welford add of the newest sample, mean / variance of the window
// (variance below 1e-13 -> 0 -> std 1; fewer than two samples -> std 1), Welford removal of the oldest sample -- in scan
// order, so every rounding matches.  mean[w], stdv[w] for the windows w = 0 .. T - m of one series p[0 .. T).
WB_HD void inc_window_stats_one(const double* p, int T, int m, double* mean, double* stdv) {
  double mu = 0.0, ns = 0.0, ss = 0.0;
  for (int t = 0; t < T; ++t) {
    const double v = p[t];
    ns += 1.0;
    const double next_m = mu + (v - mu) / ns;
    ss += (v - mu) * (v - next_m);
    mu = next_m;
    if (t >= m - 1) {
      const int w = t - (m - 1);
      double var = 0.0;
      if (!(ns <= 1)) { var = ss / ns; if (var < 1e-13) var = 0.0; }
      mean[w] = mu;
      stdv[w] = var == 0.0 ? 1.0 : sqrt(var);
      const double old = p[w];
      if (ns == 1.0) { ns = 0.0; mu = 0.0; ss = 0.0; }
      else {
        const double old_m = (ns * mu - old) / (ns - 1.0);
        ss -= (old - mu) * (old - old_m);
        mu = old_m;
        ns -= 1.0;
      }
    }
  }
}

// Interleaved series layout (KArgs::yil, k_interleave32): n series of length T regrouped 32 at a time so that element t of
// the 32 series of a group is contiguous.  Element t of series e sits at interleave32_index(e, t, T); a thread that owns
// series e walks it from interleave32_base(e, T) with an element stride of 32; the buffer holds interleave32_size(n, T)
// elements (the last group is padded).
WB_HD long long interleave32_base(long long e, int T) { return (e >> 5) * (32LL * T) + (e & 31); }
WB_HD long long interleave32_index(long long e, int t, int T) { return interleave32_base(e, T) + 32LL * t; }
WB_HD long long interleave32_size(long long n, int T) { return ((n + 31) / 32) * 32LL * T; }
// inverse, used by the copy kernel (which enumerates destination offsets so that its stores coalesce)
WB_HD void interleave32_source(long long o, int T, long long* e, int* t) {
  const long long g = o / (32LL * T);
  const long long r = o - g * 32LL * T;
  *t = (int)(r >> 5);
  *e = g * 32 + (r & 31);
}

// ------------------------------------------------------------------------------------------
// LCSS / WLCSS.  EL:1118-1183; result 1 - s / min(Tx,Ty).
// ------------------------------------------------------------------------------------------
template <bool WEIGHTED>
struct LcssPolicy {
  using real = double;  // no fp32 variant: the result is a step function of |x-y| <= eps
  static constexpr bool kMsmBand = false;
  static constexpr bool kNeedPrevX = false;
  static constexpr bool kColumnMinBound = false;
  const double* w;  // WEIGHTED: center of the signed weights table
  double eps;

  WB_HD double prev_init() const { return 0.0; }
  WB_HD double usent() const { return 0.0; }
  WB_HD double lsent() const { return 0.0; }
  WB_HD double left0(int) const { return 0.0; }
  WB_HD double diag0(int) const { return 0.0; }
  WB_HD void begin_pair(const PairCtx&) {}

  struct Row { double xi; };
  struct Col { double yj; };
  WB_HD Row row(int, double xi, double) const { Row r; r.xi = xi; return r; }
  WB_HD Col col(int, double yj, double) const { Col c; c.yj = yj; return c; }

  struct Dv { double w; };
  static constexpr bool kHasDv = WEIGHTED;
  WB_HD Dv dv(int i, int j) const { return dv_diag(i - j); }
  WB_HD Dv dv_diag(int d) const { Dv v; v.w = WEIGHTED ? ldg(w + d) : 1.0; return v; }

  WB_HD double cell(double up, double left, double diag, const Row& r, const Col& c, const Dv& d) const {
    double v = fabs(r.xi - c.yj);
    double wv = 1.0;
    if (WEIGHTED) wv = d.w;
    double hit = wv + diag;
    double miss = dmax2(left, up);
    return (v <= eps) ? hit : miss;
  }
  WB_HD double finish(double d, const Geom& g) const { return 1 - (d / (double)imin2(g.Tx, g.Ty)); }
};

// ------------------------------------------------------------------------------------------
// ERP.  EL:1273-1347.  sx/sy are the whole-series gap sums (sequential order).
// ------------------------------------------------------------------------------------------
template <class F = double>
struct ErpPolicyT {
  using real = F;
  static constexpr bool kMsmBand = false;
  static constexpr bool kNeedPrevX = false;
  static constexpr bool kColumnMinBound = false;
  F g;
  F gx_sum, gy_sum;

  WB_HD F prev_init() const { return gy_sum; }
  WB_HD F usent() const { return F(0); }
  WB_HD F lsent() const { return F(0); }
  WB_HD F left0(int) const { return gx_sum; }
  WB_HD F diag0(int i) const { return i == 0 ? F(0) : gx_sum; }
  WB_HD void begin_pair(const PairCtx& pc) { gx_sum = (F)pc.sx; gy_sum = (F)pc.sy; }

  struct Row { F xi, gx; };
  struct Col { F yj, gy; };
  WB_HD Row row(int, F xi, F) const { Row r; r.xi = xi; r.gx = fabs(xi - g); return r; }
  WB_HD Col col(int, F yj, F) const { Col c; c.yj = yj; c.gy = fabs(yj - g); return c; }

  struct Dv {};
  static constexpr bool kHasDv = false;
  WB_HD Dv dv(int, int) const { return Dv(); }
  WB_HD Dv dv_diag(int) const { return Dv(); }

  WB_HD F cell(F up, F left, F diag, const Row& r, const Col& c, const Dv&) const {
    F v = fabs(r.xi - c.yj);
    return dmin2(diag + v, dmin2(up + r.gx, left + c.gy));
  }
  WB_HD F finish(F d, const Geom&) const { return d; }
};

// ------------------------------------------------------------------------------------------
// EDR.  EL:1437-1497; eps < 0 on entry means "default": max(std_x, std_y) / 4 (EL:3762-3766).
// ------------------------------------------------------------------------------------------
struct EdrPolicy {
  using real = double;  // no fp32 variant: the result is a step function of |x-y| < eps
  static constexpr bool kMsmBand = false;
  static constexpr bool kNeedPrevX = false;
  static constexpr bool kColumnMinBound = false;
  double eps_param;  // NaN => per-pair default
  double eps;

  WB_HD double prev_init() const { return 0.0; }
  WB_HD double usent() const { return 0.0; }
  WB_HD double lsent() const { return 0.0; }
  WB_HD double left0(int) const { return 0.0; }
  WB_HD double diag0(int) const { return 0.0; }
  WB_HD void begin_pair(const PairCtx& pc) {
    eps = (eps_param != eps_param) ? dmax2(pc.sx, pc.sy) / 4.0 : eps_param;
  }

  struct Row { double xi; };
  struct Col { double yj; };
  WB_HD Row row(int, double xi, double) const { Row r; r.xi = xi; return r; }
  WB_HD Col col(int, double yj, double) const { Col c; c.yj = yj; return c; }

  struct Dv {};
  static constexpr bool kHasDv = false;
  WB_HD Dv dv(int, int) const { return Dv(); }
  WB_HD Dv dv_diag(int) const { return Dv(); }

  WB_HD double cell(double up, double left, double diag, const Row& r, const Col& c, const Dv&) const {
    double v = fabs(r.xi - c.yj);
    double sub = diag + ((v < eps) ? 0.0 : 1.0);
    return dmin2(dmin2(sub, up + 1.0), left + 1.0);
  }
  WB_HD double finish(double d, const Geom& g) const { return d / (double)imax2(g.Tx, g.Ty); }
};

// ------------------------------------------------------------------------------------------
// MSM.  EL:1583-1647.  `_msm_cost` takes C floats: operands are rounded to fp32, the two
// subtractions happen in fp32, fabs/min/+c in double.  min(|(double)a|, |(double)b|) ==
// (double)min(|a|, |b|) exactly, so one widening conversion per call suffices.
// Band quirks (column 0 always evaluated, row 0 one cell wider, stale left) live in the engine
// behind kMsmBand.
// ------------------------------------------------------------------------------------------
template <class F = double>
struct MsmPolicyT {
  using real = F;
  static constexpr bool kMsmBand = true;
  static constexpr bool kNeedPrevX = true;
  static constexpr bool kColumnMinBound = false;
  F c;  // (F)(float)c
  float cf;

  WB_HD F prev_init() const { return Num<F>::inf(); }
  WB_HD F usent() const { return F(0); }
  WB_HD F lsent() const { return F(0); }  // only for bands the stale rule cannot reach
  WB_HD F left0(int) const { return Num<F>::inf(); }
  WB_HD F diag0(int i) const { return i == 0 ? F(0) : Num<F>::inf(); }
  WB_HD void begin_pair(const PairCtx&) {}

  // _msm_cost(x, y, z) = c + (x between y and z ? 0 : min(|x-y|, |x-z|)), differences in fp32.
  // With a = x-y and b = x-z (fp32; the rounded differences keep the exact sign and are zero
  // iff the operands are equal), "between" <=> a and b have opposite signs or one is zero, and
  // in the zero case min(|a|,|b|) is 0 anyway.  So the extra cost is
  //     signbit(a) != signbit(b) ? 0 : min(|a|, |b|)
  // -- one xor + one select instead of four compares.  Per cell only ONE fp32 subtraction is
  // new: for the "up" move a = X[i]-X[i-1] is a row constant and b = X[i]-Y[j]; for the
  // "left" move a = Y[j]-X[i] = -b (negation is exact) and b = Y[j]-Y[j-1] is a column constant.
  struct Row { F xi; float xf, dx; };   // dx = (float)X[i] - (float)X[i-1]
  struct Col { F yj; float yf, dy; };   // dy = (float)Y[j] - (float)Y[j-1]
  WB_HD Row row(int, F xi, F xim) const { Row r; r.xi = xi; r.xf = (float)xi; r.dx = r.xf - (float)xim; return r; }
  WB_HD Col col(int, F yj, F yjm) const { Col c2; c2.yj = yj; c2.yf = (float)yj; c2.dy = c2.yf - (float)yjm; return c2; }

  WB_HD static unsigned fbits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u; memcpy(&u, &f, 4); return u;
#endif
  }
  // extra cost = c + (sign(a) != sign(b) ? 0 : min(|a|, |b|)).  NEG: the first operand is -a (the
  // sign test is inverted instead of negating; a zero operand makes the minimum zero either way).
  // Device: pinned with inline PTX to LOP3 (predicate out) + FMNMX + FSEL + one F2F -- left to
  // the compiler the select migrates behind the conversion (two more 32-bit selects per call), and
  // MSM is bound by exactly that pipe (profiles/r01c_ncu_msm_twe.md).
  template <bool NEG>
  WB_HD F extra(float a, float b) const {
#if defined(__CUDA_ARCH__)
    float e;
    if (NEG)
      asm("{ .reg .pred p; .reg .b32 t; lop3.b32 t, %1, %2, 0x80000000, 0x28; setp.eq.u32 p, t, 0;"
          " min.f32 %0, %3, %4; selp.f32 %0, 0f00000000, %0, p; }"
          : "=f"(e) : "r"(fbits(a)), "r"(fbits(b)), "f"(fabsf(a)), "f"(fabsf(b)));
    else
      asm("{ .reg .pred p; .reg .b32 t; lop3.b32 t, %1, %2, 0x80000000, 0x28; setp.ne.u32 p, t, 0;"
          " min.f32 %0, %3, %4; selp.f32 %0, 0f00000000, %0, p; }"
          : "=f"(e) : "r"(fbits(a)), "r"(fbits(b)), "f"(fabsf(a)), "f"(fabsf(b)));
    return c + (F)e;
#else
    const float m = fminf(fabsf(a), fabsf(b));
    const bool opposite = (((fbits(a) ^ fbits(b)) >> 31) != 0u) != NEG;
    return c + (F)(opposite ? 0.0f : m);
#endif
  }
  struct Dv {};
  static constexpr bool kHasDv = false;
  WB_HD Dv dv(int, int) const { return Dv(); }
  WB_HD Dv dv_diag(int) const { return Dv(); }

  WB_HD F cell(F up, F left, F diag, const Row& r, const Col& cl, const Dv&) const {
    const float xy = r.xf - cl.yf;                  // (float)X[i] - (float)Y[j]
    const F a = diag + fabs(r.xi - cl.yj);
    const F b = up + extra<false>(r.dx, xy);   // _msm_cost(X[i], X[i-1], Y[j])
    const F d = left + extra<true>(xy, cl.dy); // _msm_cost(Y[j], X[i], Y[j-1]): a = Y[j]-X[i] = -xy
    return dmin2(dmin2(a, b), d);
  }
  WB_HD F finish(F d, const Geom&) const { return d; }
};

// ------------------------------------------------------------------------------------------
// TWE.  EL:1733-1829.  Conventions X[-1] = Y[-1] = 0; pen = penalty + stiffness;
// tw[k] = (stiffness*2)*k is a table so that no int->double conversion sits in the cell.
// ------------------------------------------------------------------------------------------
template <class F = double>
struct TwePolicyT {
  using real = F;
  static constexpr bool kMsmBand = false;
  static constexpr bool kNeedPrevX = true;
  static constexpr bool kColumnMinBound = false;
  F pen;        // penalty + stiffness
  const F* tw;  // CENTER of the signed table tw[d] = (stiffness * 2) * |d|, d = i - j

  WB_HD F prev_init() const { return Num<F>::inf(); }
  WB_HD F usent() const { return F(0); }
  WB_HD F lsent() const { return F(0); }
  WB_HD F left0(int) const { return Num<F>::inf(); }
  WB_HD F diag0(int i) const { return i == 0 ? F(0) : Num<F>::inf(); }
  WB_HD void begin_pair(const PairCtx&) {}

  struct Row { F xi, xim, dx; };
  struct Col { F yj, yjm, dy; };
  WB_HD Row row(int, F xi, F xim) const { Row r; r.xi = xi; r.xim = xim; r.dx = fabs(xim - xi); return r; }
  WB_HD Col col(int, F yj, F yjm) const { Col c; c.yj = yj; c.yjm = yjm; c.dy = fabs(yjm - yj); return c; }
  // cooperative engine: y[j-1] is the left neighbour context's sample -- no need to hold it per column
  WB_HD static void link(Col& c, const Col& prev, bool has_prev) { c.yjm = has_prev ? prev.yj : F(0); }

  struct Dv { F t; };
  static constexpr bool kHasDv = true;
  WB_HD Dv dv(int i, int j) const { return dv_diag(i - j); }
  WB_HD Dv dv_diag(int d) const { Dv v; v.t = ldg(tw + d); return v; }

  WB_HD F cell(F up, F left, F diag, const Row& r, const Col& c, const Dv& d) const {
    F del_x = (up + r.dx) + pen;
    F del_y = (left + c.dy) + pen;
    F match = ((diag + fabs(r.xi - c.yj)) + fabs(r.xim - c.yjm)) + d.t;
    return dmin2(dmin2(del_x, del_y), match);
  }
  WB_HD F finish(F d, const Geom&) const { return d; }
};

// ColLink / link(cur, prev, has_prev) -- engines that keep the column contexts of neighbouring columns in registers
// (engine_coop.cuh, the blocked rows of engine_band.cuh): metrics whose column context carries the previous column's sample (twe: y[j-1], 0 for
// column 0) take it from the neighbouring context instead of holding it per column, so that field never occupies a
// register -- policies may define `static void link(Col&, const Col&, bool)`.
template <class M, class = void> struct ColLink {
  WB_HD static void apply(typename M::Col&, const typename M::Col&, bool) {}
};
template <class M> struct ColLink<M, decltype(M::link(*(typename M::Col*)nullptr, *(const typename M::Col*)nullptr, true))> {
  WB_HD static void apply(typename M::Col& c, const typename M::Col& prev, bool has_prev) { M::link(c, prev, has_prev); }
};

using ErpPolicy = ErpPolicyT<double>;
using MsmPolicy = MsmPolicyT<double>;
using TwePolicy = TwePolicyT<double>;

}  // namespace wb
