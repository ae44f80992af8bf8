/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement (plain C, scalar, IEEE double, no FMA contraction) of wildboar's
 * elastic-distance hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library, and only as the checker
 * or as the timed CPU baseline -- never as a fallback for the CUDA path.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here (a) bit-for-bit
 * against the reference's own Cython build (oracle/_ref, built by oracle/build_ref.sh)
 * when that build is present and (b) against the committed golden vectors in
 * tests/golden/ that were generated from that same reference build
 * (tests/golden/make_golden.py).  The SURVEY 8f restatements further down (alignments, DBA helpers,
 * subsequence distances / scans / matches / profiles / k closest windows, IncStats) are pinned the same
 * way by tests/test_next_rows.py and tests/test_subsequence_scan.py (golden vectors from
 * tests/golden/make_golden_next.py and make_golden_scan.py).
 *
 * Citations are file:line in the reference tree (/root/reference/src/wildboar):
 *   EL = distance/_elastic.pyx   CD = distance/_cdistance.pyx
 *   MI = utils/_misc.pyx         ST = utils/_stats.pyx      LB = distance/lb.py
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -pthread -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

enum {
  ORC_DTW = 0, ORC_WDTW = 1, ORC_DDTW = 2, ORC_ADTW = 3, ORC_LCSS = 4, ORC_ERP = 5,
  ORC_EDR = 6, ORC_MSM = 7, ORC_TWE = 8, ORC_WDDTW = 9, ORC_WLCSS = 10, ORC_N_METRICS = 11
};

/* metric_params surface of the reference; epsilon = NaN means
 * "EDR default": max(std_x, std_y) / 4 per pair (EL:3762-3768). */
typedef struct {
  double r;         /* Sakoe-Chiba window fraction, all metrics          */
  double g;         /* wdtw / wddtw / wlcss weight steepness, erp gap g  */
  double p;         /* adtw penalty                                      */
  double c;         /* msm cost                                          */
  double epsilon;   /* lcss / wlcss / edr threshold (NaN = edr default)  */
  double penalty;   /* twe                                               */
  double stiffness; /* twe                                               */
} orc_params;

static inline int64_t i64max(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* EL:1924-1925 */
int64_t orc_compute_r(int64_t length, double r) {
  return (int64_t)fmax(floor((double)length * r), 1.0);
}

/* ST:22-42 fast_mean_std: sequential sums, pow(x, 2.0), variance threshold 1e-13 */
double orc_std(const double *data, int64_t n) {
  double ex = 0, ex2 = 0;
  for (int64_t i = 0; i < n; i++) {
    double v = data[i];
    ex += v;
    ex2 += pow(v, 2.0);
  }
  double mean = ex / (double)n;
  ex2 = ex2 / (double)n - mean * mean;
  return ex2 > 1e-13 ? sqrt(ex2) : 0.0;
}

/* EL:3339-3341 (N = max(Tx,Ty)); EL:3418-3428 for wddtw (N = max(Tx,Ty)-2) */
void orc_weights(double g, int64_t n, double *w) {
  for (int64_t i = 0; i < n; i++) w[i] = 1.0 / (1.0 + exp(-g * ((double)i - (double)n / 2.0)));
}

/* EL:3220-3225 */
void orc_average_slope(const double *q, int64_t len, double *d) {
  int64_t j = 0;
  for (int64_t i = 1; i < len - 1; i++) {
    d[j] = ((q[i] - q[i - 1]) + ((q[i + 1] - q[i - 1]) / 2)) / 2;
    j++;
  }
}

/* EL:869-940 (weights may be NULL) */
static double dtw_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                           double *cost, double *cost_prev, const double *wv, double min_dist) {
  double w = 1.0, v;
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  v = X[0] - Y[0];
  if (wv) w = wv[0];
  cost_prev[0] = v * v * w;
  for (int64_t i = 1; i < i64min(yl, max_len); i++) {
    v = X[0] - Y[i];
    if (wv) w = wv[i - 1];
    cost_prev[i] = cost_prev[i - 1] + v * v * w;
  }
  if (max_len < yl) cost_prev[max_len] = INFINITY;
  for (int64_t i = 1; i < xl; i++) {
    int64_t j_start = i64max(0, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    if (j_start > 0) cost[j_start - 1] = INFINITY;
    double min_cost = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      double x = cost_prev[j], y, z;
      if (j > 0) { y = cost[j - 1]; z = cost_prev[j - 1]; }
      else { y = INFINITY; z = INFINITY; }
      v = X[i] - Y[j];
      if (wv) w = wv[llabs(i - j)];
      cost[j] = dmin(dmin(x, y), z) + v * v * w;
      if (cost[j] < min_cost) min_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = INFINITY;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return cost_prev[yl - 1];
}

/* EL:943-1008 */
static double adtw_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                            double *cost, double *cost_prev, double penalty, double min_dist) {
  double v;
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  v = X[0] - Y[0];
  cost_prev[0] = v * v;
  for (int64_t i = 1; i < i64min(yl, max_len); i++) {
    v = X[0] - Y[i];
    cost_prev[i] = cost_prev[i - 1] + v * v;
  }
  if (max_len < yl) cost_prev[max_len] = INFINITY;
  for (int64_t i = 1; i < xl; i++) {
    int64_t j_start = i64max(0, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    if (j_start > 0) cost[j_start - 1] = INFINITY;
    double prev_cost = INFINITY, min_cost = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      double x = cost_prev[j] + penalty, y, z;
      if (j > 0) { y = prev_cost + penalty; z = cost_prev[j - 1]; }
      else { y = INFINITY; z = INFINITY; }
      v = X[i] - Y[j];
      cost[j] = dmin(dmin(x, y), z) + v * v;
      if (cost[j] < min_cost) min_cost = cost[j];
      prev_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = INFINITY;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return cost_prev[yl - 1];
}

/* EL:1118-1183 */
static double lcss_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                            double epsilon, double *cost, double *cost_prev, const double *wv,
                            double min_dist) {
  double w = 1.0;
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  for (int64_t i = 0; i < i64min(yl, max_len); i++) cost_prev[i] = 0;
  if (max_len < yl) cost_prev[max_len] = 0;
  for (int64_t i = 0; i < xl; i++) {
    int64_t j_start = i64max(0, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    if (j_start > 0) cost[j_start - 1] = 0;
    double min_cost = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      double x = cost_prev[j], y, z;
      if (j > 0) { y = cost_prev[j - 1]; z = cost[j - 1]; }
      else { y = 0; z = 0; }
      double v = fabs(X[i] - Y[j]);
      if (wv) w = wv[llabs(i - j)];
      if (v <= epsilon) cost[j] = w + y;
      else cost[j] = dmax(z, x);
      if (cost[j] < min_cost) min_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = 0;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return 1 - (cost_prev[yl - 1] / (double)i64min(xl, yl));
}

/* EL:1273-1347 */
static double erp_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                           double g, double *gX, double *gY, double *cost, double *cost_prev,
                           double min_dist) {
  double gx_sum = 0, gy_sum = 0, v;
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  for (int64_t i = 0; i < xl; i++) { v = fabs(X[i] - g); gX[i] = v; gx_sum += v; }
  for (int64_t i = 0; i < yl; i++) { v = fabs(Y[i] - g); gY[i] = v; gy_sum += v; }
  for (int64_t i = 0; i < i64min(yl, max_len); i++) cost_prev[i] = gy_sum;
  if (max_len < yl) cost_prev[max_len] = gy_sum;
  for (int64_t i = 0; i < xl; i++) {
    int64_t j_start = i64max(0, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    if (j_start > 0) cost[j_start - 1] = 0;
    double min_cost = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      double x = cost_prev[j], y, z;
      if (j > 0) { y = cost_prev[j - 1]; z = cost[j - 1]; }
      else { y = (i == 0) ? 0 : gx_sum; z = gx_sum; }
      v = fabs(X[i] - Y[j]);
      cost[j] = dmin(y + v, dmin(x + gX[i], z + gY[j]));
      if (cost[j] < min_cost) min_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = 0;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return cost_prev[yl - 1];
}

/* EL:1437-1497 */
static double edr_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                           double epsilon, double *cost, double *cost_prev, double min_dist) {
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  for (int64_t i = 0; i < i64min(yl, max_len); i++) cost_prev[i] = 0;
  if (max_len < yl) cost_prev[max_len] = 0;
  for (int64_t i = 0; i < xl; i++) {
    int64_t j_start = i64max(0, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    if (j_start > 0) cost[j_start - 1] = 0;
    double min_cost = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      double x = cost_prev[j], y, z;
      if (j > 0) { y = cost_prev[j - 1]; z = cost[j - 1]; }
      else { y = 0; z = 0; }
      double v = fabs(X[i] - Y[j]);
      cost[j] = dmin(dmin(y + (v < epsilon ? 0 : 1), x + 1), z + 1);
      if (cost[j] < min_cost) min_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = 0;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return cost_prev[yl - 1] / (double)i64max(xl, yl);
}

/* EL:1583-1587 -- the arguments are C floats in the reference */
static inline double msm_cost(float x, float y, float z, float c) {
  if ((y <= x && x <= z) || (y >= x && x >= z)) return c;
  return c + dmin(fabs(x - y), fabs(x - z));
}

/* EL:1592-1647.  cost[j_start-1] is deliberately NOT reset (stale read);
 * we reproduce it by using the same two-buffer scheme. */
static double msm_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                           double c, double *cost, double *cost_prev, double *cost_y,
                           double min_dist) {
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  cost_prev[0] = fabs(X[0] - Y[0]);
  for (int64_t i = 1; i < i64min(yl, max_len); i++)
    cost_prev[i] = cost_prev[i - 1] + msm_cost((float)Y[i], (float)Y[i - 1], (float)X[0], (float)c);
  {
    int64_t i = max_len;
    if (i < yl)
      cost_prev[i] = cost_prev[i - 1] + msm_cost((float)Y[i], (float)Y[i - 1], (float)X[0], (float)c);
  }
  cost_y[0] = cost_prev[0];
  for (int64_t i = 1; i < xl; i++)
    cost_y[i] = cost_y[i - 1] + msm_cost((float)X[i], (float)X[i - 1], (float)Y[0], (float)c);
  for (int64_t i = 1; i < xl; i++) {
    int64_t j_start = i64max(1, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    cost[0] = cost_y[i];
    double min_cost = cost[0];
    for (int64_t j = j_start; j < j_stop; j++) {
      double a = cost_prev[j - 1] + fabs(X[i] - Y[j]);
      double b = cost_prev[j] + msm_cost((float)X[i], (float)X[i - 1], (float)Y[j], (float)c);
      double d = cost[j - 1] + msm_cost((float)Y[j], (float)X[i], (float)Y[j - 1], (float)c);
      cost[j] = dmin(dmin(a, b), d);
      if (cost[j] < min_cost) min_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = 0;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return cost_prev[yl - 1];
}

/* EL:1733-1829 */
static double twe_distance(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r,
                           double penalty, double stiffness, double *cost, double *cost_prev,
                           double min_dist) {
  int64_t max_len = i64max(0, yl - xl) + r;
  int64_t min_len = i64max(0, xl - yl);
  for (int64_t i = 0; i < i64min(yl, max_len); i++) cost_prev[i] = INFINITY;
  if (max_len < yl) cost_prev[max_len] = INFINITY;
  penalty = penalty + stiffness;
  for (int64_t i = 0; i < xl; i++) {
    int64_t j_start = i64max(0, i - min_len - r + 1);
    int64_t j_stop = i64min(yl, i + max_len);
    if (j_start > 0) cost[j_start - 1] = 0;
    double min_cost = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      double up = cost_prev[j], left, up_left, x, y;
      if (j == 0) { left = INFINITY; up_left = (i == 0) ? 0 : INFINITY; }
      else { left = cost[j - 1]; up_left = cost_prev[j - 1]; }
      x = (i == 0) ? 0 : X[i - 1];
      y = X[i];
      double del_x = up + fabs(x - y) + penalty;
      x = (j == 0) ? 0 : Y[j - 1];
      y = Y[j];
      double del_y = left + fabs(x - y) + penalty;
      x = (i == 0) ? 0 : X[i - 1];
      y = (j == 0) ? 0 : Y[j - 1];
      double match = up_left + fabs(X[i] - Y[j]) + fabs(x - y) + stiffness * 2 * (double)llabs(i - j);
      cost[j] = dmin(dmin(del_x, del_y), match);
      if (cost[j] < min_cost) min_cost = cost[j];
    }
    if (min_cost > min_dist) return INFINITY;
    if (j_stop < yl) cost[j_stop] = 0;
    double *t = cost; cost = cost_prev; cost_prev = t;
  }
  return cost_prev[yl - 1];
}

/* ------------------------------------------------------------------------------------------
 * Metric object: mirrors `cdef class Metric` reset/distance/eadistance (CD:694-761) for the
 * elastic classes EL:3126-4082.  One instance per worker thread (the reference deep-copies
 * the metric per joblib task, CD:1171).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int metric;
  orc_params p;
  int64_t Tx, Ty;
  double *cost, *cost_prev, *aux_x, *aux_y, *weights;
  const double *std_x, *std_y; /* shared, per sample (EDR default epsilon) */
} orc_metric;


static void metric_init(orc_metric *m, int metric, const orc_params *p, int64_t Tx, int64_t Ty,
                        const double *std_x, const double *std_y) {
  int64_t n = i64max(Tx, Ty);
  m->metric = metric; m->p = *p; m->Tx = Tx; m->Ty = Ty;
  m->cost = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  m->cost_prev = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  m->aux_x = (double *)malloc(sizeof(double) * (size_t)(Tx + 1));
  m->aux_y = (double *)malloc(sizeof(double) * (size_t)(Ty + 1));
  m->weights = NULL;
  m->std_x = std_x; m->std_y = std_y;
  if (metric == ORC_WDTW || metric == ORC_WLCSS) {
    m->weights = (double *)malloc(sizeof(double) * (size_t)n);
    orc_weights(p->g, n, m->weights);
  } else if (metric == ORC_WDDTW) {
    int64_t nn = i64max(Tx - 2, Ty - 2);
    if (nn > 0) { m->weights = (double *)malloc(sizeof(double) * (size_t)nn); orc_weights(p->g, nn, m->weights); }
  }
}
static void metric_free(orc_metric *m) {
  free(m->cost); free(m->cost_prev); free(m->aux_x); free(m->aux_y); free(m->weights);
}

/* `ea` = 0: Metric.distance; `ea` = 1: Metric.eadistance -- returns the raw distance the
 * reference compares against *min_dist (INFINITY when abandoned); *valid=0 for ddtw's
 * "return 0/False" early exit (EL:3297-3298). */
static double metric_eval(orc_metric *m, const double *x, int64_t ix, const double *y, int64_t iy,
                          int ea, double md, int *valid) {
  int64_t Tx = m->Tx, Ty = m->Ty;
  int64_t Tmin = i64min(Tx, Ty), Tmax = i64max(Tx, Ty);
  int64_t R = orc_compute_r(Tmin, m->p.r);
  *valid = 1;
  switch (m->metric) {
    case ORC_DTW: case ORC_WDTW:
      return sqrt(dtw_distance(x, Tx, y, Ty, R, m->cost, m->cost_prev, m->weights, ea ? md * md : INFINITY));
    case ORC_ADTW:
      return sqrt(adtw_distance(x, Tx, y, Ty, R, m->cost, m->cost_prev, m->p.p, ea ? md * md : INFINITY));
    case ORC_DDTW: case ORC_WDDTW: {
      if (Tmin < 3) { *valid = 0; return 0.0; }
      orc_average_slope(x, Tx, m->aux_x);
      orc_average_slope(y, Ty, m->aux_y);
      /* EL:3280 uses the original lengths for R, EL:3308 (eadistance) the derivative lengths */
      int64_t Rd = ea ? orc_compute_r(i64min(Tx - 2, Ty - 2), m->p.r) : R;
      return sqrt(dtw_distance(m->aux_x, Tx - 2, m->aux_y, Ty - 2, Rd, m->cost, m->cost_prev, m->weights,
                               ea ? md * md : INFINITY));
    }
    case ORC_LCSS: case ORC_WLCSS: {
      double t = INFINITY;
      if (ea && !isinf(md)) t = (double)Tmin - md * (double)Tmin; /* EL:3526-3528 */
      return lcss_distance(x, Tx, y, Ty, R, m->p.epsilon, m->cost, m->cost_prev, m->weights, t);
    }
    case ORC_ERP:
      return erp_distance(x, Tx, y, Ty, R, m->p.g, m->aux_x, m->aux_y, m->cost, m->cost_prev, ea ? md : INFINITY);
    case ORC_EDR: {
      double eps = m->p.epsilon;
      if (isnan(eps)) eps = dmax(m->std_x[ix], m->std_y[iy]) / 4.0; /* EL:3762-3766 */
      int64_t Re = ea ? orc_compute_r(Tx, m->p.r) : R;            /* EL:3833 (X twice) */
      return edr_distance(x, Tx, y, Ty, Re, eps, m->cost, m->cost_prev, ea ? md * (double)Tmax : INFINITY);
    }
    case ORC_MSM:
      return msm_distance(x, Tx, y, Ty, R, m->p.c, m->cost, m->cost_prev, m->aux_x, ea ? md : INFINITY);
    case ORC_TWE:
      return twe_distance(x, Tx, y, Ty, R, m->p.penalty, m->p.stiffness, m->cost, m->cost_prev, ea ? md : INFINITY);
  }
  *valid = 0;
  return NAN;
}

static double *sample_std(int metric, const orc_params *p, const double *x, int64_t n, int64_t T, int64_t stride) {
  if (metric != ORC_EDR || !isnan(p->epsilon)) return NULL;
  double *s = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++) s[i] = orc_std(x + i * stride, T); /* EL:3743-3752 */
  return s;
}

/* utils/_parallel.py:7-23 contiguous row blocks */
static void row_block(int64_t n, int nb, int b, int64_t *lo, int64_t *hi) {
  int64_t bs = n / nb, ov = n % nb;
  *lo = b * bs + (b < ov ? b : ov);
  *hi = *lo + bs + (b < ov ? 1 : 0);
}

static int pick_threads(int nthreads, int64_t n_work) {
  if (nthreads <= 0) {
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    nthreads = nc > 0 ? (int)nc : 1;
  }
  if (nthreads > n_work) nthreads = (int)(n_work > 0 ? n_work : 1);
  return nthreads;
}

/* ---- threading: one worker per contiguous row block, like joblib threads (CD:1193-1203) ---- */
typedef struct orc_job {
  int kind; /* 0 pairwise, 1 self, 2 paired, 3 argmin */
  int metric; const orc_params *p;
  const double *x, *y; int64_t nx, ny, Tx, Ty, xs, ys;
  const double *sx, *sy;
  double *out;
  int64_t k; const double *lower_bound; int64_t *out_idx;
  int nb;
} orc_job;
typedef struct { const orc_job *job; int b; } orc_task;

static void job_block(const orc_job *J, int b);
static void *task_main(void *arg) { orc_task *t = (orc_task *)arg; job_block(t->job, t->b); return NULL; }
static void run_job(orc_job *J) {
  if (J->nb <= 1) { J->nb = 1; job_block(J, 0); return; }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)J->nb);
  orc_task *ts = (orc_task *)malloc(sizeof(orc_task) * (size_t)J->nb);
  for (int b = 0; b < J->nb; b++) { ts[b].job = J; ts[b].b = b; pthread_create(&th[b], NULL, task_main, &ts[b]); }
  for (int b = 0; b < J->nb; b++) pthread_join(th[b], NULL);
  free(th); free(ts);
}

/* MI:18-107 bounded max-heap of the k smallest (value, index) */
typedef struct { int64_t index; double value; } heap_el;
static void heap_shift_down(heap_el *h, int64_t startpos, int64_t pos) {
  heap_el ne = h[pos];
  while (pos > startpos) {
    int64_t parent = (pos - 1) >> 1;
    if (ne.value > h[parent].value) { h[pos] = h[parent]; pos = parent; continue; }
    break;
  }
  h[pos] = ne;
}
static void heap_shift_up(heap_el *h, int64_t pos, int64_t endpos) {
  int64_t startpos = pos;
  heap_el ne = h[pos];
  int64_t child = 2 * pos + 1;
  while (child < endpos) {
    int64_t right = child + 1;
    if (right < endpos && h[child].value < h[right].value) child = right;
    h[pos] = h[child];
    pos = child;
    child = 2 * pos + 1;
  }
  h[pos] = ne;
  heap_shift_down(h, startpos, pos);
}
typedef struct { heap_el *h; int64_t n, cap; } orc_heap;
static void heap_push(orc_heap *hp, int64_t index, double value) {
  heap_el e; e.index = index; e.value = value;
  if (hp->n == 0) { hp->h[0] = e; hp->n = 1; }
  else if (hp->n < hp->cap) { hp->h[hp->n] = e; hp->n++; heap_shift_down(hp->h, 0, hp->n - 1); }
  else if (hp->h[0].value > e.value) { hp->h[0] = e; heap_shift_up(hp->h, 0, hp->cap); }
}

/* Exposed for host-side replay tests: push (index[i], value[i]) in order, dump heap array. */
void orc_heap_replay(int64_t k, int64_t n, const int64_t *index, const double *value, int64_t *out_idx,
                     double *out_val, int64_t *out_n) {
  orc_heap hp; hp.h = (heap_el *)calloc((size_t)k, sizeof(heap_el)); hp.n = 0; hp.cap = k;
  for (int64_t i = 0; i < n; i++) heap_push(&hp, index[i], value[i]);
  for (int64_t i = 0; i < k; i++) { out_idx[i] = hp.h[i].index; out_val[i] = hp.h[i].value; }
  *out_n = hp.n;
  free(hp.h);
}

static void job_block(const orc_job *J, int b) {
  int64_t lo, hi, n_work = J->nx;
  row_block(n_work, J->nb, b, &lo, &hi);
  orc_metric m;
  int valid;
  if (J->kind == 0) {        /* CD:1144-1205 */
    metric_init(&m, J->metric, J->p, J->Tx, J->Ty, J->sx, J->sy);
    for (int64_t i = lo; i < hi; i++)
      for (int64_t j = 0; j < J->ny; j++)
        J->out[i * J->ny + j] = metric_eval(&m, J->x + i * J->xs, i, J->y + j * J->ys, j, 0, INFINITY, &valid);
  } else if (J->kind == 1) { /* CD:1208-1267: upper triangle computed, mirrored, zero diagonal */
    metric_init(&m, J->metric, J->p, J->Tx, J->Tx, J->sx, J->sx);
    for (int64_t i = lo; i < hi; i++)
      for (int64_t j = i + 1; j < J->nx; j++) {
        double d = metric_eval(&m, J->x + i * J->xs, i, J->x + j * J->xs, j, 0, INFINITY, &valid);
        J->out[i * J->nx + j] = d; J->out[j * J->nx + i] = d;
      }
  } else if (J->kind == 2) { /* CD:1597-1652: first operand = the USER's y (argument swap) */
    metric_init(&m, J->metric, J->p, J->Ty, J->Tx, J->sy, J->sx);
    for (int64_t i = lo; i < hi; i++)
      J->out[i] = metric_eval(&m, J->y + i * J->ys, i, J->x + i * J->xs, i, 0, INFINITY, &valid);
  } else {                   /* CD:1270-1378 */
    metric_init(&m, J->metric, J->p, J->Tx, J->Ty, J->sx, J->sy);
    orc_heap hp; hp.h = (heap_el *)calloc((size_t)J->k, sizeof(heap_el)); hp.cap = J->k;
    for (int64_t i = lo; i < hi; i++) {
      double distance = INFINITY;
      hp.n = 0;
      for (int64_t j = 0; j < J->ny; j++) {
        if (J->lower_bound && J->lower_bound[i * J->ny + j] >= distance) continue;
        double d = metric_eval(&m, J->x + i * J->xs, i, J->y + j * J->ys, j, 1, distance, &valid);
        if (valid && d < distance) {
          heap_push(&hp, j, d);
          distance = (hp.n == hp.cap) ? hp.h[0].value : INFINITY;
        }
      }
      for (int64_t j = 0; j < J->k; j++) { J->out_idx[i * J->k + j] = hp.h[j].index; J->out[i * J->k + j] = hp.h[j].value; }
    }
    free(hp.h);
  }
  metric_free(&m);
}

int orc_pairwise(int metric, const orc_params *p, const double *x, int64_t nx, int64_t Tx, int64_t xs,
                 const double *y, int64_t ny, int64_t Ty, int64_t ys, double *out, int nthreads) {
  if (metric < 0 || metric >= ORC_N_METRICS || Tx < 1 || Ty < 1) return 1;
  orc_job J; memset(&J, 0, sizeof J);
  J.kind = 0; J.metric = metric; J.p = p; J.x = x; J.y = y; J.nx = nx; J.ny = ny; J.Tx = Tx; J.Ty = Ty; J.xs = xs; J.ys = ys;
  double *sx = sample_std(metric, p, x, nx, Tx, xs), *sy = sample_std(metric, p, y, ny, Ty, ys);
  J.sx = sx; J.sy = sy; J.out = out; J.nb = pick_threads(nthreads, nx);
  run_job(&J);
  free(sx); free(sy);
  return 0;
}

int orc_pairwise_self(int metric, const orc_params *p, const double *x, int64_t n, int64_t T, int64_t xs,
                      double *out, int nthreads) {
  if (metric < 0 || metric >= ORC_N_METRICS || T < 1) return 1;
  orc_job J; memset(&J, 0, sizeof J);
  J.kind = 1; J.metric = metric; J.p = p; J.x = x; J.nx = n; J.Tx = T; J.xs = xs;
  double *sx = sample_std(metric, p, x, n, T, xs);
  J.sx = sx; J.out = out; J.nb = pick_threads(nthreads, n);
  for (int64_t i = 0; i < n * n; i++) out[i] = 0.0;
  run_job(&J);
  free(sx);
  return 0;
}

/* `x`/`y` are the USER's x and y; the reference's swap is applied inside (kind 2). */
int orc_paired(int metric, const orc_params *p, const double *x, int64_t n, int64_t Tx, int64_t xs,
               const double *y, int64_t Ty, int64_t ys, double *out, int nthreads) {
  if (metric < 0 || metric >= ORC_N_METRICS || Tx < 1 || Ty < 1) return 1;
  orc_job J; memset(&J, 0, sizeof J);
  J.kind = 2; J.metric = metric; J.p = p; J.x = x; J.y = y; J.nx = n; J.ny = n; J.Tx = Tx; J.Ty = Ty; J.xs = xs; J.ys = ys;
  double *sx = sample_std(metric, p, x, n, Tx, xs), *sy = sample_std(metric, p, y, n, Ty, ys);
  J.sx = sx; J.sy = sy; J.out = out; J.nb = pick_threads(nthreads, n);
  run_job(&J);
  free(sx); free(sy);
  return 0;
}

/* lower_bound may be NULL (else nx*ny row-major). Output in the reference heap's array order. */
int orc_argmin(int metric, const orc_params *p, const double *x, int64_t nx, int64_t Tx, int64_t xs,
               const double *y, int64_t ny, int64_t Ty, int64_t ys, int64_t k, const double *lower_bound,
               int64_t *out_idx, double *out_dist, int nthreads) {
  if (metric < 0 || metric >= ORC_N_METRICS || Tx < 1 || Ty < 1 || k < 1) return 1;
  orc_job J; memset(&J, 0, sizeof J);
  J.kind = 3; J.metric = metric; J.p = p; J.x = x; J.y = y; J.nx = nx; J.ny = ny; J.Tx = Tx; J.Ty = Ty; J.xs = xs; J.ys = ys;
  double *sx = sample_std(metric, p, x, nx, Tx, xs), *sy = sample_std(metric, p, y, ny, Ty, ys);
  J.sx = sx; J.sy = sy; J.out = out_dist; J.out_idx = out_idx; J.k = k; J.lower_bound = lower_bound;
  J.nb = pick_threads(nthreads, nx);
  run_job(&J);
  free(sx); free(sy);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * DTW lower bounds (LB:198-432, EL:98-151, 228-260, 1076-1115, distance/dtw.py:38-40)
 * ---------------------------------------------------------------------------------------- */

/* distance/dtw.py:38-40 / EL:1917-1921 */
int64_t orc_compute_warp_width(int64_t length, double r) {
  if (r == 1) return length - 1;
  return (int64_t)floor((double)length * r);
}

/* EL:98-151 find_min_max semantics: lower[k] = min T[k-w..k+w], upper[k] = max (clipped).
 * Written directly (O(T w)); min/max of doubles are exact so any evaluation order matches. */
void orc_envelope(const double *t, int64_t n, int64_t w, double *lower, double *upper) {
  for (int64_t k = 0; k < n; k++) {
    int64_t a = i64max(0, k - w), b = i64min(n - 1, k + w);
    double lo = t[a], hi = t[a];
    for (int64_t q = a + 1; q <= b; q++) { lo = dmin(lo, t[q]); hi = dmax(hi, t[q]); }
    lower[k] = lo; upper[k] = hi;
  }
}

/* EL:228-260 cumulative_bound summed + EL:1095-1115: sqrt(sum of squared envelope excess) */
double orc_lb_keogh_one(const double *q, const double *lower, const double *upper, int64_t n) {
  double s = 0;
  for (int64_t i = 0; i < n; i++) {
    double v = q[i], d = 0;
    if (v > upper[i]) { d = v - upper[i]; d = d * d; }
    else if (v < lower[i]) { d = lower[i] - v; d = d * d; }
    s += d;
  }
  return sqrt(s);
}

/* distance/dtw.py:38-40 (_compute_warp_size) + LB:367-369 / LB:410-412: envelope half-width used by
 * DtwKeoghLowerBound: max(floor(T * r), 1), and T - 1 when that equals T. */
int64_t orc_lb_warp_size(int64_t T, double r) {
  int64_t w = (int64_t)floor((double)T * r);
  if (w < 1) w = 1;
  if (w == T) w -= 1;
  return w;
}

/* DtwKeoghLowerBound.fit(X).transform(Q), LB:359-432.  out[i * nx + j] for query i, fitted sample j.
 * kind: 0 both (max of the two directions), 1 left (query against the sample's envelope),
 * 2 right (sample against the query's envelope).  The disabled direction contributes -inf (LB:420-429). */
int orc_lb_keogh_matrix(const double *q, int64_t nq, const double *x, int64_t nx, int64_t T, double r, int kind,
                        double *out) {
  const int64_t w = orc_lb_warp_size(T, r);
  double *xl = (double *)malloc(sizeof(double) * (size_t)nx * T), *xu = (double *)malloc(sizeof(double) * (size_t)nx * T);
  double *ql = (double *)malloc(sizeof(double) * T), *qu = (double *)malloc(sizeof(double) * T);
  if (!xl || !xu || !ql || !qu) return 1;
  for (int64_t j = 0; j < nx; j++) orc_envelope(x + j * T, T, w, xl + j * T, xu + j * T);   /* fit, LB:371-374 */
  for (int64_t i = 0; i < nq; i++) {
    if (kind == 0 || kind == 2) orc_envelope(q + i * T, T, w, ql, qu);                         /* LB:417-418 */
    for (int64_t j = 0; j < nx; j++) {
      double d1 = -INFINITY, d2 = -INFINITY;
      if (kind == 0 || kind == 1) d1 = orc_lb_keogh_one(q + i * T, xl + j * T, xu + j * T, T);
      if (kind == 0 || kind == 2) d2 = orc_lb_keogh_one(x + j * T, ql, qu, T);
      out[i * nx + j] = d1 > d2 ? d1 : d2;                                                     /* LB:431 */
    }
  }
  free(xl); free(xu); free(ql); free(qu);
  return 0;
}

/* DtwKimLowerBound.fit(X).transform(Q), LB:224-311: first/last three points; the SUM of squared terms
 * (the reference applies no sqrt).  x = fitted samples (columns of the result), y = queries. */
static inline double kim_d(double a, double b) { double v = a - b; return v * v; }
int orc_lb_kim_matrix(const double *q, int64_t nq, const double *x, int64_t nx, int64_t T, double *out) {
  for (int64_t i = 0; i < nq; i++) {
    const double *Y = q + i * T;
    for (int64_t j = 0; j < nx; j++) {
      const double *X = x + j * T;
      const double x0 = X[0], x0_ = X[T - 1], y0 = Y[0], y0_ = Y[T - 1];
      double d = kim_d(x0, y0) + kim_d(x0_, y0_);
      if (T > 1) {
        const double x1 = X[1], x1_ = X[T - 2], y1 = Y[1], y1_ = Y[T - 2];
        d += dmin(kim_d(x1, y0), dmin(kim_d(x0, y1), kim_d(x1, y1)));
        d += dmin(kim_d(x1_, y1), dmin(kim_d(x0_, y1_), kim_d(x1_, y1_)));
        if (T > 2) {
          const double x2 = X[2], x2_ = X[T - 3], y2 = Y[2], y2_ = Y[T - 3];
          d += dmin(kim_d(x0, y2), dmin(kim_d(x1, y2), dmin(kim_d(x2, y2), dmin(kim_d(x2, y1), kim_d(x2, y0)))));
          d += dmin(kim_d(x0_, y2_), dmin(kim_d(x1_, y2_), dmin(kim_d(x2_, y2_), dmin(kim_d(x2_, y1_), kim_d(x2_, y0_)))));
        }
      }
      out[i * nx + j] = d;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * DTW alignment matrix, warping path, DBA (test infrastructure like the rest).
 * ------------------------------------------------------------------------------------------ */

/* _elastic.pyx:1011-1073 `_dtw_alignment`: the full (xl, yl) matrix.  The reference allocates it with
 * np.empty and writes only the band and one +inf sentinel on either side of every row; this restatement
 * writes exactly the same cells and leaves the others untouched (callers pre-fill `out`, e.g. with NaN). */
void orc_dtw_alignment(const double *X, int64_t xl, const double *Y, int64_t yl, int64_t r, const double *weights,
                       double *out) {
  double w = 1.0, v;
  const int64_t dy = i64max(0, yl - xl), dx = i64max(0, xl - yl);
  v = X[0] - Y[0];
  if (weights) w = weights[0];
  out[0] = v * v * w;
  for (int64_t i = 1; i < i64min(xl, r + 1); i++) {
    v = X[i] - Y[0];
    if (weights) w = weights[i];
    out[i * yl] = out[(i - 1) * yl] + v * v * w;
  }
  for (int64_t i = 1; i < i64min(yl, dy + r); i++) {
    v = X[0] - Y[i];
    if (weights) w = weights[i];
    out[i] = out[i - 1] + v * v * w;
  }
  if (dy + r < yl) out[dy + r] = INFINITY;
  for (int64_t i = 1; i < xl; i++) {
    const int64_t j_start = i64max(0, i - dx - r + 1), j_stop = i64min(yl, i + dy + r);
    if (j_start > 0) out[i * yl + j_start - 1] = INFINITY;
    for (int64_t j = j_start; j < j_stop; j++) {
      v = X[i] - Y[j];
      const double x = out[(i - 1) * yl + j];
      const double y = j > 0 ? out[i * yl + j - 1] : INFINITY;
      const double z = j > 0 ? out[(i - 1) * yl + j - 1] : INFINITY;
      if (weights) w = weights[llabs(i - j)];
      out[i * yl + j] = dmin(dmin(x, y), z) + v * v * w;
    }
    if (j_stop < yl) out[i * yl + j_stop] = INFINITY;
  }
}

/* distance/dtw.py:385-413 `dtw_mapping`: walk back from the last cell; np.argmin([diag, up, left]) takes the
 * FIRST minimum.  lo[i] / hi[i] = first / last column of the path in row i (a monotone path covers a contiguous
 * run of columns per row; `indicator.nonzero()` lists them row by row, ascending). */
void orc_dtw_path(const double *D, int64_t xl, int64_t yl, int32_t *lo, int32_t *hi) {
  int64_t i = xl - 1, j = yl - 1;
  for (int64_t k = 0; k < xl; k++) { lo[k] = INT32_MAX; hi[k] = -1; }
  while (i > 0 || j > 0) {
    if (j < lo[i]) lo[i] = (int32_t)j;
    if (j > hi[i]) hi[i] = (int32_t)j;
    const double od = (i > 0 && j > 0) ? D[(i - 1) * yl + j - 1] : INFINITY;
    const double ou = i > 0 ? D[(i - 1) * yl + j] : INFINITY;
    const double ol = j > 0 ? D[i * yl + j - 1] : INFINITY;
    int move = 0;
    double best = od;
    if (ou < best) { best = ou; move = 1; }
    if (ol < best) { best = ol; move = 2; }
    if (move == 0) { i--; j--; } else if (move == 1) i--; else j--;
  }
  if (0 < lo[0]) lo[0] = 0;
  if (0 > hi[0]) hi[0] = 0;
}

/* ------------------------------------------------------------------------------------------
 * subsequence search, DTW family + lcss / erp / edr / msm / twe (test infrastructure).
 * lcss_subsequence_distance EL:1186-1224, erp EL:1350-1388, edr EL:1500-1536 (epsilon resolved by the caller: NaN means
 * s_std / 4, EL:2762-2765), msm EL:1650-1686, twe EL:1832-1869: the window's *_distance() is early-abandoned against the
 * running minimum -- for these metrics that CHANGES which windows are accepted (row minima are not monotone), so the scan
 * order is part of the result; they return the minimum itself (no sqrt).
 * dtw_subsequence_distance EL:622-660, adtw_subsequence_distance EL:701-740, ddtw_subsequence_distance EL:780-815,
 * with the per-class conventions of EL:2206-2615: r = _compute_r(s_len, r) from the ORIGINAL subsequence length;
 * wdtw weights over n_timestep (EL:2370-2372), wddtw over n_timestep - 2 (EL:2540-2543); early abandoning against
 * the running minimum, strict `<` (first best window).  Returns sqrt(min) and the window start in *index
 * (left untouched -- as in the reference -- when nothing is accepted or ddtw sees s_len < 3).
 * ------------------------------------------------------------------------------------------ */
double orc_subsequence_distance(int metric, const orc_params *p, const double *S, int64_t s_len, const double *T,
                                int64_t t_len, int64_t *index) {
  const int deriv = (metric == ORC_DDTW || metric == ORC_WDDTW);
  const int64_t r = orc_compute_r(s_len, p->r);
  double *cost = (double *)malloc(sizeof(double) * (size_t)(t_len + 1));
  double *cost_prev = (double *)malloc(sizeof(double) * (size_t)(t_len + 1));
  double *weights = NULL, *S_buffer = NULL, *T_buffer = NULL;
  double *gX = (double *)malloc(sizeof(double) * (size_t)(t_len + 1));
  double *gY = (double *)malloc(sizeof(double) * (size_t)(t_len + 1));
  double min_dist = INFINITY, dist;
  const int64_t length = t_len - s_len + 1;
  const int squared = !(metric == ORC_LCSS || metric == ORC_ERP || metric == ORC_EDR || metric == ORC_MSM || metric == ORC_TWE);
  if (metric == ORC_WDTW) {
    weights = (double *)malloc(sizeof(double) * (size_t)t_len);
    orc_weights(p->g, t_len, weights);
  } else if (metric == ORC_WDDTW) {
    weights = (double *)malloc(sizeof(double) * (size_t)t_len);
    if (t_len - 2 > 0) orc_weights(p->g, t_len - 2, weights);
  }
  if (deriv) {
    if (s_len < 3) { free(cost); free(cost_prev); free(weights); free(gX); free(gY); return 0; }
    S_buffer = (double *)malloc(sizeof(double) * (size_t)t_len);
    T_buffer = (double *)malloc(sizeof(double) * (size_t)t_len);
    orc_average_slope(S, s_len, S_buffer);
  }
  for (int64_t i = 0; i < length; i++) {
    if (deriv) {
      orc_average_slope(T + i, s_len, T_buffer);
      dist = dtw_distance(S_buffer, s_len - 2, T_buffer, s_len - 2, r, cost, cost_prev, weights, min_dist);
    } else if (metric == ORC_ADTW) {
      dist = adtw_distance(S, s_len, T + i, s_len, r, cost, cost_prev, p->p, min_dist);
    } else if (metric == ORC_LCSS) { /* EL:1186-1224; threshold s_len - min_dist * s_len (min(s_len, t_len) = s_len) */
      dist = lcss_distance(S, s_len, T + i, s_len, r, p->epsilon, cost, cost_prev, NULL,
                           isinf(min_dist) ? min_dist : (double)i64min(s_len, t_len) - min_dist * (double)i64min(s_len, t_len));
    } else if (metric == ORC_ERP) { /* EL:1350-1388 */
      dist = erp_distance(S, s_len, T + i, s_len, r, p->g, gX, gY, cost, cost_prev, min_dist);
    } else if (metric == ORC_EDR) { /* EL:1500-1536: threshold min_dist * max(s_len, t_len) -- the SERIES length */
      dist = edr_distance(S, s_len, T + i, s_len, r, p->epsilon, cost, cost_prev, min_dist * (double)i64max(s_len, t_len));
    } else if (metric == ORC_MSM) { /* EL:1650-1686 */
      dist = msm_distance(S, s_len, T + i, s_len, r, p->c, cost, cost_prev, gX, min_dist);
    } else if (metric == ORC_TWE) { /* EL:1832-1869 */
      dist = twe_distance(S, s_len, T + i, s_len, r, p->penalty, p->stiffness, cost, cost_prev, min_dist);
    } else {
      dist = dtw_distance(S, s_len, T + i, s_len, r, cost, cost_prev, weights, min_dist);
    }
    if (dist < min_dist) {
      if (index) *index = i;
      min_dist = dist;
    }
  }
  free(cost); free(cost_prev); free(weights); free(S_buffer); free(T_buffer); free(gX); free(gY);
  return squared ? sqrt(min_dist) : min_dist;
}

static inline double sq_dist(double x, double y) { const double s = x - y; return s * s; }

/* EL:158-225 `constant_lower_bound` (LB_Kim over the first / last three points of the z-normalised series), term by
 * term as written there -- including the third term, which lists dist(t_y1, s_y1) twice and never dist(t_y1, s_y0):
 * the value can therefore EXCEED the DTW distance, and since the scan skips a window when it is >= the running minimum
 * (EL:413), it is part of the observable result and has to be reproduced.  The staged early returns of the reference
 * return a partial sum that is already >= best_dist; the terms are non-negative, so "skipped" <=> full sum >= best_dist. */
static double ucr_lb_kim(const double *S, double s_mean, double s_std, const double *T, double t_mean, double t_std,
                         int64_t length) {
  if (t_std == 0) return 0;
  const double t_x0 = (T[0] - t_mean) / t_std, t_y0 = (T[length - 1] - t_mean) / t_std;
  const double s_x0 = (S[0] - s_mean) / s_std, s_y0 = (S[length - 1] - s_mean) / s_std;
  double min_dist = sq_dist(t_x0, s_x0) + sq_dist(t_y0, s_y0);
  const double t_x1 = (T[1] - t_mean) / t_std, s_x1 = (S[1] - s_mean) / s_std;
  min_dist += dmin(dmin(sq_dist(t_x1, s_x0), sq_dist(t_x0, s_x1)), sq_dist(t_x1, s_x1));
  const double t_y1 = (T[length - 2] - t_mean) / t_std, s_y1 = (S[length - 2] - s_mean) / s_std;
  min_dist += dmin(dmin(sq_dist(t_y1, s_y1), sq_dist(t_y0, s_y1)), sq_dist(t_y1, s_y1));
  const double t_x2 = (T[2] - t_mean) / t_std, s_x2 = (S[2] - s_mean) / s_std;
  min_dist += dmin(dmin(sq_dist(t_x0, s_x2), dmin(dmin(sq_dist(t_x1, s_x2), sq_dist(t_x2, s_x2)), sq_dist(t_x2, s_x1))),
                   sq_dist(t_x2, s_x0));
  const double t_y2 = (T[length - 3] - t_mean) / t_std, s_y2 = (S[length - 3] - s_mean) / s_std;
  min_dist += dmin(dmin(sq_dist(t_y0, s_y2), dmin(dmin(sq_dist(t_y1, s_y2), sq_dist(t_y2, s_y2)), sq_dist(t_y2, s_y1))),
                   sq_dist(t_y2, s_y0));
  return min_dist;
}

/* scaled_dtw (UCR suite) subsequence search, ScaledDtwSubsequenceMetric EL:1928-2060.
 * Follows scaled_dtw_subsequence_distance EL:353-482 for the window statistics (running ex / ex2, the oldest sample
 * subtracted again, std = 1 when the variance is not positive) and the selection rule (`dist < min_dist`, first best
 * window, sqrt of the minimum), and inner_scaled_dtw_subsequence_distance EL:263-345 for the DP (band |i - j| <= r with
 * r = _compute_warp_width(s_len, r) EL:1917-1921, v = (S[i] - s_mean) / s_std - (X[j] - mean) / std).  The LB_Kim
 * prefilter IS restated (ucr_lb_kim above: it is not a valid bound and changes results); the LB_Keogh bounds and the
 * cumulative-bound abandoning (EL:424-470, 333-334) are valid lower bounds that only skip or cut short windows whose
 * distance cannot be below the running minimum, and are not restated.  s_std == 0 is passed as 1 by the caller
 * (_cdistance.pyx:370).  Needs s_len >= 3 (the reference reads S[1], S[2] unconditionally). */
double orc_scaled_dtw_subsequence(const double *S, int64_t s_len, double s_mean, double s_std, const double *T,
                                  int64_t t_len, double rfrac, int64_t *index) {
  const int64_t r = (rfrac == 1.0) ? s_len - 1 : (int64_t)floor((double)s_len * rfrac);
  double *cost = (double *)malloc(sizeof(double) * (size_t)(2 * r + 2));
  double *cost_prev = (double *)malloc(sizeof(double) * (size_t)(2 * r + 2));
  double ex = 0, ex2 = 0, min_dist = INFINITY;
  for (int64_t t = 0; t < t_len; t++) {
    const double cur = T[t];
    ex += cur;
    ex2 += cur * cur;
    if (t >= s_len - 1) {
      const int64_t I = t - (s_len - 1);
      const double *X = T + I;
      const double mean = ex / (double)s_len;
      const double tmp = ex2 / (double)s_len - mean * mean;
      const double std = tmp > 0 ? sqrt(tmp) : 1.0;
      if (!(ucr_lb_kim(S, s_mean, s_std, X, mean, std, s_len) < min_dist)) { ex -= X[0]; ex2 -= X[0] * X[0]; continue; }
      double *c = cost, *cp = cost_prev;
      int64_t k = 0;
      for (int64_t i = 0; i < 2 * r + 1; i++) { c[i] = INFINITY; cp[i] = INFINITY; }
      for (int64_t i = 0; i < s_len; i++) {
        k = i64max(0, r - i);
        for (int64_t j = i64max(0, i - r); j < i64min(s_len, i + r + 1); j++) {
          double v = (S[i] - s_mean) / s_std;
          v -= (X[j] - mean) / std;
          if (i == 0 && j == 0) c[k] = v * v;
          else {
            const double y = (j - 1 < 0 || k - 1 < 0) ? INFINITY : c[k - 1];
            const double x = (i - 1 < 0 || k + 1 > 2 * r) ? INFINITY : cp[k + 1];
            const double z = (i - 1 < 0 || j - 1 < 0) ? INFINITY : cp[k];
            c[k] = dmin(dmin(x, y), z) + v * v;
          }
          k++;
        }
        double *tswap = c; c = cp; cp = tswap;
      }
      const double dist = cp[k - 1];
      if (dist < min_dist) { if (index) *index = I; min_dist = dist; }
      ex -= X[0];
      ex2 -= X[0] * X[0];
    }
  }
  free(cost); free(cost_prev);
  return sqrt(min_dist);
}

/* ------------------------------------------------------------------------------------------
 * the generic scaled subsequence metrics `scaled_<metric>` = ScaledSubsequenceMetricWrap(Metric),
 * CD:470-551 (`_distance`): the subsequence is z-normalised with (s_mean, s_std) handed in by the caller
 * (np.mean / np.std, std <= 1e-13 -> 0 -> 1.0, CD:453-467, 283-298), every window with the running IncStats
 * (ST:45-93: Welford add / remove, variance < 1e-13 -> 0 -> std 1), then `wrap._eadistance(s_buffer, window, &min_dist)`
 * decides (strict <, early abandoning against the running minimum).  `wrap.reset(X, X)` sizes the weight vectors from the
 * SERIES length (wdtw EL:3334-3341: t_len; wddtw EL:3415-3428: t_len - 2).  EDR's default epsilon comes from
 * fast_mean_std of the two normalised buffers (EL:3856-3861).  ddtw / wddtw with s_len < 3: nothing is accepted (EL:3297).
 * ------------------------------------------------------------------------------------------ */
typedef struct { double mean, n_samples, sum_square, sum; } inc_stats;
static void inc_add(inc_stats *s, double w, double v) {
  s->n_samples += w;
  double next_m = s->mean + (v - s->mean) / s->n_samples;
  s->sum_square += (v - s->mean) * (v - next_m);
  s->mean = next_m;
  s->sum += w * v;
}
static void inc_remove(inc_stats *s, double w, double v) {
  if (s->n_samples == 1.0) { s->n_samples = 0.0; s->mean = 0.0; s->sum_square = 0.0; }
  else {
    double old_m = (s->n_samples * s->mean - v) / (s->n_samples - w);
    s->sum_square -= (v - s->mean) * (v - old_m);
    s->mean = old_m;
    s->n_samples -= w;
  }
  s->sum -= w * v;
}
static double inc_variance(const inc_stats *s) {
  if (s->n_samples <= 1) return 0;
  double var = s->sum_square / s->n_samples;
  if (var < 1e-13) var = 0.0;
  return var;
}

/* mean / std of every window as the wrap computes them (exported for tests of the device statistics kernel) */
void orc_inc_window_stats(const double *x, int64_t t_len, int64_t s_len, double *mean, double *std) {
  inc_stats st = {0, 0, 0, 0};
  for (int64_t i = 0; i < s_len - 1; i++) inc_add(&st, 1.0, x[i]);
  for (int64_t i = 0; i < t_len - s_len + 1; i++) {
    inc_add(&st, 1.0, x[i + s_len - 1]);
    double sd = inc_variance(&st);
    sd = (sd == 0.0) ? 1.0 : sqrt(sd);
    mean[i] = st.mean; std[i] = sd;
    inc_remove(&st, 1.0, x[i]);
  }
}

/* Metric._eadistance(sb, xb) of two equal-length buffers against the running bound md (EL:3188-3213, 3289-3320, 3376-3401,
 * 3508-3536, 3638-3665, 3847-3881, 3953-3979, 4053-4079): the value the reference compares with md (sqrt domain for the DTW
 * family); *valid = 0 for ddtw / wddtw below three samples (returns False).  weights: over the series length (reset(X, X)). */
typedef struct { double *cost, *cost_prev, *a1, *a2, *weights; } ea_scratch;
static double ea_eval(int metric, const orc_params *p, const double *sb, const double *xb, int64_t s_len, double md,
                      ea_scratch *w, int *valid) {
  const int64_t r = orc_compute_r(s_len, p->r);
  *valid = 1;
  switch (metric) {
    case ORC_DTW: case ORC_WDTW:
      return sqrt(dtw_distance(sb, s_len, xb, s_len, r, w->cost, w->cost_prev, w->weights, md * md));
    case ORC_ADTW:
      return sqrt(adtw_distance(sb, s_len, xb, s_len, r, w->cost, w->cost_prev, p->p, md * md));
    case ORC_DDTW: case ORC_WDDTW:
      if (s_len < 3) { *valid = 0; return 0; }
      orc_average_slope(sb, s_len, w->a1); orc_average_slope(xb, s_len, w->a2);
      return sqrt(dtw_distance(w->a1, s_len - 2, w->a2, s_len - 2, orc_compute_r(s_len - 2, p->r), w->cost, w->cost_prev,
                               w->weights, md * md));
    case ORC_LCSS: case ORC_WLCSS:
      return lcss_distance(sb, s_len, xb, s_len, r, p->epsilon, w->cost, w->cost_prev, metric == ORC_WLCSS ? w->weights : NULL,
                           isinf(md) ? INFINITY : (double)s_len - md * (double)s_len);
    case ORC_ERP:
      return erp_distance(sb, s_len, xb, s_len, r, p->g, w->a1, w->a2, w->cost, w->cost_prev, md);
    case ORC_EDR: {
      double eps = p->epsilon;
      if (isnan(eps)) eps = dmax(orc_std(sb, s_len), orc_std(xb, s_len)) / 4.0;
      return edr_distance(sb, s_len, xb, s_len, r, eps, w->cost, w->cost_prev, md * (double)s_len);
    }
    case ORC_MSM:
      return msm_distance(sb, s_len, xb, s_len, r, p->c, w->cost, w->cost_prev, w->a1, md);
    case ORC_TWE:
      return twe_distance(sb, s_len, xb, s_len, r, p->penalty, p->stiffness, w->cost, w->cost_prev, md);
  }
  *valid = 0;
  return NAN;
}
static void ea_scratch_init(ea_scratch *w, int metric, const orc_params *p, int64_t t_len) {
  size_t n = (size_t)(t_len + 2);
  w->cost = (double *)malloc(sizeof(double) * n); w->cost_prev = (double *)malloc(sizeof(double) * n);
  w->a1 = (double *)malloc(sizeof(double) * n); w->a2 = (double *)malloc(sizeof(double) * n);
  w->weights = NULL;
  if (metric == ORC_WDTW || metric == ORC_WLCSS) { w->weights = (double *)malloc(sizeof(double) * n); orc_weights(p->g, t_len, w->weights); }
  if (metric == ORC_WDDTW) { w->weights = (double *)malloc(sizeof(double) * n); if (t_len - 2 > 0) orc_weights(p->g, t_len - 2, w->weights); }
}
static void ea_scratch_free(ea_scratch *w) { free(w->cost); free(w->cost_prev); free(w->a1); free(w->a2); free(w->weights); }

/* k = 1: ScaledSubsequenceMetricWrap._distance (CD:494-551).  k >= 1 with out_idx / out_dist: the paired scans of
 * argmin_subsequence_distance, `_ArgminSubsequenceDistance` (scaled == 0, raw buffers) and `_ScaledArgminSubsequenceDistance`
 * (scaled != 0) CD:1380-1548: the running bound is the heap maximum once the k-heap is full (MI:62-107), the result the heap
 * array (heap order; entries >= *n_found are whatever calloc left: the reference reads uninitialised memory there). */
static double subsequence_ea_scan_w(int metric, const orc_params *p, const double *S, int64_t s_len, double s_mean,
                                    double s_std, const double *T, int64_t t_len, int scaled, int64_t k, int64_t *out_idx,
                                    double *out_dist, int64_t *n_found, int64_t *index, int64_t weight_len) {
  size_t n = (size_t)(t_len + 2);
  /* weight tables span the series the metric was reset() with; the dilated profile compares shorter windows (CD:1769) */
  ea_scratch w; ea_scratch_init(&w, metric, p, weight_len > 0 ? weight_len : t_len);
  double *sb = (double *)malloc(sizeof(double) * n), *xb = (double *)malloc(sizeof(double) * n);
  double *mean = (double *)malloc(sizeof(double) * n), *std = (double *)malloc(sizeof(double) * n);
  orc_heap hp; hp.h = (heap_el *)calloc((size_t)k, sizeof(heap_el)); hp.n = 0; hp.cap = k;
  double min_dist = INFINITY;
  if (scaled) {
    for (int64_t i = 0; i < s_len; i++) sb[i] = (S[i] - s_mean) / s_std;
    orc_inc_window_stats(T, t_len, s_len, mean, std);
  } else memcpy(sb, S, sizeof(double) * (size_t)s_len);
  for (int64_t i = 0; i < t_len - s_len + 1; i++) {
    const double *X = T + i;
    if (scaled) { for (int64_t j = 0; j < s_len; j++) xb[j] = (T[i + j] - mean[i]) / std[i]; X = xb; }
    int valid;
    const double dist = ea_eval(metric, p, sb, X, s_len, min_dist, &w, &valid);
    if (valid && dist < min_dist) {
      heap_push(&hp, i, dist);
      if (index) *index = i;
      if (k == 1) min_dist = dist;
    }
    if (k > 1 || !valid) min_dist = (hp.n == hp.cap) ? hp.h[0].value : INFINITY;
  }
  for (int64_t j = 0; j < k && out_idx; j++) { out_idx[j] = hp.h[j].index; out_dist[j] = hp.h[j].value; }
  if (n_found) *n_found = hp.n;
  const double best = (k == 1 && hp.n == 1) ? hp.h[0].value : min_dist;
  free(sb); free(xb); free(mean); free(std); free(hp.h); ea_scratch_free(&w);
  return best;
}

static double subsequence_ea_scan(int metric, const orc_params *p, const double *S, int64_t s_len, double s_mean,
                                  double s_std, const double *T, int64_t t_len, int scaled, int64_t k, int64_t *out_idx,
                                  double *out_dist, int64_t *n_found, int64_t *index) {
  return subsequence_ea_scan_w(metric, p, S, s_len, s_mean, s_std, T, t_len, scaled, k, out_idx, out_dist, n_found, index, 0);
}

double orc_scaled_subsequence_distance(int metric, const orc_params *p, const double *S, int64_t s_len, double s_mean,
                                       double s_std, const double *T, int64_t t_len, int64_t *index) {
  return subsequence_ea_scan(metric, p, S, s_len, s_mean, s_std, T, t_len, 1, 1, NULL, NULL, NULL, index);
}

int64_t orc_argmin_subsequence(int metric, const orc_params *p, const double *S, int64_t s_len, double s_mean, double s_std,
                               const double *T, int64_t t_len, int scaled, int64_t k, int64_t *out_idx, double *out_dist) {
  int64_t n_found = 0;
  subsequence_ea_scan(metric, p, S, s_len, s_mean, s_std, T, t_len, scaled, k, out_idx, out_dist, &n_found, NULL);
  return n_found;
}

/* ------------------------------------------------------------------------------------------
 * `_matches` / `_distance_profile` of the elastic subsequence metrics -- subsequence_match,
 * paired_subsequence_match and distance_profile of _distance.py:732-1080, 1477-1600 end here (CD:338-372: the profile is
 * `_matches` with threshold = +inf).  out[w] = distance of window w when the reference reports it as a match under
 * `threshold`, NaN otherwise; returns the number of matches.  Rules restated:
 *   unscaled dtw / wdtw / adtw / ddtw / wddtw (*_subsequence_matches EL:658-700, 737-779, 821-868): the DP runs against
 *     threshold^2, match iff dist <= threshold^2, reported sqrt(dist);
 *   unscaled lcss / erp / edr / msm / twe (EL:1227-1270, 1391-1434, 1539-1580, 1689-1730, 1872-1914): the DP is abandoned
 *     against s_len - threshold * s_len (lcss, +inf stays +inf), threshold * max(s_len, t_len) (edr), threshold (others);
 *     match iff dist <= threshold;
 *   scaled_dtw (scaled_dtw_matches EL:485-619): windows with LB_Kim >= threshold^2 are skipped (that bound is not valid,
 *     see ucr_lb_kim), match iff dist <= threshold^2, sqrt; the LB_Keogh bounds / cumulative abandoning are valid and are
 *     not restated;
 *   scaled_<metric> wraps (ScaledSubsequenceMetricWrap._matches CD:553-606): Metric._eadistance with *min_dist =
 *     threshold, i.e. match iff dist < threshold (STRICT).
 * ------------------------------------------------------------------------------------------ */
int64_t orc_subsequence_matches(int metric, const orc_params *p, const double *S, int64_t s_len, double s_mean, double s_std,
                                const double *T, int64_t t_len, int scaled, double threshold, double *out) {
  const int64_t nw = t_len - s_len + 1;
  int64_t n_matches = 0;
  size_t n = (size_t)(t_len + 2);
  for (int64_t w = 0; w < nw; w++) out[w] = NAN;
  if (scaled && metric == ORC_DTW) {
    const int64_t r = orc_compute_warp_width(s_len, p->r);
    double *cost = (double *)malloc(sizeof(double) * (size_t)(2 * r + 2));
    double *cost_prev = (double *)malloc(sizeof(double) * (size_t)(2 * r + 2));
    const double thr = threshold * threshold;
    double ex = 0, ex2 = 0;
    for (int64_t t = 0; t < t_len; t++) {
      const double cur = T[t];
      ex += cur; ex2 += cur * cur;
      if (t < s_len - 1) continue;
      const int64_t I = t - (s_len - 1);
      const double *X = T + I;
      const double mean = ex / (double)s_len;
      const double tmp = ex2 / (double)s_len - mean * mean;
      const double std = tmp > 0 ? sqrt(tmp) : 1.0;
      ex -= X[0]; ex2 -= X[0] * X[0];
      if (!(ucr_lb_kim(S, s_mean, s_std, X, mean, std, s_len) < thr)) continue;
      double *c = cost, *cp = cost_prev;
      int64_t k = 0;
      for (int64_t i = 0; i < 2 * r + 1; i++) { c[i] = INFINITY; cp[i] = INFINITY; }
      for (int64_t i = 0; i < s_len; i++) {
        k = i64max(0, r - i);
        for (int64_t j = i64max(0, i - r); j < i64min(s_len, i + r + 1); j++) {
          double v = (S[i] - s_mean) / s_std;
          v -= (X[j] - mean) / std;
          if (i == 0 && j == 0) c[k] = v * v;
          else {
            const double y = (j - 1 < 0 || k - 1 < 0) ? INFINITY : c[k - 1];
            const double x = (i - 1 < 0 || k + 1 > 2 * r) ? INFINITY : cp[k + 1];
            const double z = (i - 1 < 0 || j - 1 < 0) ? INFINITY : cp[k];
            c[k] = dmin(dmin(x, y), z) + v * v;
          }
          k++;
        }
        double *tswap = c; c = cp; cp = tswap;
      }
      const double dist = cp[k - 1];
      if (dist <= thr) { out[I] = sqrt(dist); n_matches++; }
    }
    free(cost); free(cost_prev);
    return n_matches;
  }
  const int deriv = (metric == ORC_DDTW || metric == ORC_WDDTW);
  double *cost = (double *)malloc(sizeof(double) * n), *cost_prev = (double *)malloc(sizeof(double) * n);
  double *sb = (double *)malloc(sizeof(double) * n), *xb = (double *)malloc(sizeof(double) * n);
  double *a1 = (double *)malloc(sizeof(double) * n), *a2 = (double *)malloc(sizeof(double) * n);
  double *weights = NULL, *mean = (double *)malloc(sizeof(double) * n), *std = (double *)malloc(sizeof(double) * n);
  if (metric == ORC_WDTW) { weights = (double *)malloc(sizeof(double) * n); orc_weights(p->g, t_len, weights); }
  if (metric == ORC_WDDTW) { weights = (double *)malloc(sizeof(double) * n); if (t_len - 2 > 0) orc_weights(p->g, t_len - 2, weights); }
  if (scaled) {
    for (int64_t i = 0; i < s_len; i++) sb[i] = (S[i] - s_mean) / s_std;
    orc_inc_window_stats(T, t_len, s_len, mean, std);
  } else {
    memcpy(sb, S, sizeof(double) * (size_t)s_len);
  }
  if (deriv && s_len >= 3) orc_average_slope(sb, s_len, a1);
  const int64_t r = orc_compute_r(s_len, p->r);
  for (int64_t i = 0; i < nw; i++) {
    const double *X = T + i;
    if (scaled) { for (int64_t j = 0; j < s_len; j++) xb[j] = (T[i + j] - mean[i]) / std[i]; X = xb; }
    double dist, thr = threshold;
    int sq = 0;
    switch (metric) {
      case ORC_DTW: case ORC_WDTW:
        sq = 1; thr = threshold * threshold;
        dist = dtw_distance(sb, s_len, X, s_len, r, cost, cost_prev, weights, thr); break;
      case ORC_ADTW:
        sq = 1; thr = threshold * threshold;
        dist = adtw_distance(sb, s_len, X, s_len, r, cost, cost_prev, p->p, thr); break;
      case ORC_DDTW: case ORC_WDDTW:
        if (s_len < 3) continue; /* EL:843-844 (no matches) / EL:3297 (nothing accepted) */
        sq = 1; thr = threshold * threshold;
        orc_average_slope(X, s_len, a2);
        dist = dtw_distance(a1, s_len - 2, a2, s_len - 2, scaled ? orc_compute_r(s_len - 2, p->r) : r, cost, cost_prev, weights, thr);
        break;
      case ORC_LCSS:
        dist = lcss_distance(sb, s_len, X, s_len, r, p->epsilon, cost, cost_prev, NULL,
                             isinf(threshold) ? INFINITY : (double)s_len - threshold * (double)s_len);
        break;
      case ORC_ERP:
        dist = erp_distance(sb, s_len, X, s_len, r, p->g, a1, a2, cost, cost_prev, threshold); break;
      case ORC_EDR: {
        double eps = p->epsilon;
        if (isnan(eps)) eps = dmax(orc_std(sb, s_len), orc_std(X, s_len)) / 4.0; /* scaled wrap only; unscaled: resolved by the caller */
        dist = edr_distance(sb, s_len, X, s_len, r, eps, cost, cost_prev, threshold * (double)(scaled ? s_len : i64max(s_len, t_len)));
        break;
      }
      case ORC_MSM:
        dist = msm_distance(sb, s_len, X, s_len, r, p->c, cost, cost_prev, a2, threshold); break;
      case ORC_TWE:
        dist = twe_distance(sb, s_len, X, s_len, r, p->penalty, p->stiffness, cost, cost_prev, threshold); break;
      default: dist = NAN;
    }
    if (scaled) {
      if (sq) dist = sqrt(dist);
      if (dist < threshold) { out[i] = dist; n_matches++; }
    } else if (dist <= thr) { out[i] = sq ? sqrt(dist) : dist; n_matches++; }
  }
  free(cost); free(cost_prev); free(sb); free(xb); free(a1); free(a2); free(weights); free(mean); free(std);
  return n_matches;
}

/* orc_argmin_subsequence with the weight tables of wdtw / wddtw sized for a series of `weight_len` points (wddtw: weight_len - 2):
 * what `_dilated_distance_profile` sees after metric.reset(X, X) (CD:1769) when it compares windows shorter than the series. */
int64_t orc_argmin_subsequence_w(int metric, const orc_params *p, const double *S, int64_t s_len, double s_mean, double s_std,
                                 const double *T, int64_t t_len, int scaled, int64_t k, int64_t weight_len, int64_t *out_idx,
                                 double *out_dist) {
  int64_t n_found = 0;
  subsequence_ea_scan_w(metric, p, S, s_len, s_mean, s_std, T, t_len, scaled, k, out_idx, out_dist, &n_found, NULL, weight_len);
  return n_found;
}
