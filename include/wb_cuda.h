/*
 * wb_cuda.h -- C ABI of libwbcuda.so: B200 (sm_100a) kernels for wildboar's elastic-distance
 * hot path.  Plain pointers and sizes only; no exceptions cross this boundary.
 *
 * Each entry point replaces one Cython batch driver of the reference
 * (src/wildboar/distance/_cdistance.pyx, "CD") specialised to the elastic Metric classes of
 * src/wildboar/distance/_elastic.pyx ("EL"); the reference-side binding is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - return 0 on success, non-zero on error; wb_cuda_last_error() returns a thread-local
 *     message for the last failing call on this thread.
 *   - host entry points take caller-owned HOST buffers (float64, last axis contiguous,
 *     `*_stride` = elements between consecutive samples, dim already selected -- the
 *     reference's TSArray view `&X[i, dim, 0]`, utils/__init__.pxd:4); the library stages,
 *     shards rows over `devices` and gathers.  They may be called with the GIL released.
 *   - `*_dev` entry points take DEVICE pointers on the current device and enqueue on `stream`
 *     (a cudaStream_t); they do not synchronise except where stated.
 *   - there is NO CPU fallback: without an sm_100 device every compute call fails.
 */
#ifndef WB_CUDA_H
#define WB_CUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* metric ids == keys of _METRICS (distance/_distance.py:186-205) that are elastic */
enum {
  WB_DTW = 0,   /* DtwMetric                   EL:3126 */
  WB_WDTW = 1,  /* WeightedDtwMetric           EL:3322 */
  WB_DDTW = 2,  /* DerivativeDtwMetric         EL:3228 */
  WB_ADTW = 3,  /* AmercingDtwMetric           EL:3344 */
  WB_LCSS = 4,  /* LcssMetric                  EL:3433 */
  WB_ERP = 5,   /* ErpMetric                   EL:3564 */
  WB_EDR = 6,   /* EdrMetric                   EL:3671 */
  WB_MSM = 7,   /* MsmMetric                   EL:3887 */
  WB_TWE = 8,   /* TweMetric                   EL:3985 */
  WB_WDDTW = 9, /* WeightedDerivativeDtwMetric EL:3404 (Tx <= Ty only: the reference overflows otherwise) */
  WB_WLCSS = 10 /* WeightedLcssMetric          EL:3542 */
};

/* metric_params of the reference constructors (already validated by the caller) */
typedef struct wb_params {
  double r;         /* all: Sakoe-Chiba window fraction in [0,1]                   */
  double g;         /* wdtw/wddtw/wlcss: weight steepness; erp: gap value          */
  double p;         /* adtw: penalty                                               */
  double c;         /* msm: cost                                                   */
  double epsilon;   /* lcss/wlcss/edr: threshold; NaN = edr default max(std)/4     */
  double penalty;   /* twe                                                         */
  double stiffness; /* twe                                                         */
  int32_t engine;   /* 0 auto, 1 force row-scan engine, 2 force strip engine, 3 force band-register engine,
                       4 force the cooperative (lanes-per-pair) engine */
  int32_t precision; /* 0: fp64, bit-equal to the reference (default); 1: fp32 arithmetic (<= 1e-4 relative;
                        lcss / wlcss / edr always run in fp64); 2: fp64 with the DTW-family cost folded into the minimum
                        by ONE fused multiply-add (<= 1e-12 relative, not bit-equal: the reference build has no FMA;
                        every other metric computes exactly as in mode 0)                                  */
} wb_params;

/* timing / work counters of the last call (optional out-parameter) */
typedef struct wb_stats {
  double kernel_ms;     /* device time of the DP kernels (CUDA events), max over devices */
  double total_ms;      /* device time incl. staging copies, max over devices            */
  int64_t cells;        /* DP cells evaluated (reference cell count)          */
  int64_t pairs;        /* pairs evaluated                                               */
  int32_t launches;     /* kernels launched by this call                                 */
  int32_t engine;       /* engine used for the DP (1 row-scan, 2 strip, 3 band-register, 4 cooperative) */
  int64_t lb_kim_pruned;   /* argmin cascade: pairs pruned by LB_Kim                     */
  int64_t lb_keogh_pruned; /* argmin cascade: pairs pruned by LB_Keogh (either direction)  */
  int32_t strip_w;      /* strip engine, last launch: strip width W (columns held in registers); cooperative engine:
                           band coordinates per lane                                                           */
  int32_t strip_nr;     /* rows per fast-path iteration (in-thread wavefront depth); cooperative engine: lanes per pair */
  int32_t strip_warps;  /* warps per CTA                                                       */
  int32_t strip_gring;  /* 1: boundary buffers in global memory (L2), 0: shared memory          */
  int64_t ambiguous;    /* argmin in neighbour-set mode (use_device_lb bit 1): queries whose k-nearest SET depends on the
                           scan's history (a pair outside the set ties with the kth distance) -- redo them without the bit */
} wb_stats;

int wb_cuda_device_count(void);
const char *wb_cuda_last_error(void);

/* out[i*ny + j] = metric(x[i], y[j]).  Replaces _pairwise_distance, CD:1184-1205. */
int wb_cuda_pairwise(int metric, const wb_params *params,
                     const double *x, int64_t nx, int64_t Tx, int64_t x_stride,
                     const double *y, int64_t ny, int64_t Ty, int64_t y_stride,
                     double *out, const int *devices, int n_devices, wb_stats *stats);

/* out (n*n): j > i computed, mirrored to [j][i], zero diagonal.
 * Replaces _singleton_pairwise_distance, CD:1248-1267. */
int wb_cuda_pairwise_self(int metric, const wb_params *params,
                          const double *x, int64_t n, int64_t T, int64_t x_stride,
                          double *out, const int *devices, int n_devices, wb_stats *stats);

/* out[i] = metric(y[i], x[i]) -- the reference evaluates the operands swapped (CD:1632-1647);
 * x and y here are the USER's x and y.  Replaces _paired_distance, CD:1632-1652. */
int wb_cuda_paired(int metric, const wb_params *params,
                   const double *x, int64_t n, int64_t Tx, int64_t x_stride,
                   const double *y, int64_t Ty, int64_t y_stride,
                   double *out, const int *devices, int n_devices, wb_stats *stats);

/* Multivariate forms: x is (nx, n_dims, Tx) with `x_stride` elements between samples and
 * `x_dim_stride` elements between the dimensions of one sample (the reference's TSArray,
 * utils/__init__.pxd:4), y likewise; y == NULL selects the self join (_singleton_pairwise_distance).
 * combine 0 = dim="mean": ONE (nx, ny) matrix, the per-dimension distances summed in dimension order and
 * divided by n_dims on the device (bit-equal to np.mean(list, axis=0), _distance.py:1250-1253 / 1289-1292);
 * combine 1 = dim="full": out is (n_dims, nx, ny).  The dimensions are uploaded once and the result crosses
 * PCIe once.  Replaces the per-dimension Python loops of pairwise_distance (_distance.py:1245-1297) and
 * paired_distance (_distance.py:1163-1169; out is (n,) or (n_dims, n); operands swapped as in wb_cuda_paired). */
int wb_cuda_pairwise_nd(int metric, const wb_params *params,
                        const double *x, int64_t nx, int64_t n_dims, int64_t Tx, int64_t x_stride, int64_t x_dim_stride,
                        const double *y, int64_t ny, int64_t Ty, int64_t y_stride, int64_t y_dim_stride,
                        int combine, double *out, const int *devices, int n_devices, wb_stats *stats);
int wb_cuda_paired_nd(int metric, const wb_params *params,
                      const double *x, int64_t n, int64_t n_dims, int64_t Tx, int64_t x_stride, int64_t x_dim_stride,
                      const double *y, int64_t Ty, int64_t y_stride, int64_t y_dim_stride,
                      int combine, double *out, const int *devices, int n_devices, wb_stats *stats);

/* k nearest y for every x under the reference's sequential early-abandoning scan.
 * out_idx / out_dist: (nx, k) in the reference heap's array order (utils/_misc.pyx:62-107).
 * lower_bound: optional (nx, ny) matrix, pairs with lower_bound >= running threshold are
 * skipped (CD:1331).  use_device_lb bit 0 (dtw only): additionally prune with the on-device
 * LB_Kim / LB_Keogh cascade; for k = 1 the thresholds are also seeded with the exact distance to
 * sketch-nearest candidates (neither ever changes the result).
 * use_device_lb bit 1 (value 2, with bit 0; 1 < k <= 8): neighbour-SET mode for callers that only count
 * the k nearest (KNeighborsClassifier.predict_proba, _neighbors.py:275-287: class votes).  The
 * thresholds are seeded for k > 1 as well (kth smallest candidate distance), so out_idx / out_dist hold
 * the same k neighbours as the reference but in THIS scan's heap order; stats->ambiguous counts the
 * queries for which a pair outside the set ties with the kth distance (which of the tied pairs the
 * reference keeps depends on its scan history) -- the caller repeats the call without the bit if it is
 * not zero.
 * Replaces _argmin_distance, CD:1348-1378. */
int wb_cuda_argmin(int metric, const wb_params *params,
                   const double *x, int64_t nx, int64_t Tx, int64_t x_stride,
                   const double *y, int64_t ny, int64_t Ty, int64_t y_stride,
                   int64_t k, const double *lower_bound, int use_device_lb,
                   int64_t *out_idx, double *out_dist,
                   const int *devices, int n_devices, wb_stats *stats);

/* Device-resident fitted set: the estimators that call this path keep their training set
 * `_fit_X` (distance/_neighbors.py:76, 239) and query it repeatedly (kneighbors :141-150, predict_proba
 * :262-283, KMeans assign :320-347).  wb_cuda_fit uploads y (ny, n_dims, Ty) ONCE to every listed device;
 * the *_fitted calls then move only the queries and the result.  They shard the query rows over the
 * devices of the fitted set and otherwise behave exactly like wb_cuda_pairwise_nd / wb_cuda_argmin. */
typedef struct wb_fitted wb_fitted;
int wb_cuda_fit(const double *y, int64_t ny, int64_t n_dims, int64_t Ty, int64_t y_stride, int64_t y_dim_stride,
                const int *devices, int n_devices, wb_fitted **out);
void wb_cuda_fit_free(wb_fitted *fit);
int wb_cuda_pairwise_fitted(int metric, const wb_params *params,
                            const double *x, int64_t nx, int64_t n_dims, int64_t Tx, int64_t x_stride, int64_t x_dim_stride,
                            const wb_fitted *fit, int combine, double *out, wb_stats *stats);
int wb_cuda_argmin_fitted(int metric, const wb_params *params,
                          const double *x, int64_t nx, int64_t Tx, int64_t x_stride,
                          const wb_fitted *fit, int64_t k, const double *lower_bound, int use_device_lb,
                          int64_t *out_idx, double *out_dist, wb_stats *stats);

/* DTW alignments and DBA.
 *
 * wb_cuda_dtw_paths: optimal warping paths of n_pairs pairs (a[ia[p]], b[ib[p]]) (ia / ib NULL: pair p uses
 * series p), a: (na, Ta) rows of the DP, b: (nb, Tb) columns, host buffers.  Replaces `_dtw_alignment`
 * (_elastic.pyx:1011-1073) + the Python back-walk `dtw_mapping` (distance/dtw.py:385-413): band
 * max(floor(max(Ta, Tb) r), 1) (dtw.py:38-40), optional `weights` (max(Ta, Tb) values, cost v*v*weights[|i-j|]).
 * path_lo / path_hi: (n_pairs, Ta) int32, the path occupies columns lo..hi of each row (== the rows of
 * `indicator.nonzero()`); cost (optional, n_pairs): D[Ta-1][Tb-1]; out_matrix (optional, n_pairs x Ta x Tb):
 * the alignment matrix with +inf outside the band (the reference leaves those cells uninitialised).
 *
 * wb_cuda_dba_epoch: one majorize-minimize step of `_mm_dtw_average` (distance/dtw.py:655-690) for K
 * barycentres at once against a resident sample set.  Cluster c owns members[member_offsets[c] ..
 * member_offsets[c+1]) (sample indices, ascending).  do_update != 0: every member is aligned with means_in[c]
 * and means_out[c] = z / V accumulated in the reference's order (bit-equal); do_update == 0: means_out =
 * means_in.  dist_out[q] = metric(means_out[c], sample members[q]) for metric WB_DTW / WB_WDTW (the cost
 * function of dtw_average, dtw.py:590-605; the caller averages with numpy).  sample_weight: optional, one per
 * SAMPLE of the fitted set; weights: the alignment's weight vector (wdtw: jeong_weight(max(Tm, T), g) computed
 * by the caller with numpy, dtw.py:347-374) or NULL. */
int wb_cuda_dtw_paths(const double *a, int64_t na, int64_t Ta, int64_t a_stride,
                      const double *b, int64_t nb, int64_t Tb, int64_t b_stride,
                      const int64_t *ia, const int64_t *ib, int64_t n_pairs, double r, const double *weights,
                      int32_t *path_lo, int32_t *path_hi, double *cost, double *out_matrix, int device, wb_stats *stats);
int wb_cuda_dba_epoch(const wb_fitted *fit, int metric, const wb_params *params,
                      const double *means_in, int64_t K, int64_t Tm,
                      const int64_t *member_offsets, const int64_t *members, const double *sample_weight,
                      const double *weights, int do_update, double *means_out, double *dist_out, wb_stats *stats);

/* Subsequence search with the elastic metrics: for every subsequence k (s[s_offsets[k] .. s_offsets[k+1]),
 * length m_k <= T) and sample i, the minimum over the sliding windows w = 0 .. T - m_k of
 * metric(subsequence, x[i][w : w + m_k]) and the FIRST window that attains it under the reference's scan.  Replaces
 * `_pairwise_subsequence_distance` / `_paired_subsequence_distance` (_cdistance.pyx:1018-1062) over the elastic
 * SubsequenceMetric classes of _distance.py:143-178.
 * paired == 0: out_dist / out_idx are (nx, n_s); paired != 0 (n_s == nx): subsequence i against sample i, (nx).
 *
 * scaled == 0, DTW family (Dtw / WeightedDtw / AmercingDtw / DerivativeDtw / WeightedDerivativeDtw SubsequenceMetric,
 * _elastic.pyx:2206-2615; dtw_subsequence_distance :622-660, adtw :701-740, ddtw :780-815): band from the subsequence
 * length (_compute_r(m_k, r)), comparison in the squared-cost domain, sqrt of the minimum; wdtw / wddtw weights span the
 * SERIES length (T, T - 2).  The reference abandons windows early against the running minimum, which never changes
 * the result for these metrics; the device evaluates every window (all windows of all samples are the `y` operand of
 * ONE pairwise launch per subsequence, addressed with stride 1) and takes the first minimum.
 *
 * scaled == 0, lcss / erp / edr / msm / twe (Lcss / Erp / Edr / Msm / Twe SubsequenceMetric, _elastic.pyx:2616-3124;
 * *_subsequence_distance :1186, :1350, :1500, :1650, :1832): here the early abandoning against the running minimum DOES
 * decide which windows are accepted (row minima are not monotone; lcss compares a similarity with a distance bound,
 * edr scales the bound with the SERIES length), so the scan is replayed exactly: one DP launch records every window's
 * distance and the maximum of its row minima, one warp per sample then walks the windows in order.  edr's default epsilon
 * (params->epsilon NaN) is std / 4 of each subsequence: pass it in s_epsilon (n_s values, computed by the caller with numpy
 * as ScaledSubsequenceMetric.from_array does, _cdistance.pyx:453-467); s_epsilon is ignored otherwise and may be NULL.
 *
 * scaled != 0, metric WB_DTW: `scaled_dtw`, the UCR-suite search of ScaledDtwSubsequenceMetric (_elastic.pyx:1928-2060,
 * scaled_dtw_subsequence_distance :353-482, inner_scaled_dtw_subsequence_distance :263-345): `s` holds the subsequences
 * already z-normalised by the caller ((s - mean) / std with numpy's mean / std, _cdistance.pyx:453-467), every window is
 * normalised with its running mean / std in the reference's summation order, band |i - j| <= _compute_warp_width(m, r)
 * (_elastic.pyx:1917-1921); the reference's LB_Kim prefilter is part of the observable result and is replayed.
 *
 * scaled != 0, any other metric: `scaled_<metric>` = ScaledSubsequenceMetricWrap(Metric) (_cdistance.pyx:470-551): `s`
 * z-normalised by the caller as above, every window z-normalised with the reference's running IncStats
 * (utils/_stats.pyx:45-93), then Metric._eadistance() against the running minimum -- replayed exactly as above (weights
 * of wdtw / wddtw over T / T - 2 as wrap.reset(X, X) sizes them; edr's default epsilon from the two normalised buffers). */
int wb_cuda_subsequence(int metric, const wb_params *params,
                        const double *s, const int64_t *s_offsets, int64_t n_s,
                        const double *x, int64_t nx, int64_t T, int64_t x_stride,
                        int paired, int scaled, const double *s_epsilon, double *out_dist, int64_t *out_idx,
                        const int *devices, int n_devices, wb_stats *stats);

/* Subsequence matches / distance profile: the dense form of SubsequenceMetric._matches
 * (_cdistance.pyx:311-372; the *_subsequence_matches of _elastic.pyx:658-700, 737-779, 821-868, 1227-1270, 1391-1434,
 * 1539-1580, 1689-1730, 1872-1914, scaled_dtw_matches :485-619, ScaledSubsequenceMetricWrap._matches _cdistance.pyx:553-606).
 * Replaces the per-sample loops of `_subsequence_match` / `_paired_subsequence_match` (_cdistance.pyx:1065-1141) and
 * `_distance_profile` (_cdistance.pyx:1655-1725, = _matches with threshold +inf).
 * s: (n_s, m) dense; n_s == 1: the one subsequence against every sample, n_s == nx: subsequence i against sample i.
 * out: (nx, T - m + 1); out[i][w] = distance of window w of sample i where the reference reports a match under
 * `threshold` (+inf: every window, the distance profile), NaN elsewhere.  Match rules as in the reference: unscaled
 * metrics `dist <= threshold` (DTW family in the squared-cost domain), the scaled wraps `dist < threshold`; a window whose
 * DP the reference abandons against the threshold (lcss: m - threshold m, edr: threshold max(m, T) resp. threshold m) is
 * not a match; scaled_dtw skips windows whose (invalid) LB_Kim is >= threshold^2.  scaled / s_epsilon as in
 * wb_cuda_subsequence (the caller z-normalises s and resolves edr's default epsilon from the statistics the reference
 * uses on that path: numpy's for subsequence_match, sequential sums for distance_profile, _cdistance.pyx:167-185). */
int wb_cuda_subsequence_profile(int metric, const wb_params *params,
                                const double *s, int64_t n_s, int64_t m,
                                const double *x, int64_t nx, int64_t T, int64_t x_stride,
                                int scaled, const double *s_epsilon, double threshold, double *out,
                                const int *devices, int n_devices, wb_stats *stats);

/* The k closest windows of sample i to subsequence i.  Replaces `_argmin_subsequence_distance`
 * (_cdistance.pyx:1380-1600: `_ArgminSubsequenceDistance` for scaled == 0, `_ScaledArgminSubsequenceDistance` otherwise), the
 * driver behind argmin_subsequence_distance (_distance.py:1636-1790): per sample a sequential scan of the windows with
 * Metric._eadistance against the running bound (+inf until the k-heap is full, then its maximum, utils/_misc.pyx:62-107).
 * s: (nx, m) dense; scaled != 0: s already z-normalised by the caller with fast_mean_std (utils/_stats.pyx:22-42, std 0 -> 1)
 * and every window z-normalised on the device with the running IncStats; wdtw / wddtw weights over T / T - 2.
 * out_idx / out_dist: (nx, k) in the heap's array order.  The dilated / padded distance profile (`_dilated_distance_profile`,
 * _cdistance.pyx:804-935) uses this entry point with k = 1, the dilated windows as `s` and the kernel parts as one-window samples;
 * weight_len > 0 sizes the wdtw / wddtw weight tables for a series of that many points (the series the reference reset() the
 * metric with, _cdistance.pyx:1769) instead of T; 0 = T.  Same device scheme as the profile (all windows of all samples in
 * one DP launch) followed by the exact replay of the scan, one warp per sample. */
int wb_cuda_subsequence_argmin(int metric, const wb_params *params,
                               const double *s, int64_t n_s, int64_t m,
                               const double *x, int64_t nx, int64_t T, int64_t x_stride,
                               int scaled, int64_t k, int64_t weight_len, int64_t *out_idx, double *out_dist,
                               const int *devices, int n_devices, wb_stats *stats);

/* Device-resident variant of wb_cuda_pairwise: d_x (nx, Tx), d_y (ny, Ty), d_out (nx, ny) are
 * dense row-major DEVICE arrays on the current device.  Enqueues on `stream`; fills `stats`
 * (after synchronising the stream) when stats != NULL. */
int wb_cuda_pairwise_dev(int metric, const wb_params *params,
                         const double *d_x, int64_t nx, int64_t Tx,
                         const double *d_y, int64_t ny, int64_t Ty,
                         double *d_out, void *stream, wb_stats *stats);

/* Lower-bound matrices of wildboar.distance.lb, host buffers, one device.
 * out[i * nx + j] for query i (rows of q) and fitted sample j (rows of x); both (n, T) float64 with
 * `*_stride` elements between consecutive samples.
 *
 * wb_cuda_lb_keogh replaces DtwKeoghLowerBound(r, kind).fit(x).transform(q), distance/lb.py:359-432
 * (a Python double loop over _dtw_lb_keogh, _elastic.pyx:1095-1115): envelope half-width
 * max(floor(T r), 1) (T - 1 when that equals T, lb.py:367-369), kind 0 = "both" (maximum of the two
 * directions), 1 = "left" (query against the sample's envelope), 2 = "right".
 * wb_cuda_lb_kim replaces DtwKimLowerBound().fit(x).transform(q), distance/lb.py:224-311 (sum of the
 * squared first/last-three-point terms; the reference applies no square root). */
int wb_cuda_lb_keogh(const double *q, int64_t nq, int64_t q_stride,
                     const double *x, int64_t nx, int64_t x_stride, int64_t T,
                     double r, int kind, double *out, int device, wb_stats *stats);
int wb_cuda_lb_kim(const double *q, int64_t nq, int64_t q_stride,
                   const double *x, int64_t nx, int64_t x_stride, int64_t T,
                   double *out, int device, wb_stats *stats);

/* dtw_envelop and dtw_lb_keogh of wildboar.distance.dtw, batched over n series (all dense, C order).
 * wb_cuda_dtw_envelope: lower/upper[i][k] = min/max of x[i][max(0,k-w) .. min(T-1,k+w)]; `w` is the resolved warp size
 *   (distance/dtw.py:155-190: max(floor(T r), 1), T - 1 when that equals T); fails for w outside [0, T) like
 *   _dtw_envelop (EL:1076-1092).
 * wb_cuda_dtw_lb_keogh_terms: cb[i][k] = squared excess of x[i][k] over [lower[i][k], upper[i][k]], min_dist[i] =
 *   sqrt(sum_k cb[i][k]) summed in time order -- _dtw_lb_keogh, EL:1095-1115 (cumulative_bound EL:228-260). */
int wb_cuda_dtw_envelope(const double *x, int64_t n, int64_t T, int64_t x_stride, int64_t w,
                         double *lower, double *upper, int device, wb_stats *stats);
int wb_cuda_dtw_lb_keogh_terms(const double *x, const double *lower, const double *upper, int64_t n, int64_t T,
                               double *min_dist, double *cb, int device, wb_stats *stats);

/* Page-locked host memory for RESULT buffers (optional).  The host entry points accept any caller-owned `out`; when it
 * is page-locked (from here, cudaHostAlloc or cudaHostRegister) the result slabs arrive by asynchronous DMA at PCIe
 * speed instead of through the driver's pageable staging.  Released blocks are pooled (WILDBOAR_CUDA_PINNED_POOL_MB,
 * default 4096) because page-locking costs ~0.3 ms per MB.  This is the allocation the reference does with
 * `np.empty((x.shape[0], y.shape[0]))` in _pairwise_distance / _singleton_pairwise_distance (CD:1197, 1260).
 * wb_cuda_host_alloc returns NULL when no page-locked memory can be had (the caller then uses ordinary memory). */
void *wb_cuda_host_alloc(size_t bytes);
void wb_cuda_host_free(void *p);

/* Measured FP64 issue rate of the current device: runs a register-only DADD/DMUL chain on
 * every SM and returns FP64 warp-lane instructions per second (the ALU roofline denominator,
 * SURVEY 8d).  mix: 0 = DADD only, 1 = DTW cell mix (3 arithmetic + 2 compare/select). */
int wb_cuda_fp64_peak(int mix, double *inst_per_s, double *sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* WB_CUDA_H */
