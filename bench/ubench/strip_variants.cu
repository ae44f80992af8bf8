// DEV TOOLING (not part of the product library): times strip-kernel variants (strip width W,
// in-thread wavefront depth NR, warps per CTA) on a cfg3-shaped problem and runs FP64 issue
// microbenchmarks.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
//   -std=c++17 -lineinfo -o strip_variants strip_variants.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../wildboar_b200/csrc/dispatch.cuh"
#include "../../wildboar_b200/csrc/kernels.cuh"
#include "../../wildboar_b200/csrc/prep.hpp"

using namespace wb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ---------------- issue-rate microbenchmarks (inline PTX so ptxas keeps the mix) -------------
template <int MIX>
__global__ void __launch_bounds__(256) k_issue(int iters, double seed, double* out, unsigned long long* cyc) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  int i0 = (int)seed, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
  float f0 = (float)seed, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  const double inc = seed * 1e-9 + 1e-7;
  const int k = threadIdx.x | 1;
  unsigned long long c0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#define D(a) asm volatile("add.f64 %0, %0, %1;" : "+d"(a) : "d"(inc));
#define I(a) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a) : "r"(k));
#define L(a) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(k));
#define F(a) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(1.0001f));
#define S(a, b) asm volatile("{ .reg .pred q; setp.lt.s32 q, %1, %2; selp.b32 %0, %0, %1, q; }" : "+r"(a) : "r"(b), "r"(k));
    if (MIX == 0) { D(a0) D(a1) D(a2) D(a3) D(a4) D(a5) D(a6) D(a7) }
    if (MIX == 1) { D(a0) L(i0) D(a1) L(i1) D(a2) L(i2) D(a3) L(i3) D(a4) L(i4) D(a5) L(i5) D(a6) L(i6) D(a7) L(i7) }
    if (MIX == 2) { D(a0) I(i0) D(a1) I(i1) D(a2) I(i2) D(a3) I(i3) D(a4) I(i4) D(a5) I(i5) D(a6) I(i6) D(a7) I(i7) }
    if (MIX == 3) { D(a0) L(i0) F(f0) D(a1) L(i1) F(f1) D(a2) L(i2) F(f2) D(a3) L(i3) F(f3) D(a4) L(i4) F(f4) D(a5) L(i5) F(f5) D(a6) L(i6) F(f6) D(a7) L(i7) F(f7) }
    if (MIX == 4) { L(i0) L(i1) L(i2) L(i3) L(i4) L(i5) L(i6) L(i7) }
    if (MIX == 5) { L(i0) F(f0) L(i1) F(f1) L(i2) F(f2) L(i3) F(f3) L(i4) F(f4) L(i5) F(f5) L(i6) F(f6) L(i7) F(f7) }
    if (MIX == 6) { D(a0) L(i0) L(i1) D(a1) L(i2) L(i3) D(a2) L(i4) L(i5) D(a3) L(i6) L(i7) D(a4) L(i0) L(i1) D(a5) L(i2) L(i3) D(a6) L(i4) L(i5) D(a7) L(i6) L(i7) }
  }
  unsigned long long c1 = clock64();
  long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  out[gtid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (double)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7) + (double)(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int MIX>
static void run_issue(const char* name, int ninst, int sms, int warps_per_smsp) {
  const int threads = 256, iters = 1 << 14;
  const int ctas_per_sm = warps_per_smsp * 4 * 32 / threads;
  const int grid = sms * ctas_per_sm;
  double* out; unsigned long long* cyc;
  CK(cudaMalloc(&out, sizeof(double) * grid * threads)); CK(cudaMalloc(&cyc, sizeof(unsigned long long) * grid));
  k_issue<MIX><<<grid, threads>>>(iters, 1.5, out, cyc); CK(cudaDeviceSynchronize());
  k_issue<MIX><<<grid, threads>>>(iters, 2.5, out, cyc); CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> h(grid);
  CK(cudaMemcpy(h.data(), cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : h) avg += (double)v; avg /= grid;
  // warp-instructions issued per SMSP per cycle
  double per_smsp = (double)warps_per_smsp * iters * ninst / avg;
  printf("issue %-28s warps/SMSP=%d  inst/cycle/SMSP=%.3f  cycles/iter/warp-set=%.2f\n", name, warps_per_smsp, per_smsp, avg / iters);
  cudaFree(out); cudaFree(cyc);
}

// ---------------- strip kernel variants ----------------
struct Result { double ms; double gcups; double checksum; };

template <class M, int W, int NR, int NWARPS, int MINB, bool GRING = false>
static Result run_variant(const char* pname, int ops, const M& m, const void* dxv, const void* dyv, long long nx, long long ny, int T, int R,
                          double* dout, unsigned long long* counter, int sms, int reps) {
  using F = typename M::real;
  const F* dx = (const F*)dxv; const F* dy = (const F*)dyv;
  KArgsT<F> a; memset(&a, 0, sizeof a);
  a.x = dx; a.y = dy; a.nx = nx; a.ny = ny; a.Tx = T; a.Ty = T; a.g = make_geom(T, T, R);
  a.NS = strip_ring_slots(a.g, W); a.out = dout; a.ld = ny; a.counter = counter; a.mode = PM_PAIRWISE;
  a.nyb = (ny + 31) / 32; a.ntasks = nx * a.nyb; a.ys = T;
  size_t smem = GRING ? 0 : (size_t)NWARPS * a.NS * 32 * sizeof(F);
  auto kern = k_strip<M, W, NWARPS * 32, MINB, false, NR, GRING>;
  Result r{0, 0, 0};
  if (smem > 232448) { printf("strip W=%2d NR=%d warps/CTA=%d: does not fit in shared memory\n", W, NR, NWARPS); r.ms = -1; return r; }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NWARPS * 32, smem));
  if (per_sm < 1) { r.ms = -2; return r; }
  if (GRING) { if (per_sm > MINB) per_sm = MINB; CK(cudaMalloc(&a.gring, (size_t)sms * per_sm * NWARPS * a.NS * 32 * sizeof(F))); }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps + 1; ++rep) {
    CK(cudaMemset(counter, 0, sizeof(unsigned long long)));
    CK(cudaEventRecord(e0));
    kern<<<sms * per_sm, NWARPS * 32, smem>>>(a, m);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  int64_t C = (int64_t)T * (2 * R - 1) - (int64_t)R * (R - 1);
  r.ms = best; r.gcups = (double)nx * ny * C / (best * 1e-3) / 1e9;
  std::vector<double> h(std::min<long long>(ny, 4096));
  CK(cudaMemcpy(h.data(), dout + (nx - 1) * ny, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  for (double v : h) r.checksum += v;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  if (GRING) cudaFree(a.gring);
  printf("%-5s %s W=%2d NR=%d warps/CTA=%d CTAs/SM=%d regs=%3d smem=%6zu  %8.2f ms  %8.1f GCUPS  (%.1f%% of FP64 peak, %d ops/cell)  chk=%.6f\n",
         pname, GRING ? "gring" : "strip", W, NR, NWARPS, per_sm, fa.numRegs, smem, r.ms, r.gcups, 100.0 * r.gcups * ops / 18448.0, ops, r.checksum);
  fflush(stdout);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return r;
}

int main(int argc, char** argv) {
  int nx = argc > 1 ? atoi(argv[1]) : 1024, ny = argc > 2 ? atoi(argv[2]) : 10000, T = argc > 3 ? atoi(argv[3]) : 512;
  double rr = argc > 4 ? atof(argv[4]) : 0.1;
  int reps = argc > 5 ? atoi(argv[5]) : 2;
  int dev = 0, sms = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs=%d  problem %d x %d, T=%d, r=%.3f\n", sms, nx, ny, T, rr);
  if (argc <= 7) for (int w : {8}) {
    run_issue<0>("DADD", 8, sms, w);
    run_issue<1>("DADD+LOP(alu) 1:1", 16, sms, w);
    run_issue<2>("DADD+IMAD(fma) 1:1", 16, sms, w);
    run_issue<3>("DADD+LOP+FFMA 1:1:1", 24, sms, w);
    run_issue<6>("DADD+2LOP 1:2", 24, sms, w);
  }
  if (argc <= 7) { run_issue<4>("LOP only", 8, sms, 8); run_issue<5>("LOP+FFMA 1:1", 16, sms, 8); }

  std::vector<double> hx((size_t)nx * T), hy((size_t)ny * T);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0 - 0.5; };
  for (int i = 0; i < nx; ++i) { double acc = 0; for (int t = 0; t < T; ++t) { acc += rnd(); hx[(size_t)i * T + t] = acc; } }
  for (int i = 0; i < ny; ++i) { double acc = 0; for (int t = 0; t < T; ++t) { acc += rnd(); hy[(size_t)i * T + t] = acc; } }
  double *dx, *dy, *dout; unsigned long long* counter;
  CK(cudaMalloc(&dx, hx.size() * 8)); CK(cudaMalloc(&dy, hy.size() * 8)); CK(cudaMalloc(&dout, (size_t)nx * ny * 8));
  CK(cudaMalloc(&counter, 8));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, hy.data(), hy.size() * 8, cudaMemcpyHostToDevice));
  int R = (int)compute_r(T, rr);
  DtwPolicy<false, false> m; m.w = nullptr; m.p = 0;
  // tables for the weighted metrics / twe (signed, pointer to the centre)
  std::vector<double> hw = make_weights(0.05, T), htw = make_tw(0.001, T + 1);
  double *dw, *dtw;
  CK(cudaMalloc(&dw, hw.size() * 8)); CK(cudaMalloc(&dtw, htw.size() * 8));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dtw, htw.data(), htw.size() * 8, cudaMemcpyHostToDevice));
  DtwPolicy<true, false> mw; mw.w = dw + table_center(T); mw.p = 0;
  DtwPolicy<false, true> ma; ma.w = nullptr; ma.p = 1.0;
  LcssPolicy<false> ml; ml.w = nullptr; ml.eps = 1.0;
  ErpPolicy me; me.g = 0; me.gx_sum = 0; me.gy_sum = 0;
  EdrPolicy md; md.eps_param = 0.25; md.eps = 0.25;
  MsmPolicy mm; mm.cf = 1.0f; mm.c = 1.0;
  TwePolicy mt; mt.pen = 1.001; mt.tw = dtw + table_center(T + 1);
  // fp32 mode: float copies of the operands and tables
  std::vector<float> hxf(hx.begin(), hx.end()), hyf(hy.begin(), hy.end()), hwf(hw.begin(), hw.end()), htwf(htw.begin(), htw.end());
  float *dxf, *dyf, *dwf, *dtwf;
  CK(cudaMalloc(&dxf, hxf.size() * 4)); CK(cudaMalloc(&dyf, hyf.size() * 4)); CK(cudaMalloc(&dwf, hwf.size() * 4)); CK(cudaMalloc(&dtwf, htwf.size() * 4));
  CK(cudaMemcpy(dxf, hxf.data(), hxf.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dyf, hyf.data(), hyf.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwf, hwf.data(), hwf.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dtwf, htwf.data(), htwf.size() * 4, cudaMemcpyHostToDevice));
  DtwPolicy<false, false, float> fm; fm.w = nullptr; fm.p = 0;
  DtwPolicy<false, true, float> fa; fa.w = nullptr; fa.p = 1.0f;
  MsmPolicyT<float> fmm; fmm.cf = 1.0f; fmm.c = 1.0f;
  TwePolicyT<float> ft; ft.pen = 1.001f; ft.tw = dtwf + table_center(T + 1);
  int set = argc > 6 ? atoi(argv[6]) : 0;
#define RVF(P, NAME, OPS, OBJ, W, NR, NW, MB, GR) run_variant<P, W, NR, NW, MB, GR>(NAME, OPS, OBJ, dxf, dyf, nx, ny, T, R, dout, counter, sms, reps);
#define FG(W, NR, NW, MB) RVF(decltype(fm), "dtw32", 3, fm, W, NR, NW, MB, true)
#define FV(W, NR, NW, MB) RVF(decltype(fm), "dtw32", 3, fm, W, NR, NW, MB, false)
#define FO(W, NR, NW, MB) RVF(decltype(fa), "adtw32", 5, fa, W, NR, NW, MB, true) RVF(decltype(fmm), "msm32", 8, fmm, W, NR, NW, MB, true) RVF(decltype(ft), "twe32", 10, ft, W, NR, NW, MB, true)
#define RV(P, NAME, OPS, OBJ, W, NR, NW, MB, GR) run_variant<P, W, NR, NW, MB, GR>(NAME, OPS, OBJ, dx, dy, nx, ny, T, R, dout, counter, sms, reps);
#define V(W, NR, NW, MB) RV(decltype(m), "dtw", 5, m, W, NR, NW, MB, false)
#define G(W, NR, NW, MB) RV(decltype(m), "dtw", 5, m, W, NR, NW, MB, true)
#define GP(W, NR, NW, MB) RV(decltype(mw), "wdtw", 6, mw, W, NR, NW, MB, true) RV(decltype(ma), "adtw", 7, ma, W, NR, NW, MB, true) \
  RV(decltype(ml), "lcss", 4, ml, W, NR, NW, MB, true) RV(decltype(me), "erp", 6, me, W, NR, NW, MB, true) RV(decltype(md), "edr", 7, md, W, NR, NW, MB, true) \
  RV(decltype(mm), "msm", 8, mm, W, NR, NW, MB, true) RV(decltype(mt), "twe", 10, mt, W, NR, NW, MB, true)
#define GL(W, NR, NW, MB) RV(decltype(mm), "msm", 8, mm, W, NR, NW, MB, true) RV(decltype(mt), "twe", 10, mt, W, NR, NW, MB, true)
  if (set == 0) {  // dtw, tall bands (cfg3 / cfg2 shapes)
    G(12, 4, 12, 1) G(12, 6, 12, 1) G(12, 8, 12, 1) G(8, 2, 16, 1) G(8, 4, 16, 1) G(8, 6, 16, 1) G(8, 8, 16, 1)
    G(10, 6, 14, 1) G(10, 5, 14, 1) G(14, 6, 10, 1) G(16, 6, 8, 1) G(6, 4, 20, 1) G(6, 6, 20, 1)
  } else if (set == 1) {  // dtw, narrow bands (cfg1 / cfg4 shapes)
    V(8, 2, 8, 2) V(8, 4, 8, 2) V(8, 6, 8, 2) V(8, 2, 16, 1) V(6, 2, 8, 2) V(6, 2, 10, 2) V(4, 2, 12, 2)
  } else if (set == 2) {  // the other metrics, cfg2 shape
    GP(8, 4, 12, 1) GP(8, 4, 16, 1) GP(8, 2, 16, 1) GP(12, 4, 12, 1) GP(12, 6, 12, 1) GP(8, 6, 12, 1)
  } else if (set == 7) {  // more (W, NR, warps) points for the non-dtw metrics
    GP(10, 6, 14, 1) GP(10, 4, 14, 1) GP(8, 6, 16, 1) GP(10, 6, 12, 1) GP(12, 4, 14, 1) GP(6, 6, 20, 1)
  } else if (set == 4) {  // fp32 mode, dtw (tall bands: global buffers; also shared memory, which now fits more warps)
    FG(12, 6, 12, 1) FG(12, 4, 16, 1) FG(16, 4, 16, 1) FG(16, 6, 16, 1) FG(16, 8, 16, 1) FG(24, 4, 12, 1) FG(24, 6, 12, 1) FG(16, 4, 24, 1) FG(16, 4, 12, 2) FG(32, 4, 8, 1) FG(20, 6, 16, 1)
    FV(16, 4, 12, 1) FV(16, 6, 12, 1) FV(24, 4, 12, 1) FV(12, 6, 12, 1)
  } else if (set == 5) {  // fp32 mode, narrow bands (shared memory)
    FV(8, 2, 8, 2) FV(8, 4, 8, 2) FV(16, 4, 8, 2) FV(16, 2, 8, 2) FV(16, 4, 16, 1) FV(12, 4, 16, 1) FV(16, 4, 12, 2)
  } else if (set == 6) {  // fp32 mode, other metrics
    FO(12, 6, 12, 1) FO(16, 4, 16, 1) FO(16, 6, 16, 1) FO(24, 4, 12, 1) FO(8, 4, 16, 1)
  } else {  // long series (cfg5 shape): msm / twe
    GL(16, 4, 8, 1) GL(12, 4, 8, 1) GL(12, 4, 12, 1) GL(8, 4, 12, 1) GL(8, 4, 16, 1) GL(8, 2, 16, 1) GL(12, 6, 12, 1) GL(8, 4, 8, 2)
    GL(10, 4, 14, 1) GL(10, 6, 14, 1) GL(8, 6, 16, 1) GL(12, 4, 14, 1)
  }
  printf("done\n");
  return 0;
}
