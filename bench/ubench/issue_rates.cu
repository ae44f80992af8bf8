// DEV TOOLING: how do FP64 instructions share the issue port with ALU/FMA-pipe instructions on
// B200?  Whole-chip wall-clock rates (CUDA events), register-only inline-PTX loops.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

#define D(a) asm volatile("add.f64 %0, %0, %1;" : "+d"(a) : "d"(inc));
#define M(a) asm volatile("mul.f64 %0, %0, %1;" : "+d"(a) : "d"(one));
#define S(a) asm volatile("selp.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(k), "r"(k));   /* never folds: predicate from k */
#define SP(a, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.b32 %0, %0, %1, q; }" : "+r"(a) : "r"(k), "r"(p));
#define X(a) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(k));
#define I(a) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a) : "r"(k));
#define F(a) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(1.0001f));
// DTW cell on one chain: v = x - y; m = min(min(up, left), diag); d = m + v*v   (selects via setp.f64 + selp.f64)
#define CELL(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; selp.f64 m, %1, %0, q; setp.lt.f64 q, %2, m; selp.f64 m, %2, m, q; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));

// variants of the min implementation
#define CELL_SEL(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b32 al, ah, bl, bh; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; mov.b64 {al, ah}, %1; mov.b64 {bl, bh}, %0; selp.b32 al, al, bl, q; selp.b32 ah, ah, bh, q; mov.b64 m, {al, ah}; setp.lt.f64 q, %2, m; mov.b64 {bl, bh}, %2; selp.b32 al, bl, al, q; selp.b32 ah, bh, ah, q; mov.b64 m, {al, ah}; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));
#define CELL_NOSEL(d, up, dg, y) asm volatile("{ .reg .pred q, q2; .reg .f64 v; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; setp.lt.f64 q2, %2, %0; mul.f64 v, v, v; add.f64 %0, %0, v; @q add.s32 %5, %5, 1; @q2 add.s32 %5, %5, 2; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y), "r"(i0));
#define CELL_NOSETP(d, up, dg, y, p) asm volatile("{ .reg .pred q; .reg .f64 v, m; sub.f64 v, %3, %4; setp.ne.s32 q, %5, 0; selp.f64 m, %1, %0, q; selp.f64 m, %2, m, q; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y), "r"(p));
#define CELL_ARITH(d, up, dg, y) asm volatile("{ .reg .f64 v; sub.f64 v, %3, %4; mul.f64 v, v, v; add.f64 %0, %0, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));
#define CELL_INT1(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b64 a, b, c; sub.f64 v, %3, %4; mov.b64 a, %1; mov.b64 b, %0; min.u64 c, a, b; mov.b64 m, c; setp.lt.f64 q, %2, m; selp.f64 m, %2, m, q; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));


// ---- select-pipe experiments: where do the four 32-bit selects of the two mins execute? ----
// HYB: per min, low half via selp.b32 (ALU pipe), high half via a predicated IMAD (FMA-heavy pipe; `one` is an opaque runtime 1)
#define CELL_HYB(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b32 al, ah, bl, bh; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; mov.b64 {al, ah}, %1; mov.b64 {bl, bh}, %0; selp.b32 bl, al, bl, q; @q mad.lo.u32 bh, ah, %5, 0; mov.b64 m, {bl, bh}; setp.lt.f64 q, %2, m; mov.b64 {al, ah}, %2; selp.b32 bl, al, bl, q; @q mad.lo.u32 bh, ah, %5, 0; mov.b64 m, {bl, bh}; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y), "r"(onei));
// IMAD: all four selects as predicated IMADs
#define CELL_IMAD(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b32 al, ah, bl, bh; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; mov.b64 {al, ah}, %1; mov.b64 {bl, bh}, %0; @q mad.lo.u32 bl, al, %5, 0; @q mad.lo.u32 bh, ah, %5, 0; mov.b64 m, {bl, bh}; setp.lt.f64 q, %2, m; mov.b64 {al, ah}, %2; @q mad.lo.u32 bl, al, %5, 0; @q mad.lo.u32 bh, ah, %5, 0; mov.b64 m, {bl, bh}; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y), "r"(onei));
// PMOV: predicated mov.b32 (ptxas picks the pipe)
#define CELL_PMOV(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b32 al, ah, bl, bh; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; mov.b64 {al, ah}, %1; mov.b64 {bl, bh}, %0; @q mov.b32 bl, al; @q mov.b32 bh, ah; mov.b64 m, {bl, bh}; setp.lt.f64 q, %2, m; mov.b64 {al, ah}, %2; @q mov.b32 bl, al; @q mov.b32 bh, ah; mov.b64 m, {bl, bh}; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));
// PADD: second min folded into two predicated DADDs (no selects for it): q ? dg + c : m + c
#define CELL_PADD(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; selp.f64 m, %1, %0, q; setp.lt.f64 q, %2, m; mul.f64 v, v, v; @q add.f64 %0, %2, v; @!q add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));
// HYBPADD: first min hybrid (ALU + IMAD), second min as predicated DADDs
#define CELL_HYBPADD(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b32 al, ah, bl, bh; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; mov.b64 {al, ah}, %1; mov.b64 {bl, bh}, %0; selp.b32 bl, al, bl, q; @q mad.lo.u32 bh, ah, %5, 0; mov.b64 m, {bl, bh}; setp.lt.f64 q, %2, m; mul.f64 v, v, v; @q add.f64 %0, %2, v; @!q add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y), "r"(onei));
// SHF: high half via funnel shift by 0 under predicate? (ALU) -- control: selp both halves but independent int regs not from FP64 results
#define CELL_FSELONLY(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; selp.f64 m, %1, %0, q; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));
// MIXSEL: low halves via selp.b32 (SEL), high halves via selp.f32 (FSEL): are they different pipes?
#define CELL_MIXSEL(d, up, dg, y) asm volatile("{ .reg .pred q; .reg .f64 v, m; .reg .b32 al, bl; .reg .f32 ah, bh; sub.f64 v, %3, %4; setp.lt.f64 q, %1, %0; mov.b64 {al, ah}, %1; mov.b64 {bl, bh}, %0; selp.b32 bl, al, bl, q; selp.f32 bh, ah, bh, q; mov.b64 m, {bl, bh}; setp.lt.f64 q, %2, m; mov.b64 {al, ah}, %2; selp.b32 bl, al, bl, q; selp.f32 bh, ah, bh, q; mov.b64 m, {bl, bh}; mul.f64 v, v, v; add.f64 %0, m, v; }" : "+d"(d) : "d"(up), "d"(dg), "d"(xx), "d"(y));

template <int MIX>
__global__ void __launch_bounds__(256) k(int iters, double seed, double* out) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
  int j0 = i0 * 3, j1 = i1 * 3, j2 = i2 * 3, j3 = i3 * 3, j4 = i4 * 3, j5 = i5 * 3, j6 = i6 * 3, j7 = i7 * 3;
  float f0 = 1, f1 = 2, f2 = 3, f3 = 4, f4 = 5, f5 = 6, f6 = 7, f7 = 8;
  const double inc = seed * 1e-9 + 1e-7, one = 1.0 + seed * 1e-12, xx = seed * 0.37;
  const int k = (int)blockIdx.x | 1;
  const int onei = (iters > 0) ? 1 : (int)seed;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MIX == 0) { D(a0) D(a1) D(a2) D(a3) D(a4) D(a5) D(a6) D(a7) }
    if (MIX == 1) { D(a0) SP(i0, j0) D(a1) SP(i1, j1) D(a2) SP(i2, j2) D(a3) SP(i3, j3) D(a4) SP(i4, j4) D(a5) SP(i5, j5) D(a6) SP(i6, j6) D(a7) SP(i7, j7) }
    if (MIX == 2) { D(a0) X(i0) X(j0) D(a1) X(i1) X(j1) D(a2) X(i2) X(j2) D(a3) X(i3) X(j3) D(a4) X(i4) X(j4) D(a5) X(i5) X(j5) D(a6) X(i6) X(j6) D(a7) X(i7) X(j7) }
    if (MIX == 3) { X(i0) X(j0) X(i1) X(j1) X(i2) X(j2) X(i3) X(j3) X(i4) X(j4) X(i5) X(j5) X(i6) X(j6) X(i7) X(j7) }
    if (MIX == 4) { D(a0) I(i0) D(a1) I(i1) D(a2) I(i2) D(a3) I(i3) D(a4) I(i4) D(a5) I(i5) D(a6) I(i6) D(a7) I(i7) }
    if (MIX == 5) { D(a0) X(i0) I(j0) D(a1) X(i1) I(j1) D(a2) X(i2) I(j2) D(a3) X(i3) I(j3) D(a4) X(i4) I(j4) D(a5) X(i5) I(j5) D(a6) X(i6) I(j6) D(a7) X(i7) I(j7) }
    if (MIX == 6) { CELL(a0, a4, a1, a5) CELL(a1, a5, a2, a6) CELL(a2, a6, a3, a7) CELL(a3, a7, a0, a4) CELL(a4, a0, a5, a1) CELL(a5, a1, a6, a2) CELL(a6, a2, a7, a3) CELL(a7, a3, a4, a0) }
    if (MIX == 7) { D(a0) X(i0) D(a1) X(i1) D(a2) X(i2) D(a3) X(i3) D(a4) X(i4) D(a5) X(i5) D(a6) X(i6) D(a7) X(i7) }
    if (MIX == 8) { X(i0) I(j0) X(i1) I(j1) X(i2) I(j2) X(i3) I(j3) X(i4) I(j4) X(i5) I(j5) X(i6) I(j6) X(i7) I(j7) }
    if (MIX == 9) { D(a0) M(a1) D(a2) M(a3) D(a4) M(a5) D(a6) M(a7) }
#define ALL8(C) C(a0, a4, a1, a5) C(a1, a5, a2, a6) C(a2, a6, a3, a7) C(a3, a7, a0, a4) C(a4, a0, a5, a1) C(a5, a1, a6, a2) C(a6, a2, a7, a3) C(a7, a3, a4, a0)
    if (MIX == 10) { ALL8(CELL_SEL) }
    if (MIX == 11) { ALL8(CELL_NOSEL) }
    if (MIX == 12) { CELL_NOSETP(a0, a4, a1, a5, j0) CELL_NOSETP(a1, a5, a2, a6, j1) CELL_NOSETP(a2, a6, a3, a7, j2) CELL_NOSETP(a3, a7, a0, a4, j3) CELL_NOSETP(a4, a0, a5, a1, j4) CELL_NOSETP(a5, a1, a6, a2, j5) CELL_NOSETP(a6, a2, a7, a3, j6) CELL_NOSETP(a7, a3, a4, a0, j7) }
    if (MIX == 13) { ALL8(CELL_ARITH) }
    if (MIX == 14) { ALL8(CELL_INT1) }
    if (MIX == 15) { ALL8(CELL_HYB) }
    if (MIX == 16) { ALL8(CELL_IMAD) }
    if (MIX == 17) { ALL8(CELL_PMOV) }
    if (MIX == 18) { ALL8(CELL_PADD) }
    if (MIX == 19) { ALL8(CELL_HYBPADD) }
    if (MIX == 20) { ALL8(CELL_FSELONLY) }
    if (MIX == 21) { ALL8(CELL_MIXSEL) }
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (double)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7 + j0 + j1 + j2 + j3 + j4 + j5 + j6 + j7) + (double)(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
}

template <int MIX>
void run(const char* name, int n_inst, int n_fp64, int sms, int warps_per_smsp, double clk_ghz) {
  const int iters = 1 << 14, threads = 256;
  const int grid = sms * (warps_per_smsp * 4 * 32 / threads);
  double* out; CK(cudaMalloc(&out, sizeof(double) * grid * threads));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); k<MIX><<<grid, threads>>>(iters, 1.5 + rep, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
  }
  const double warp_inst = (double)grid * (threads / 32) * iters * n_inst;
  const double cyc = best * 1e-3 * clk_ghz * 1e9;
  printf("%-34s w/SMSP=%2d  %7.3f ms  inst/clk/SMSP=%.3f  fp64 inst/clk/SMSP=%.3f  clk/iter/warp-slot=%.2f\n", name, warps_per_smsp, best,
         warp_inst / (cyc * sms * 4), warp_inst * n_fp64 / n_inst / (cyc * sms * 4), cyc * sms * 4 / ((double)grid * (threads / 32) * iters) );
  cudaFree(out);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const double clk = 1.965;
  for (int w : {8}) {
    run<0>("8 DADD", 8, 8, sms, w, clk);
    run<9>("4 DADD + 4 DMUL", 8, 8, sms, w, clk);
    run<7>("8 DADD + 8 LOP", 16, 8, sms, w, clk);
    run<1>("8 DADD + 8 (ISETP+SEL)", 24, 8, sms, w, clk);
    run<2>("8 DADD + 16 LOP", 24, 8, sms, w, clk);
    run<4>("8 DADD + 8 IMAD", 16, 8, sms, w, clk);
    run<5>("8 DADD + 8 LOP + 8 IMAD", 24, 8, sms, w, clk);
    run<3>("16 LOP", 16, 0, sms, w, clk);
    run<8>("8 LOP + 8 IMAD", 16, 0, sms, w, clk);
    run<6>("8 DTW cells (5 fp64 + 2 sel.f64)", 72, 40, sms, w, clk);
    run<10>("8 cells, selp.b32 halves", 72, 40, sms, w, clk);
    run<11>("8 cells, 2 DSETP no selects", 56, 40, sms, w, clk);
    run<12>("8 cells, 4 FSEL no DSETP", 64, 24, sms, w, clk);
    run<13>("8 cells, arithmetic only", 24, 24, sms, w, clk);
    run<14>("8 cells, 1 min.u64 + 1 DSETP min", 72, 32, sms, w, clk);
    run<15>("8 cells, selects: SEL lo + IMAD hi", 72, 40, sms, w, clk);
    run<16>("8 cells, selects: 4 pred IMAD", 72, 40, sms, w, clk);
    run<17>("8 cells, selects: 4 pred MOV", 72, 40, sms, w, clk);
    run<18>("8 cells, min2 as 2 pred DADD", 64, 48, sms, w, clk);
    run<19>("8 cells, min1 SEL+IMAD, min2 pred DADD", 64, 48, sms, w, clk);
    run<21>("8 cells, selects: SEL lo + FSEL hi", 72, 40, sms, w, clk);
    run<20>("8 cells, one min only (DSETP+2FSEL)", 48, 32, sms, w, clk);
  }
  return 0;
}
