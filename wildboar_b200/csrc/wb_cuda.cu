// libwbcuda.so: C ABI (include/wb_cuda.h) + launch logic + multi-GPU row sharding.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false ...
// -fmad=false is REQUIRED: the reference build contains no FMA, and every DP value is
// reproduced bit for bit only if a*b+c is two roundings ("No FMA").
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <limits>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/wb_cuda.h"
#include "dispatch.cuh"
#include "kernels.cuh"
#include "prep.hpp"
#include "argmin.cuh"
#include "lb.cuh"
#include "dba.cuh"

namespace wb {

static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }

#define WB_CK(call)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      char buf_[512];                                                                            \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
               cudaGetErrorString(e_));                                                          \
      set_err(buf_);                                                                             \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

// Device buffers owned by one call on one device; freed (stream-ordered) on destruction.
struct Workspace {
  cudaStream_t stream;
  std::vector<void*> bufs;
  // host staging that must outlive async copies (deque: references stay valid across later emplace_back calls)
  std::deque<std::vector<double>> host_keep;
  std::deque<std::vector<float>> host_keep_f;
  std::deque<std::vector<int>> host_keep_i;
  explicit Workspace(cudaStream_t s) : stream(s) {}
  ~Workspace() { for (void* p : bufs) cudaFreeAsync(p, stream); }
  template <class T> int alloc(T** p, size_t n) {
    void* q = nullptr;
    WB_CK(cudaMallocAsync(&q, std::max<size_t>(n, 1) * sizeof(T), stream));
    bufs.push_back(q);
    *p = (T*)q;
    return 0;
  }
  // Engine scratch (row-scan rows, global boundary rings): ONE buffer reused by every DP launch of this workspace -- the
  // launches of a workspace are serialised on its stream, and chunked callers (argmin, subsequence scans) launch hundreds
  // of times, so a fresh buffer per launch would pile up gigabytes until the call ends.  Grows, never shrinks.
  void* scratch = nullptr; size_t scratch_bytes = 0;
  template <class T> int scratch_get(T** p, size_t n) {
    const size_t need = std::max<size_t>(n, 1) * sizeof(T);
    if (need > scratch_bytes) {
      void* q = nullptr;
      WB_CK(cudaMallocAsync(&q, need, stream));
      bufs.push_back(q);  // the previous, smaller one stays alive until destruction (kernels in flight may still use it)
      scratch = q; scratch_bytes = need;
    }
    *p = (T*)scratch;
    return 0;
  }
};

struct DeviceInfo { int sms; int max_smem_optin; int cc_major; int cc_minor; };
// Keep stream-ordered allocations cached in the device's default pool between calls (the
// default release threshold of 0 hands every buffer back to the OS at each synchronisation).
static void retain_pool_memory(int dev) {
  static std::atomic<unsigned long long> done{0};
  if (dev < 0 || dev >= 64 || (done.load() >> dev) & 1ULL) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done.fetch_or(1ULL << dev);
}

static int device_info(DeviceInfo* di) {
  int dev = 0;
  WB_CK(cudaGetDevice(&dev));
  retain_pool_memory(dev);
  WB_CK(cudaDeviceGetAttribute(&di->sms, cudaDevAttrMultiProcessorCount, dev));
  WB_CK(cudaDeviceGetAttribute(&di->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  WB_CK(cudaDeviceGetAttribute(&di->cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  WB_CK(cudaDeviceGetAttribute(&di->cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (di->cc_major != 10) {
    set_err("wildboar_b200 requires an sm_100 (B200) device; there is no CPU or other-arch fallback");
    return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Persistent per-device execution contexts.  A host call used to create two streams and half a dozen
// events and to query the device attributes every time (~0.2 ms: more than the kernel of a 200 x 200
// call).  Contexts are leased from a per-device free list instead (one lease per call and device, so the
// library stays re-entrant: concurrent calls get different contexts) and live until the process ends.
// ------------------------------------------------------------------------------------------
struct DevCtx {
  int dev = -1;
  DeviceInfo di{};
  cudaStream_t st = nullptr, cst = nullptr;  // compute / copy-back streams
  cudaEvent_t done[2]{}, drained[2]{}, k0[2]{}, k1[2]{};  // slab pipeline (device_worker)
};
static std::mutex g_ctx_mu;
static std::vector<DevCtx*> g_ctx_free[64];

static DevCtx* ctx_acquire(int dev) {
  if (cudaSetDevice(dev) != cudaSuccess) { set_err("cudaSetDevice failed"); cudaGetLastError(); return nullptr; }
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!g_ctx_free[dev].empty()) { DevCtx* c = g_ctx_free[dev].back(); g_ctx_free[dev].pop_back(); return c; }
  }
  DevCtx* c = new DevCtx;
  c->dev = dev;
  if (device_info(&c->di)) { delete c; return nullptr; }
  bool ok = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->cst, cudaStreamNonBlocking) == cudaSuccess;
  for (int b = 0; b < 2 && ok; ++b)
    ok = cudaEventCreateWithFlags(&c->done[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->drained[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreate(&c->k0[b]) == cudaSuccess && cudaEventCreate(&c->k1[b]) == cudaSuccess;
  if (!ok) { set_err("could not create the per-device streams / events"); cudaGetLastError(); delete c; return nullptr; }
  return c;
}
static void ctx_release(DevCtx* c) {
  if (!c) return;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (c->dev >= 0 && c->dev < 64) g_ctx_free[c->dev].push_back(c);
}
struct CtxLease {
  DevCtx* c;
  explicit CtxLease(int dev) : c(ctx_acquire(dev)) {}
  ~CtxLease() { ctx_release(c); }
  CtxLease(const CtxLease&) = delete;
  CtxLease& operator=(const CtxLease&) = delete;
};

// ------------------------------------------------------------------------------------------
// Pinned host memory for results (wb_cuda_host_alloc / wb_cuda_host_free, include/wb_cuda.h).  A device-to-host
// copy into pageable memory is staged by the driver through its own bounce buffers (~10 GB/s and
// synchronous); into page-locked memory it is one DMA at PCIe speed and truly asynchronous.  Page-locking is
// expensive (~0.3 ms per MB), so released blocks are kept in a pool (bounded by WILDBOAR_CUDA_PINNED_POOL_MB,
// default 4096) and handed out again to later calls of similar size.
// ------------------------------------------------------------------------------------------
// (the pool lives in a heap object that is never destroyed: a background page-locking thread may still hold the mutex
// while the process runs its static destructors)
struct PinPool {
  std::mutex mu;
  std::map<void*, size_t> live;         // blocks handed out: capacity
  std::multimap<size_t, void*> free_;   // pooled blocks by capacity
  size_t pooled = 0;
  size_t pending = 0;                   // large blocks being page-locked in the background
};
static PinPool& pin_pool() { static PinPool* p = new PinPool; return *p; }
#define g_pin_mu (pin_pool().mu)
#define g_pin_live (pin_pool().live)
#define g_pin_free (pin_pool().free_)
#define g_pin_pooled (pin_pool().pooled)
#define g_pin_pending (pin_pool().pending)
static size_t pin_pool_cap() {
  static size_t cap = [] {
    const char* e = getenv("WILDBOAR_CUDA_PINNED_POOL_MB");
    const long long mb = e ? atoll(e) : 4096;
    return (size_t)std::max<long long>(mb, 0) << 20;
  }();
  return cap;
}
static void* pinned_alloc(size_t bytes) {
  if (bytes == 0) bytes = 1;
  const size_t cap = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);  // 2 MB granules
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pin_free.lower_bound(cap);
    if (it != g_pin_free.end() && it->first <= cap + cap / 4 + ((size_t)8 << 20)) {
      void* p = it->second;
      g_pin_live[p] = it->first;
      g_pin_pooled -= it->first;
      g_pin_free.erase(it);
      return p;
    }
  }
  // Page-locking costs ~0.5 ms per MB (0.4 s for the 800 MB matrix of cfg3) -- more than the pageable copy path loses on
  // a result that size.  Large blocks are therefore never allocated on the caller's time: the first request is answered
  // with "none" (the caller uses ordinary memory), a background thread locks a block of that size into the pool, and the
  // following calls find it there.
  if (cap > ((size_t)256 << 20)) {
    if (cap <= pin_pool_cap()) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
      std::thread([cap, dev]() {
        if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return; }  // the caller's device: no stray context elsewhere
        {
          std::lock_guard<std::mutex> lk(g_pin_mu);
          if (g_pin_pooled + g_pin_pending + cap > pin_pool_cap()) return;
          g_pin_pending += cap;
        }
        void* q = nullptr;
        const bool ok = cudaHostAlloc(&q, cap, cudaHostAllocPortable) == cudaSuccess;
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pin_pending -= cap;
        if (ok) { g_pin_free.emplace(cap, q); g_pin_pooled += cap; }
      }).detach();
    }
    return nullptr;
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pin_live[p] = cap;
  return p;
}
static void pinned_free(void* p) {
  if (!p) return;
  size_t cap = 0;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pin_live.find(p);
    if (it == g_pin_live.end()) return;  // not ours
    cap = it->second;
    g_pin_live.erase(it);
    if (g_pin_pooled + cap <= pin_pool_cap()) { g_pin_free.emplace(cap, p); g_pin_pooled += cap; return; }
  }
  cudaFreeHost(p);
}
// page-locked (cudaHostAlloc / cudaHostRegister) host memory?
static bool is_pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// DP cells the reference evaluates per pair: sum_i (j_stop(i) - j_start(i)).
static int64_t cells_per_pair(int Tx, int Ty, int R) {
  Geom g = make_geom(Tx, Ty, R);
  int64_t c = 0;
  for (int i = 0; i < Tx; ++i) {
    int js = std::max(0, i - g.a), je = std::min(Ty, i + g.max_len);
    if (je > js) c += je - js;
  }
  return c;
}

// The global boundary buffers are written once and read once one strip later; between the two, every other
// warp's buffer traffic and the streaming result stores pass through L2.  Marking the buffer range as an
// L2 PERSISTING access-policy window (misses: streaming) keeps the dirty lines from being evicted to HBM
// before they are consumed (profiles/r01d_l2_persist.md).  WILDBOAR_CUDA_L2_PERSIST=0 disables it.
static bool l2_persist_window(cudaStream_t st, void* base, size_t bytes) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("WILDBOAR_CUDA_L2_PERSIST"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return false;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof attr);
  if (base == nullptr || bytes == 0) {  // reset
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    return false;
  }
  int dev = 0, max_persist = 0, max_window = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  if (max_persist <= 0 || max_window <= 0) return false;
  static std::atomic<unsigned long long> limit_set{0};
  if (dev < 64 && !((limit_set.load() >> dev) & 1ULL)) {
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
    limit_set.fetch_or(1ULL << dev);
  }
  // only when the whole buffer set fits the persisting carve-out: a partial window (T = 4096: 256 MB of
  // buffers) takes L2 away from the rest and measured 4 % slower
  if (bytes > (size_t)max_persist || bytes > (size_t)max_window) return false;
  const size_t win = bytes;
  attr.accessPolicyWindow.base_ptr = base;
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)win);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

// Strip-kernel configurations (measured on B200: profiles/r01c_variants.md).
//   NARROW (H < 32)       : W = 8 (W = 4 for 8 <= H < 16), NR = 2, boundary buffers in SHARED memory,
//                           two CTAs of 8 warps per SM.
//   L2 (H >= 32, <= 230 slots): boundary buffers in GLOBAL memory (L2 resident: 148 SMs x warps x slots x
//                           256 B stays well under the 126 MB L2), per-policy (W, NR, warps).
//   TALL (> 230 slots)    : global buffers, W = 8, NR = 4, 16 warps per SM (the buffers no longer fit
//                           in L2; more warps hide the extra latency).
// Shared-memory buffers would cap cfg3 at 6-7 warps per SM (33 KB per warp); the global buffers
// lift that to 12-16 and are what makes T = 4096 bands (110 KB per warp) run at all.
// (W, NR, warps) per policy from the sweep of profiles/r01h_variants_metrics.md: W = 10, NR = 6 everywhere; 14 warps at
// 128 registers, or 12 warps at 168 registers for the cells that need the registers (twe, msm, edr).
// TALL bands (T = 4096 sweep, profiles/r01h_variants_cfg5_filled.log): twe W = 12, NR = 6, 12 warps; msm W = 8, NR = 4, 12 warps.
template <class M> struct StripCfg { static constexpr int WL = 10, NRL = 6, NWL = 14, WT = 8, NRT = 4, NWT = 16; };
template <> struct StripCfg<TwePolicy> { static constexpr int WL = 10, NRL = 6, NWL = 12, WT = 12, NRT = 6, NWT = 12; };
template <> struct StripCfg<MsmPolicy> { static constexpr int WL = 10, NRL = 6, NWL = 12, WT = 8, NRT = 4, NWT = 12; };
template <> struct StripCfg<EdrPolicy> { static constexpr int WL = 10, NRL = 6, NWL = 12, WT = 8, NRT = 4, NWT = 16; };

// Launch shape for fewer warp tasks than resident warp slots.  Filling whole CTAs one after the other (grid =
// ceil(tasks / warps per CTA)) leaves the SMs unevenly loaded -- cfg1: 1250 tasks in 157 CTAs of 8 warps puts 16 warps on
// nine SMs and 8 on the others, and the launch lasts as long as the fullest SM.  Instead: as few CTA layers as needed
// (grid = a multiple of the SM count, the block scheduler deals CTAs round-robin over the SMs) and only as many warps per
// CTA as the tasks need, so every SM gets the same number of warps (+-1).
static void balance_launch(long long ntasks, int sms, int per_sm, int max_warps, long long* grid, int* warps) {
  const long long slots = (long long)sms * per_sm * max_warps;
  if (ntasks >= slots) { *grid = (long long)sms * per_sm; *warps = max_warps; return; }
  if (ntasks <= sms) { *grid = std::max<long long>(ntasks, 1); *warps = 1; return; }
  const long long layers = std::min<long long>(per_sm, (ntasks + (long long)sms * max_warps - 1) / ((long long)sms * max_warps));
  *grid = (long long)sms * layers;
  *warps = (int)std::min<long long>(max_warps, (ntasks + *grid - 1) / *grid);
}

template <class M, int W, int NT, int MINB, bool EA, int NR, bool GRING>
static int launch_strip_cfg(Workspace& ws, KArgsT<typename M::real> a, const M& m, int nwarps, int sms, size_t smem_cap, wb_stats* cfg) {
  using F = typename M::real;
  cudaStream_t st = ws.stream;
  auto kern = k_strip<M, W, NT, MINB, EA, NR, GRING>;
  a.NS = strip_ring_slots(a.g, W);
  const size_t per_warp = (size_t)a.NS * 32 * sizeof(F);
  if (!GRING) nwarps = (int)std::max<size_t>(1, std::min<size_t>((size_t)nwarps, smem_cap / per_warp));
  size_t smem = GRING ? 0 : (size_t)nwarps * per_warp;
  if (!GRING) WB_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  WB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem));
  if (per_sm < 1) { set_err("strip kernel does not fit on an SM"); return 1; }
  per_sm = std::min(per_sm, MINB);
  long long grid = (long long)sms * per_sm;
  if (a.ntasks < grid * nwarps) {
    // (WILDBOAR_CUDA_TASK_LANES = lt < 32 puts fewer pairs into a warp task -- more, emptier warps.  Measured on cfg1,
    // profiles/r02r_task_lanes_cfg1.txt: lt = 32 0.132 ms, 24 0.164 ms, 16 0.236 ms: the launch is bound by instructions
    // issued, not by latency, so the knob stays a knob.)
    if (a.mode == PM_PAIRWISE || a.mode == PM_SELF || a.mode == PM_PAIRED) {
      static const int lt_env = [] { const char* e = getenv("WILDBOAR_CUDA_TASK_LANES"); return e ? atoi(e) : 0; }();
      if (lt_env >= 1 && lt_env < 32) {
        const long long n_in_task_dim = (a.mode == PM_PAIRED) ? a.nx : a.ny;
        a.lt = lt_env;
        a.nyb = (a.ny + lt_env - 1) / lt_env;
        a.ntasks = ((a.mode == PM_PAIRED) ? 1 : a.nx) * ((n_in_task_dim + lt_env - 1) / lt_env);
      }
    }
    balance_launch(a.ntasks, sms, per_sm, nwarps, &grid, &nwarps);
    if (!GRING) smem = (size_t)nwarps * per_warp;
  }
  bool window_set = false;
  if (GRING) {
    F* ring = nullptr;
    const size_t ring_bytes = (size_t)grid * nwarps * a.NS * 32 * sizeof(F);
    if (ws.scratch_get(&ring, (size_t)grid * nwarps * a.NS * 32)) return 1;
    a.gring = ring;
    window_set = l2_persist_window(st, ring, ring_bytes);
  }
  if (cfg) { cfg->strip_w = W; cfg->strip_nr = NR; cfg->strip_warps = nwarps; cfg->strip_gring = GRING ? 1 : 0; }
  kern<<<(unsigned)grid, nwarps * 32, smem, st>>>(a, m);
  WB_CK(cudaGetLastError());
  if (window_set) l2_persist_window(st, nullptr, 0);
  return 0;
}

template <class M, bool EA>
static int launch_strip(Workspace& ws, const KArgsT<typename M::real>& a, const M& m, size_t smem_cap, int sms, wb_stats* cfg) {
  using C = StripCfg<M>;
  const size_t per_warp = (size_t)strip_ring_slots(a.g, 16) * 32 * sizeof(typename M::real);  // widest strips
  if constexpr (sizeof(typename M::real) == 4) {
    // optional fp32 mode (profiles/r01d_variants_fp32.md): half-size boundary buffers -> shared memory holds
    // 12 warps for cfg3-like bands; otherwise global buffers with wider strips
    if (a.g.H >= 32) {
      const size_t pw12 = (size_t)strip_ring_slots(a.g, 12) * 32 * sizeof(float);
      const size_t pw16 = (size_t)strip_ring_slots(a.g, 16) * 32 * sizeof(float);
      const size_t ring_budget = (size_t)8 << 30;
      const int cap = (int)std::max<size_t>(1, ring_budget / (pw16 * (size_t)sms));
      if (std::is_same<M, DtwPolicy<false, false, float>>::value && pw12 * 12 <= smem_cap && a.g.H >= 24)
        return launch_strip_cfg<M, 12, 384, 1, EA, 6, false>(ws, a, m, 12, sms, smem_cap, cfg);
      if (strip_ring_slots(a.g, 16) <= 230) return launch_strip_cfg<M, 16, 512, 1, EA, 6, true>(ws, a, m, std::min(16, cap), sms, smem_cap, cfg);
      return launch_strip_cfg<M, 16, 512, 1, EA, 4, true>(ws, a, m, std::min(16, cap), sms, smem_cap, cfg);
    }
    if (a.g.H >= 16 || a.g.H < 8) return launch_strip_cfg<M, 8, 256, 2, EA, 2, false>(ws, a, m, 8, sms, smem_cap, cfg);
    return launch_strip_cfg<M, 4, 256, 2, EA, 2, false>(ws, a, m, 8, sms, smem_cap, cfg);
  }
  if (a.g.H >= 32) {
    // keep the global rings of one launch below ~8 GB whatever the series length
    const size_t ring_budget = (size_t)8 << 30;
    int cap = (int)std::max<size_t>(1, ring_budget / (per_warp * (size_t)sms));
    const bool fits_l2 = strip_ring_slots(a.g, C::WL) <= 230;  // slots are bounded by Tx as well as by H
    if (fits_l2 && a.g.H >= 2 * C::WL)
      return launch_strip_cfg<M, C::WL, C::NWL * 32, 1, EA, C::NRL, true>(ws, a, m, std::min(C::NWL, cap), sms, smem_cap, cfg);
    if (!fits_l2)
      return launch_strip_cfg<M, C::WT, C::NWT * 32, 1, EA, C::NRT, true>(ws, a, m, std::min(C::NWT, cap), sms, smem_cap, cfg);
    return launch_strip_cfg<M, 8, 256, 2, EA, 2, true>(ws, a, m, std::min(8, cap), sms, smem_cap, cfg);
  }
  // (a deeper in-thread wavefront, NR = 4, for launches with fewer warp tasks than warp slots was measured on cfg1:
  // 0.175 ms against 0.163 ms with NR = 2 -- profiles/r02e_engines_small.jsonl -- and dropped)
  static const int narrow_w = [] { const char* e = getenv("WILDBOAR_CUDA_STRIP_NARROW_W"); return e ? atoi(e) : 0; }();  // tuning knob
  if (narrow_w == 4 && a.g.H >= 8) return launch_strip_cfg<M, 4, 256, 2, EA, 2, false>(ws, a, m, 8, sms, smem_cap, cfg);
  if (a.g.H >= 16 || a.g.H < 8) return launch_strip_cfg<M, 8, 256, 2, EA, 2, false>(ws, a, m, 8, sms, smem_cap, cfg);
  return launch_strip_cfg<M, 4, 256, 2, EA, 2, false>(ws, a, m, 8, sms, smem_cap, cfg);
}

// Cooperative engine (engine_coop.cuh, k_coop): lanes-per-pair G and cells-per-lane W for a geometry.  Three
// instantiations: W = 8 (256 threads, two CTAs per SM) for bands up to 256 coordinates, W = 13 (384 threads, one CTA)
// up to 416, W = 4 for the narrow bands the other two cannot tile; G = the smallest power of two that holds the layout.  Returns false when no layout tiles the band exactly.
template <class M>
static bool coop_pick(const Geom& g, int* W, int* G, CoopLayout* lay) {
  const int ws[3] = {8, 13, 4};  // W = 4 tiles every band of 12 .. 128 coordinates (the layouts of 8 and 13 leave gaps below 49)
  for (int q = 0; q < 3; ++q) {
    for (int gg = 2; gg <= 32; gg *= 2) {
      if (coop_supported<M>(g, ws[q], gg) && coop_layout(g, ws[q], gg, lay)) { *W = ws[q]; *G = gg; return true; }
    }
  }
  return false;
}

template <class M, int W, int U, int NT, int MINB>
static int launch_coop_cfg(Workspace& ws, KArgsT<typename M::real> a, const M& m, int G, const CoopLayout& lay, int sms, wb_stats* cfg) {
  auto kern = k_coop<M, W, U, NT, MINB>;
  int per_sm = 0;
  WB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0));
  if (per_sm < 1) { set_err("cooperative kernel does not fit on an SM"); return 1; }
  const long long ntasks = (a.npairs * G + 31) / 32;
  long long grid = 0;
  int nwarps = NT / 32;
  balance_launch(ntasks, sms, per_sm, NT / 32, &grid, &nwarps);
  if (cfg) { cfg->strip_w = W; cfg->strip_nr = G; cfg->strip_warps = nwarps; cfg->strip_gring = 0; }
  kern<<<(unsigned)grid, nwarps * 32, 0, ws.stream>>>(a, m, G, lay);
  WB_CK(cudaGetLastError());
  return 0;
}

// Band-register engine: run the interior rows blocked (engine_band.cuh)?  Measured (profiles/r02h_band_blocked.txt): with
// DENSE y rows (argmin: every lane walks its own row, one 32-byte sector per lane and load) the blocked rows -- one load
// per column per four rows instead of two per cell -- are 2-2.7x faster (2000 x 20 000 x 128, r = 0.05: msm 381 -> 181 ms,
// twe 399 -> 157 ms, lcss 199 -> 66 ms); with coalesced y (sliding windows read in place, interleaved materialised
// windows) loads are cheap, the kernel lives on resident warps, and the doubled register count (msm 90 -> 198) makes it
// 1.5-3x SLOWER (msm scan 39 -> 117 ms).  So: blocked exactly for dense rows.
template <class A>
static bool band_blocked_auto(const A& a, bool abandoning) {
  (void)abandoning;
  return !a.yil && a.ys == a.Ty;
}

// What one DP launch needs to know.
struct DpCall {
  int metric;
  wb_params p;
  const double* x; long long nx; int Tx;  // dense device arrays, ORIGINAL series
  const double* y; long long ny; int Ty;
  int mode;
  double* out; long long ld;
  double* out_m;       // optional (row-scan only)
  const double* thr;   // optional raw-domain abandon thresholds per x row
  long long thr_ld, thr_div;  // thr_div > 0: thresholds per (x row, group of thr_div consecutive y series), KArgs::thr_ld / thr_div
  int ea;              // eadistance() variants of R (ddtw EL:3308, edr EL:3833)
  long long row0; int mirror;
  bool need_rowmin;    // force the row-scan engine (exact replay needs row minima)
  const int2* list; const int* list_len;  // PM_LIST (argmin cascade): device-resident survivor list
  long long list_n;                       // PM_LISTP: number of list entries (known on the host)
  // prepared operands (filled by prepare_operands; reusable across chunked launches)
  const double* px; const double* py; int ptx, pty;
  const double* sx; const double* sy; const double* sy2;
  Tables tab;
  // fp32 mode (optional; p.precision == 1 and the metric has a float variant): float copies
  bool fp32;
  const float* pxf; const float* pyf;
  TablesT<float> tabf;
  int R;
  bool degenerate;     // ddtw with min(T) < 3: every distance is 0 (EL:3270)
  // multivariate dim="mean": accumulate into / scale the stored value (kernels.cuh, combine_dims)
  int acc; double div;
  // subsequence search: y series start every `ys` elements (0: dense rows), raw = no final sqrt
  long long ys; int raw;
  // optional copy of the prepared y, interleaved in groups of 32 series (k_interleave32): used by the row-scan and band
  // kernels (one thread per series of a warp task -> coalesced loads); the strip kernels keep reading `py`
  const double* pyi;
};

// Slope transforms, per-series scalars and lookup tables for one (x, y) operand pair.
static int prepare_operands(Workspace& ws, DpCall& c) {
  cudaStream_t st = ws.stream;
  const int Tmin = std::min(c.Tx, c.Ty);
  c.R = (int)compute_r(Tmin, c.p.r);
  c.px = c.x; c.py = c.y; c.ptx = c.Tx; c.pty = c.Ty;
  c.sx = c.sy = nullptr;
  c.tab.weights = c.tab.tw = nullptr;
  c.degenerate = false;
  if (is_derivative(c.metric)) {
    if (Tmin < 3) { c.degenerate = true; return 0; }
    double *dx = nullptr, *dy = nullptr;
    if (ws.alloc(&dx, (size_t)c.nx * (c.Tx - 2))) return 1;
    k_slope<<<1024, 256, 0, st>>>(c.x, c.nx, c.Tx, dx);
    if (c.y == c.x && c.ny == c.nx && c.Ty == c.Tx) dy = dx;
    else {
      if (ws.alloc(&dy, (size_t)c.ny * (c.Ty - 2))) return 1;
      k_slope<<<1024, 256, 0, st>>>(c.y, c.ny, c.Ty, dy);
    }
    WB_CK(cudaGetLastError());
    c.px = dx; c.py = dy; c.ptx = c.Tx - 2; c.pty = c.Ty - 2;
    if (c.ea) c.R = (int)compute_r(std::min(c.ptx, c.pty), c.p.r);
  }
  if (c.metric == M_EDR && c.ea) c.R = (int)compute_r(c.Tx, c.p.r);
  const int nmax = std::max(c.ptx, c.pty);
  if (c.metric == M_WDTW || c.metric == M_WLCSS || c.metric == M_WDDTW || c.metric == M_TWE) {
    // signed tables (index = i - j); kernels get the pointer to the centre entry
    const int64_t tn = c.metric == M_TWE ? nmax + 1 : nmax;
    ws.host_keep.push_back(c.metric == M_TWE ? make_tw(c.p.stiffness, tn) : make_weights(c.p.g, tn));
    std::vector<double>& h = ws.host_keep.back();
    double* d = nullptr;
    if (ws.alloc(&d, h.size())) return 1;
    WB_CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    if (c.metric == M_TWE) c.tab.tw = d + table_center(tn); else c.tab.weights = d + table_center(tn);
  }
  if (c.metric == M_ERP || (c.metric == M_EDR && std::isnan(c.p.epsilon))) {
    const int kind = c.metric == M_ERP ? 0 : 1;
    double *sx = nullptr, *sy = nullptr;
    if (ws.alloc(&sx, (size_t)c.nx)) return 1;
    k_series_stat<<<(unsigned)((c.nx + 127) / 128), 128, 0, st>>>(c.px, c.nx, c.ptx, kind, c.p.g, sx, c.ptx);
    if (c.py == c.px && c.ny == c.nx && c.pty == c.ptx) sy = sx;
    else {
      if (ws.alloc(&sy, (size_t)c.ny)) return 1;
      k_series_stat<<<(unsigned)((c.ny + 127) / 128), 128, 0, st>>>(c.py, c.ny, c.pty, kind, c.p.g, sy, c.pty);
    }
    WB_CK(cudaGetLastError());
    c.sx = sx; c.sy = sy;
  }
  c.fp32 = c.p.precision == 1 && has_fp32_variant(c.metric);
  c.pxf = c.pyf = nullptr; c.tabf.weights = c.tabf.tw = nullptr;
  if (c.fp32) {
    float *fx = nullptr, *fy = nullptr;
    if (ws.alloc(&fx, (size_t)c.nx * c.ptx)) return 1;
    k_to_float<<<1024, 256, 0, st>>>(c.px, c.nx * (long long)c.ptx, fx);
    if (c.py == c.px && c.ny == c.nx && c.pty == c.ptx) fy = fx;
    else {
      if (ws.alloc(&fy, (size_t)c.ny * c.pty)) return 1;
      k_to_float<<<1024, 256, 0, st>>>(c.py, c.ny * (long long)c.pty, fy);
    }
    WB_CK(cudaGetLastError());
    c.pxf = fx; c.pyf = fy;
    if (c.tab.weights || c.tab.tw) {
      const std::vector<double>& h = ws.host_keep.back();
      ws.host_keep_f.emplace_back(h.begin(), h.end());
      std::vector<float>& hf = ws.host_keep_f.back();
      float* d = nullptr;
      if (ws.alloc(&d, hf.size())) return 1;
      WB_CK(cudaMemcpyAsync(d, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice, st));
      const int64_t tn = c.metric == M_TWE ? nmax + 1 : nmax;
      if (c.metric == M_TWE) c.tabf.tw = d + table_center(tn); else c.tabf.weights = d + table_center(tn);
    }
  }
  return 0;
}

// Dense prepared y rows -> groups of 32 interleaved series for the thread-per-series kernels that walk their series row
// by row (row-scan, band): many short dense rows are otherwise read with one 32-byte sector per lane and load.
static int interleave_y(Workspace& ws, DpCall& c) {
  c.pyi = nullptr;
  if (c.degenerate || c.fp32 || c.ys != 0 || c.ny < 32 || c.pty < 1) return 0;
  double* d = nullptr;
  if (ws.alloc(&d, (size_t)interleave32_size(c.ny, c.pty))) return 1;
  k_interleave32<<<148 * 8, 256, 0, ws.stream>>>(c.py, c.ny, c.pty, d);
  WB_CK(cudaGetLastError());
  c.pyi = d;
  return 0;
}

// Launch the DP over rows [r0, r0+nrows) of the prepared x against columns [c0, c0+ncols) of
// the prepared y.  out/out_m/thr are indexed relative to (r0, c0) / r0.
template <class F>
static int launch_dp_t(Workspace& ws, const DeviceInfo& di, const DpCall& c, long long r0, long long nrows,
                       long long c0, long long ncols, double* out, long long ld, double* out_m, const double* thr,
                       wb_stats* stats) {
  cudaStream_t st = ws.stream;
  if (nrows <= 0 || ncols <= 0) return 0;
  if (c.degenerate) {
    if (c.acc) return 0;  // every dimension is degenerate alike: the zeros of dimension 0 stand (0 / n_dims == 0)
    if (c.mode == PM_PAIRED) WB_CK(cudaMemsetAsync(out, 0, sizeof(double) * nrows, st));
    else WB_CK(cudaMemset2DAsync(out, ld * sizeof(double), 0, ncols * sizeof(double), nrows, st));
    return 0;
  }
  constexpr bool kF32 = sizeof(F) == 4;
  KArgsT<F> a;
  memset(&a, 0, sizeof a);
  const long long ystep = c.ys > 0 ? c.ys : c.pty;
  if constexpr (kF32) { a.x = c.pxf + r0 * c.ptx; a.y = c.pyf + c0 * ystep; }
  else { a.x = c.px + r0 * c.ptx; a.y = c.py + c0 * ystep; }
  a.nx = nrows; a.ny = ncols; a.Tx = c.ptx; a.Ty = c.pty;
  a.g = make_geom(c.ptx, c.pty, c.R);
  a.g.raw = c.raw;
  a.ys = c.ys > 0 ? c.ys : c.pty;
  a.sx = c.sx ? c.sx + r0 : nullptr; a.sy = c.sy ? c.sy + c0 : nullptr; a.sy2 = c.sy2 ? c.sy2 + c0 : nullptr;
  a.out = out; a.ld = ld; a.out_m = out_m; a.thr = thr;
  a.thr_ld = c.thr_ld; a.thr_div = thr ? c.thr_div : 0;
  a.mode = c.mode; a.row0 = c.row0 + r0 - c0; a.mirror = c.mirror;
  a.list = c.list; a.list_len = c.list_len;
  a.acc = c.acc; a.div = c.div;
  a.lt = 32;
  a.nyb = (ncols + 31) / 32;
  a.yil = 0;
  a.ntasks = (c.mode == PM_PAIRED) ? (nrows + 31) / 32 : nrows * a.nyb;
  if (c.mode == PM_LISTP) a.ntasks = (c.list_n + 31) / 32;
  unsigned long long* counter = nullptr;
  if (ws.alloc(&counter, 1)) return 1;
  WB_CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
  a.counter = counter;

  const size_t smem_cap = (size_t)di.max_smem_optin;
  int engine = 0;
  int rc = 0;
  auto body = [&](auto m) {
    using M = decltype(m);
    bool strip_ok = strip_supported<M>(a.g, 4) && !c.need_rowmin &&
                    !(out_m != nullptr) && c.p.engine != 1 && c.p.engine != 3 && c.p.engine != 4;
    if (thr && !M::kColumnMinBound) strip_ok = false;  // exact abandoning needs row minima
    if (c.p.engine == 2 && !strip_ok) { set_err("strip engine forced but not applicable"); rc = 1; return; }
    if constexpr (!kF32) {
      if (!strip_ok && c.pyi && c0 == 0 && c.ys == 0) { a.y = c.pyi; a.yil = 1; }
    }
    // Cooperative engine (a group of lanes per pair, band row in registers): forced (engine 4), or chosen when a thread
    // per pair is the wrong shape -- (1) tall bands whose per-pair boundary buffers would spill out of L2 (T = 4096,
    // r = 0.05: 256 MB), (2) too few pairs to give every scheduler a few warps of 32 pairs.
    if constexpr (!kF32) {
      const bool plain = !c.need_rowmin && out_m == nullptr && thr == nullptr && c.mode != PM_LIST && !a.yil;
      if (plain && (c.p.engine == 0 || c.p.engine == 4)) {
        int cw = 0, cg = 0;
        CoopLayout lay;
        const long long npairs = (c.mode == PM_PAIRED) ? nrows : (c.mode == PM_LISTP ? c.list_n : nrows * ncols);
        bool use = coop_pick<M>(a.g, &cw, &cg, &lay);
        if (use && c.p.engine == 0) {
          // measured (profiles/r02d_engines_coop_v3.jsonl): with fewer thread-per-pair warps than ~4 per SM the strip
          // engine idles most schedulers (1 x 2000 x 512 dtw: 1.57 ms against 0.16 ms here); for tall bands (T = 4096,
          // r = 0.05) the DTW family is 24 % faster here (1643 against 1320 GCUPS) while msm / twe only win while the
          // strip engine is short of warps (32-row shares) -- their cells carry more context per column
          const long long tpp_warps = (npairs + 31) / 32;
          const bool tall = a.g.H >= 32 && strip_ring_slots(a.g, StripCfg<M>::WL) > 230;
          const bool few = tpp_warps < (long long)di.sms * 4;
          use = strip_ok ? (few || (tall && (M::kColumnMinBound || tpp_warps < (long long)di.sms * 8))) : false;
        }
        if (use) {
          engine = 4;
          a.npairs = npairs;
          rc = (cw == 8) ? launch_coop_cfg<M, 8, 8, 256, 2>(ws, a, m, cg, lay, di.sms, stats)
             : (cw == 4) ? launch_coop_cfg<M, 4, 8, 256, 2>(ws, a, m, cg, lay, di.sms, stats)
                         : launch_coop_cfg<M, 13, 4, 384, 1>(ws, a, m, cg, lay, di.sms, stats);
          return;
        }
        if (c.p.engine == 4) { set_err("cooperative engine forced but not applicable"); rc = 1; return; }
      } else if (c.p.engine == 4) { set_err("cooperative engine forced but not applicable"); rc = 1; return; }
    }
    if (strip_ok) {
      engine = 2;
      if constexpr (M::kColumnMinBound) {
        if (thr) { rc = launch_strip<M, true>(ws, a, m, smem_cap, di.sms, stats); return; }
      }
      rc = launch_strip<M, false>(ws, a, m, smem_cap, di.sms, stats);
    } else if (!kF32 && c.p.engine != 1 && band_supported<M>(a.g, 32)) {
      // rows in order + row minima, narrow band, equal lengths: the band lives in registers (engine_band.cuh)
      engine = 3;
      constexpr int NT = 128;
      auto go = [&](auto kern) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0) != cudaSuccess || per_sm < 1) {
          set_err("band kernel occupancy query failed"); rc = 1; return;
        }
        long long grid = (long long)di.sms * per_sm;
        const long long need = (a.ntasks * 32 + NT - 1) / NT;
        grid = std::max<long long>(1, std::min(grid, need));
        kern<<<(unsigned)grid, NT, 0, st>>>(a, m);
        if (cudaGetLastError() != cudaSuccess) { set_err("band kernel launch failed"); rc = 1; }
      };
      // blocked interior rows (engine_band.cuh): about half the instructions per cell but twice the registers, i.e. half
      // the resident warps -- WILDBOAR_CUDA_BAND_BLOCKED = 0 / 1 forces, default: see band_blocked_auto
      static const int blk_env = [] { const char* e = getenv("WILDBOAR_CUDA_BAND_BLOCKED"); return e ? atoi(e) : -1; }();
      const bool blk = a.g.H <= 16 && (blk_env >= 0 ? blk_env != 0 : band_blocked_auto(a, thr != nullptr));
      if (a.yil) {
        if constexpr (!kF32) {
          if (a.g.H <= 8) { if (blk) go(k_band<M, 8, NT, 32, true>); else go(k_band<M, 8, NT, 32>); }
          else if (a.g.H <= 16) { if (blk) go(k_band<M, 16, NT, 32, true>); else go(k_band<M, 16, NT, 32>); }
          else go(k_band<M, 32, NT, 32>);
        }
      } else if (a.g.H <= 8) { if (blk) go(k_band<M, 8, NT, 1, true>); else go(k_band<M, 8, NT>); }
      else if (a.g.H <= 16) { if (blk) go(k_band<M, 16, NT, 1, true>); else go(k_band<M, 16, NT>); }
      else go(k_band<M, 32, NT>);
    } else {
      if (c.p.engine == 3) { set_err("band engine forced but not applicable"); rc = 1; return; }
      engine = 1;
      constexpr int NT = 128;
      int per_sm = 0;
      auto kern = k_rowscan<M, NT>;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0) != cudaSuccess || per_sm < 1) {
        set_err("row-scan kernel occupancy query failed"); rc = 1; return;
      }
      per_sm = std::min(per_sm, 4);
      long long grid = (long long)di.sms * per_sm;
      long long need = (a.ntasks * 32 + NT - 1) / NT;
      grid = std::max<long long>(1, std::min(grid, need));
      a.srows = std::max(c.ptx, c.pty) + 1;
      a.sstride = grid * NT;
      F* scratch = nullptr;
      if (ws.scratch_get(&scratch, (size_t)2 * a.srows * a.sstride)) { rc = 1; return; }
      a.scratch = scratch;
      bool launched = false;
      if constexpr (!kF32) {
        if (a.yil) { k_rowscan<M, NT, 32><<<(unsigned)grid, NT, 0, st>>>(a, m); launched = true; }
      }
      if (!launched) kern<<<(unsigned)grid, NT, 0, st>>>(a, m);
      if (cudaGetLastError() != cudaSuccess) { set_err("row-scan kernel launch failed"); rc = 1; }
    }
  };
  bool known;
  if constexpr (kF32) known = with_policy_f32(c.metric, c.p, c.tabf, body);
  else known = with_policy(c.metric, c.p, c.tab, body);
  if (!known) { set_err("unknown metric id"); return 1; }
  if (rc) return rc;
  if (stats) {
    stats->engine = engine;
    stats->launches += 1;
    long long pairs = (c.mode == PM_PAIRED) ? nrows : nrows * ncols;
    if (c.mode == PM_LISTP) pairs = c.list_n;
    if (c.mode == PM_SELF) {
      pairs = 0;
      for (long long i = 0; i < nrows; ++i) {
        long long ig = a.row0 + i;
        pairs += std::max<long long>(0, ncols - 1 - ig);
      }
    }
    stats->pairs += pairs;
    stats->cells += pairs * cells_per_pair(c.ptx, c.pty, c.R);
  }
  return 0;
}

static int launch_dp(Workspace& ws, const DeviceInfo& di, const DpCall& c, long long r0, long long nrows,
                     long long c0, long long ncols, double* out, long long ld, double* out_m, const double* thr,
                     wb_stats* stats) {
  if (c.fp32) return launch_dp_t<float>(ws, di, c, r0, nrows, c0, ncols, out, ld, out_m, thr, stats);
  return launch_dp_t<double>(ws, di, c, r0, nrows, c0, ncols, out, ld, out_m, thr, stats);
}

// ------------------------------------------------------------------------------------------
// Host-buffer drivers: stage, shard rows over devices, gather.
// ------------------------------------------------------------------------------------------
}  // namespace wb
// Device-resident fitted set (include/wb_cuda.h, wb_cuda_fit): a dense (n_dims, n, T) copy per device.
struct wb_fitted {
  int64_t n, nd, T;
  std::vector<int> devs;
  std::vector<double*> ptr;
  // per device: the reference-side operands of argmin's LB cascade, built on first use (argmin.cuh, LbCascCache)
  mutable std::deque<wb::LbCascCache> casc;
};
namespace wb {

struct HostJob {
  int kind;  // 0 pairwise, 1 self, 2 paired, 3 argmin
  int metric; wb_params p;
  const double* x; int64_t nx, Tx, xs;
  const double* y; int64_t ny, Ty, ys;
  double* out;
  int64_t k; const double* lower_bound; int use_device_lb; int64_t* out_idx;
  // multivariate (pairwise / self / paired): n_dims >= 1 dimensions, `xds` / `yds` elements between the
  // dimensions of one sample; combine 0 = "mean" (one matrix), 1 = "full" (n_dims matrices)  (DI:1289-1297)
  int64_t nd, xds, yds; int combine;
  const wb_fitted* fit;  // kinds 0 / 3: the second operand is already resident on the devices
  int* self_mirrored;    // kind 1, in/out: non-null = the worker MAY mirror the lower triangle on the device; set to 1 when it did
};

static int h2d_rows(double* dst, const double* src, int64_t rows, int64_t T, int64_t stride, cudaStream_t st) {
  if (rows <= 0) return 0;
  if (stride == T) WB_CK(cudaMemcpyAsync(dst, src, sizeof(double) * rows * T, cudaMemcpyHostToDevice, st));
  else WB_CK(cudaMemcpy2DAsync(dst, sizeof(double) * T, src, sizeof(double) * stride, sizeof(double) * T, rows,
                               cudaMemcpyHostToDevice, st));
  return 0;
}

// h2d_rows for LARGE pageable sources: piece by piece through two page-locked staging buffers that several host threads
// fill (stage.hpp) while the DMA of the previous piece runs -- ~3x the rate of the driver's own pageable path (one staging
// thread).  Page-locked or small sources, or no staging memory to be had: plain h2d_rows.  Blocks until the last piece has
// left its staging buffer (the copies themselves are ordered on `st` like any other).
static int h2d_rows_staged(double* dst, const double* src, int64_t rows, int64_t T, int64_t stride, cudaStream_t st) {
  const size_t piece_bytes = (size_t)16 << 20;
  if (rows <= 0) return 0;
  if ((size_t)rows * T * sizeof(double) < 2 * piece_bytes || is_pinned_host(src) || getenv("WILDBOAR_CUDA_NO_STAGING"))
    return h2d_rows(dst, src, rows, T, stride, st);
  char* buf[2] = {(char*)pinned_alloc(piece_bytes), nullptr};
  buf[1] = buf[0] ? (char*)pinned_alloc(piece_bytes) : nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ok = buf[0] && buf[1] && cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess;
  int rc = 0;
  if (!ok) rc = h2d_rows(dst, src, rows, T, stride, st);
  else {
    const int64_t per = std::max<int64_t>(1, (int64_t)(piece_bytes / (sizeof(double) * (size_t)T)));
    bool used[2] = {false, false};
    int n = 0;
    for (int64_t r0 = 0; r0 < rows && !rc; r0 += per, ++n) {
      const int64_t nr = std::min(per, rows - r0);
      const int b = n & 1;
      if (used[b] && cudaEventSynchronize(ev[b]) != cudaSuccess) { rc = 1; break; }
      HostCopyPool::get().copy_rows(buf[b], (const char*)(src + r0 * stride), (size_t)nr, sizeof(double) * (size_t)T, sizeof(double) * (size_t)stride);
      if (cudaMemcpyAsync(dst + r0 * T, buf[b], sizeof(double) * nr * T, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaEventRecord(ev[b], st) != cudaSuccess) { set_err("staged host-to-device copy failed"); rc = 1; break; }
      used[b] = true;
    }
    for (int b = 0; b < 2; ++b) if (used[b]) cudaEventSynchronize(ev[b]);
  }
  for (int b = 0; b < 2; ++b) { if (ev[b]) cudaEventDestroy(ev[b]); if (buf[b]) pinned_free(buf[b]); }
  return rc;
}

struct Timer {
  cudaEvent_t a, b; cudaStream_t st; bool ok;
  explicit Timer(cudaStream_t s) : st(s) { ok = cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess; }
  ~Timer() { if (ok) { cudaEventDestroy(a); cudaEventDestroy(b); } }
  void start() { if (ok) cudaEventRecord(a, st); }
  void stop() { if (ok) cudaEventRecord(b, st); }
  double ms() { float f = 0; if (ok && cudaEventSynchronize(b) == cudaSuccess) cudaEventElapsedTime(&f, a, b); return f; }
};

// rows [lo, hi) of the job on device `dev`
static int device_worker(const HostJob& J, int dev, int64_t lo, int64_t hi, wb_stats* st_out) {
  CtxLease lease(dev);  // persistent streams / events of this device (one lease per call: re-entrant)
  if (!lease.c) return 1;
  DevCtx& cx = *lease.c;
  const DeviceInfo di = cx.di;
  cudaStream_t st = cx.st, cst = cx.cst;
  int rc = 0;
  wb_stats stats;
  memset(&stats, 0, sizeof stats);
  {
    Workspace ws(st);
    Timer total(st);
    total.start();
    const int64_t rows = hi - lo;
    const int64_t nd = std::max<int64_t>(J.nd, 1);
    double *dx = nullptr, *dy = nullptr;
    bool argmin_piped = false;
    do {
      // operands are staged densely as (n_dims, samples, T)
      int64_t xn = 0, xT = 0, yn = 0, yT = 0;  // samples / length of the staged first and second operand
      if (J.kind == 1) {
        // self join: every device holds all of x (columns); its rows are [lo, hi)
        yn = J.nx; yT = J.Tx; xn = J.nx; xT = J.Tx;
        if ((rc = ws.alloc(&dy, (size_t)nd * J.nx * J.Tx))) break;
        for (int64_t d = 0; d < nd && !rc; ++d) rc = h2d_rows(dy + d * J.nx * J.Tx, J.x + d * J.xds, J.nx, J.Tx, J.xs, st);
        if (rc) break;
        dx = dy;
      } else if (J.kind == 2) {
        // paired: first operand is the USER's y (CD:1632-1647)
        xn = rows; xT = J.Ty; yn = rows; yT = J.Tx;
        if ((rc = ws.alloc(&dx, (size_t)nd * rows * J.Ty))) break;
        if ((rc = ws.alloc(&dy, (size_t)nd * rows * J.Tx))) break;
        for (int64_t d = 0; d < nd && !rc; ++d) {
          if ((rc = h2d_rows(dx + d * rows * J.Ty, J.y + d * J.yds + lo * J.ys, rows, J.Ty, J.ys, st))) break;
          rc = h2d_rows(dy + d * rows * J.Tx, J.x + d * J.xds + lo * J.xs, rows, J.Tx, J.xs, st);
        }
        if (rc) break;
      } else {
        xn = rows; xT = J.Tx; yn = J.ny; yT = J.Ty;
        if ((rc = ws.alloc(&dx, (size_t)nd * rows * J.Tx))) break;
        if (J.fit) {
          for (size_t q = 0; q < J.fit->devs.size(); ++q) if (J.fit->devs[q] == dev) dy = J.fit->ptr[q];
          if (!dy) { set_err("the fitted set is not resident on this device"); rc = 1; break; }
        } else if ((rc = ws.alloc(&dy, (size_t)nd * J.ny * J.Ty))) break;
        // argmin of dtw / wdtw / adtw (fp64) against HOST references: the rows are uploaded piecewise while the scan is
        // already running (run_argmin); every other case needs the whole second operand first (slopes, per-series
        // scalars, float copies, interleaved rows are derived from all of it)
        argmin_piped = J.kind == 3 && !J.fit && nd == 1 && J.p.precision != 1 && piped_piece_bytes() > 0 &&
                       (long long)sizeof(double) * J.ny * J.Ty > 2 * piped_piece_bytes() &&
                       (J.metric == M_DTW || J.metric == M_WDTW || J.metric == M_ADTW);
        for (int64_t d = 0; d < nd && !rc; ++d) {
          if ((rc = h2d_rows(dx + d * rows * J.Tx, J.x + d * J.xds + lo * J.xs, rows, J.Tx, J.xs, st))) break;
          if (!J.fit && !argmin_piped) rc = h2d_rows_staged(dy + d * J.ny * J.Ty, J.y + d * J.yds, J.ny, J.Ty, J.ys, st);
        }
        if (rc) break;
      }
      // one prepared call per dimension (slopes, per-series scalars and tables are per dimension)
      std::vector<DpCall> cs((size_t)nd);
      for (int64_t d = 0; d < nd; ++d) {
        DpCall& c = cs[(size_t)d];
        memset(&c, 0, sizeof c);
        c.metric = J.metric; c.p = J.p;
        const double* xd = dx + d * xn * xT;
        const double* yd = dy + d * yn * yT;
        if (J.kind == 2) { c.x = xd; c.nx = rows; c.Tx = (int)xT; c.y = yd; c.ny = rows; c.Ty = (int)yT; c.mode = PM_PAIRED; }
        else if (J.kind == 1) { c.x = xd + lo * J.Tx; c.nx = rows; c.Tx = (int)J.Tx; c.y = yd; c.ny = J.nx; c.Ty = (int)J.Tx; c.mode = PM_SELF; c.row0 = lo; }
        else { c.x = xd; c.nx = rows; c.Tx = (int)J.Tx; c.y = yd; c.ny = J.ny; c.Ty = (int)J.Ty; c.mode = PM_PAIRWISE; }
        if (nd > 1 && J.combine == 0) { c.acc = d > 0; c.div = (d == nd - 1) ? (double)nd : 0.0; }
      }
      DpCall& c = cs[0];

      if (J.kind == 3) {
        c.ea = 1;
        if ((rc = prepare_operands(ws, c))) break;
        ArgminIo io;
        if (J.fit) for (size_t q = 0; q < J.fit->devs.size() && q < J.fit->casc.size(); ++q) if (J.fit->devs[q] == dev) io.casc_cache = &J.fit->casc[q];
        io.k = J.k; io.lower_bound = J.lower_bound ? J.lower_bound + lo * J.ny : nullptr; io.lb_ld = J.ny;
        io.out_idx = J.out_idx + lo * J.k; io.out_dist = J.out + lo * J.k; io.use_device_lb = J.use_device_lb;
        char* stage_bufs[2] = {nullptr, nullptr};
        if (argmin_piped) {
          io.y_host = J.y; io.y_hs = J.ys; io.y_dev = dy; io.up_stream = cst;
          // pageable references: staged through two page-locked buffers by several copy threads (stage.hpp)
          if (!is_pinned_host(J.y) && !getenv("WILDBOAR_CUDA_NO_STAGING")) {
            const size_t pb = (size_t)piped_piece_bytes() + sizeof(double) * (size_t)J.Ty;
            stage_bufs[0] = (char*)pinned_alloc(pb);
            stage_bufs[1] = stage_bufs[0] ? (char*)pinned_alloc(pb) : nullptr;
            if (stage_bufs[0] && stage_bufs[1]) { io.stage_buf[0] = stage_bufs[0]; io.stage_buf[1] = stage_bufs[1]; }
          }
        }
        rc = run_argmin(ws, di, c, io, &stats,
                        [&](long long r0, long long nr, long long c0, long long nc, double* o, long long ld, double* om,
                            const double* thr, wb_stats* s) { return launch_dp(ws, di, c, r0, nr, c0, nc, o, ld, om, thr, s); });
        for (char* sbuf : stage_bufs) if (sbuf) pinned_free(sbuf);  // run_argmin has waited for the last DMA out of them
        if (rc) break;
        WB_CK(cudaStreamSynchronize(st));
        break;
      }

      for (int64_t d = 0; d < nd && !rc; ++d) rc = prepare_operands(ws, cs[(size_t)d]);
      if (rc) break;
      // result matrices: 1 (single dimension / "mean") or n_dims ("full"); dimensions per matrix: n_dims or 1
      const int64_t nmat = (nd > 1 && J.combine == 1) ? nd : 1;
      const int64_t dpm = nd / nmat;
      if (J.kind == 2) {
        double* dout = nullptr;
        if ((rc = ws.alloc(&dout, (size_t)rows))) break;
        for (int64_t mi = 0; mi < nmat && !rc; ++mi) {
          Timer kt(st); kt.start();
          for (int64_t d = mi * dpm; d < (mi + 1) * dpm && !rc; ++d)
            rc = launch_dp(ws, di, cs[(size_t)d], 0, rows, 0, rows, dout, 1, nullptr, nullptr, &stats);
          if (rc) break;
          kt.stop();
          WB_CK(cudaMemcpyAsync(J.out + mi * J.nx + lo, dout, sizeof(double) * rows, cudaMemcpyDeviceToHost, st));
          WB_CK(cudaStreamSynchronize(st));
          stats.kernel_ms += kt.ms();
        }
        break;
      }
      // pairwise / self: chunk the row block so result slabs stream back while the next chunk computes
      const int64_t ncols = c.ny;
      size_t slab_budget = (size_t)48 << 20;  // small slabs: the un-overlapped tail copy stays short
      if (const char* e = getenv("WILDBOAR_CUDA_SLAB_KB")) {  // test knob: forces many slabs on small inputs
        const long long v = atoll(e);
        if (v >= 1) slab_budget = (size_t)v << 10;
      }
      int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(rows, (int64_t)(slab_budget / (sizeof(double) * std::max<int64_t>(ncols, 1)))));
      const int64_t nchunks = (rows + chunk - 1) / chunk;
      const int64_t nunits = nchunks * nmat;  // streaming unit u = (matrix u / nchunks, row chunk u % nchunks)
      // page-locked result (wb_cuda_host_alloc, or registered by the caller): the slab copies are asynchronous DMAs and
      // the whole pipeline is ordered by events, the host only enqueues; pageable: the copy blocks the host anyway
      const bool out_pinned = is_pinned_host(J.out);
      // Self join on ONE device: the lower triangle is written by the kernel itself into a device-resident n x n matrix
      // (CD:1240-1246 copies it from the upper one; the values are the same doubles), so no host-side transpose pass is
      // needed.  Row chunk u is final once the kernels of chunks 0..u have run (the mirrored entries of row j come from
      // rows i < j), so the chunks still stream back in order while later chunks compute.
      bool full_self = false;
      if (J.kind == 1 && J.self_mirrored && lo == 0 && hi == J.nx && nmat == 1) {
        size_t fr = 0, tot = 0;
        const size_t need = sizeof(double) * (size_t)rows * (size_t)ncols;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && need <= fr / 2) full_self = true;
        cudaGetLastError();
      }
      double* dbuf[2] = {nullptr, nullptr};
      if (full_self) {
        if ((rc = ws.alloc(&dbuf[0], (size_t)rows * ncols))) break;
        WB_CK(cudaMemsetAsync(dbuf[0], 0, sizeof(double) * (size_t)rows * ncols, st));  // zero diagonal
        for (int64_t d = 0; d < nd; ++d) cs[(size_t)d].mirror = 1;
        *J.self_mirrored = 1;
      } else {
        if (J.kind == 1 && J.self_mirrored) *J.self_mirrored = 0;
        if ((rc = ws.alloc(&dbuf[0], (size_t)chunk * ncols))) break;
        if (nunits > 1 && (rc = ws.alloc(&dbuf[1], (size_t)chunk * ncols))) break;
      }
      bool timed[2] = {false, false};
      auto harvest = [&](int b) {  // kernel time of the unit that last used event pair b
        if (!timed[b]) return;
        float f = 0;
        if (cudaEventSynchronize(cx.k1[b]) == cudaSuccess && cudaEventElapsedTime(&f, cx.k0[b], cx.k1[b]) == cudaSuccess) stats.kernel_ms += f;
        timed[b] = false;
      };
      // software pipeline: the kernels of unit u+1 are enqueued BEFORE the copy of unit u is issued (a copy into
      // pageable memory blocks the host until the slab has been staged; the device keeps computing meanwhile)
      auto launch_unit = [&](int64_t u) -> int {
        const int b = (int)(u & 1);
        const int64_t mi = u / nchunks, ci = u % nchunks;
        const int64_t r0 = ci * chunk, nr = std::min(chunk, rows - r0);
        double* const slab = full_self ? dbuf[0] + r0 * ncols : dbuf[b];
        harvest(b);
        // the copy of unit u-2 must have left the buffer before unit u overwrites it
        if (!full_self && u >= 2) WB_CK(cudaStreamWaitEvent(st, cx.drained[b], 0));
        if (J.kind == 1 && !full_self) WB_CK(cudaMemsetAsync(slab, 0, sizeof(double) * nr * ncols, st));
        WB_CK(cudaEventRecord(cx.k0[b], st));
        for (int64_t d = mi * dpm; d < (mi + 1) * dpm; ++d)
          if (launch_dp(ws, di, cs[(size_t)d], r0, nr, 0, ncols, slab, ncols, nullptr, nullptr, &stats)) return 1;
        WB_CK(cudaEventRecord(cx.k1[b], st));
        WB_CK(cudaEventRecord(cx.done[b], st));
        timed[b] = true;
        return 0;
      };
      auto copy_unit = [&](int64_t u) -> int {
        const int b = (int)(u & 1);
        const int64_t mi = u / nchunks, ci = u % nchunks;
        const int64_t r0 = ci * chunk, nr = std::min(chunk, rows - r0);
        double* const slab = full_self ? dbuf[0] + r0 * ncols : dbuf[b];
        WB_CK(cudaStreamWaitEvent(cst, cx.done[b], 0));
        if (cudaMemcpyAsync(J.out + mi * J.nx * ncols + (lo + r0) * ncols, slab, sizeof(double) * nr * ncols, cudaMemcpyDeviceToHost, cst) != cudaSuccess) {
          set_err("device-to-host copy of the result slab failed"); return 1;
        }
        WB_CK(cudaEventRecord(cx.drained[b], cst));
        return 0;
      };
      if ((rc = launch_unit(0))) break;
      for (int64_t u = 0; u < nunits && !rc; ++u) {
        if (u + 1 < nunits && (rc = launch_unit(u + 1))) break;
        rc = copy_unit(u);
      }
      (void)out_pinned;
      if (!rc) { harvest(0); harvest(1); }
      if (cudaStreamSynchronize(cst) != cudaSuccess && !rc) { set_err("device-to-host copy of the result slab failed"); rc = 1; }
      if (rc) break;
      WB_CK(cudaStreamSynchronize(st));
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); stats.total_ms = total.ms(); }
  }
  cudaStreamSynchronize(st);
  cudaStreamSynchronize(cst);
  if (st_out) *st_out = stats;
  return rc;
}

// utils/_parallel.py:7-23: contiguous row blocks, the first (n % G) one row longer
static void row_blocks(int64_t n, int G, std::vector<int64_t>& off) {
  off.assign(G + 1, 0);
  int64_t bs = n / G, ov = n % G;
  for (int b = 0; b < G; ++b) off[b + 1] = off[b] + bs + (b < ov ? 1 : 0);
}
// self join: rows i cost (n-1-i) pairs; cut so every device gets ~equal pair counts
static void tri_blocks(int64_t n, int G, std::vector<int64_t>& off) {
  off.assign(G + 1, n);
  off[0] = 0;
  const double total = 0.5 * (double)n * (double)(n - 1);
  int64_t i = 0;
  double acc = 0;
  for (int b = 1; b < G; ++b) {
    const double target = total * b / G;
    while (i < n && acc + (double)(n - 1 - i) <= target) { acc += (double)(n - 1 - i); ++i; }
    off[b] = i;
  }
}

static int run_host_job(const HostJob& J, const int* devices, int n_devices, wb_stats* stats) {
  int ndev_avail = 0;
  if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail < 1) {
    set_err("no CUDA device available: wildboar_b200 has no CPU fallback");
    return 1;
  }
  std::vector<int> devs;
  if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
  else devs.push_back(0);
  for (int d : devs) if (d < 0 || d >= ndev_avail) { set_err("invalid device ordinal"); return 1; }
  const int64_t n_work = J.nx;
  int G = (int)std::min<int64_t>((int64_t)devs.size(), std::max<int64_t>(n_work, 1));
  std::vector<int64_t> off;
  if (J.kind == 1) tri_blocks(n_work, G, off); else row_blocks(n_work, G, off);
  std::vector<wb_stats> sts(G);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  int mirrored = 0;
  if (G == 1) {
    HostJob J1 = J;
    if (J.kind == 1) J1.self_mirrored = &mirrored;  // one device: the kernel writes the lower triangle as well
    rcs[0] = device_worker(J1, devs[0], off[0], off[1], &sts[0]);
    errs[0] = g_err;
  } else {
    std::vector<std::thread> th;
    for (int b = 0; b < G; ++b)
      th.emplace_back([&, b]() { rcs[b] = device_worker(J, devs[b], off[b], off[b + 1], &sts[b]); errs[b] = g_err; });
    for (auto& t : th) t.join();
  }
  for (int b = 0; b < G; ++b) if (rcs[b]) { set_err(errs[b]); return rcs[b]; }
  if (J.kind == 1 && !mirrored) {
    // lower triangle = copy of the upper one (CD:1240-1246): blocked transpose, destination row blocks spread over
    // host threads (disjoint writes)
    const int64_t n = J.nx, B = 64;
    const int64_t nmat = (J.nd > 1 && J.combine == 1) ? J.nd : 1;
    const int64_t nblk = (n + B - 1) / B;
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, n * n * nmat / (1 << 18)}));
    auto mirror_blocks = [&](int t) {
      for (int64_t mi = 0; mi < nmat; ++mi) {
        double* o = J.out + mi * n * n;
        // destination block row jb; block rows are dealt out round-robin (row jb costs jb + 1 blocks)
        for (int64_t jbi = t; jbi < nblk; jbi += nth) {
          const int64_t jb = jbi * B;
          for (int64_t ib = 0; ib <= jb; ib += B)
            for (int64_t i = ib; i < std::min(ib + B, n); ++i)
              for (int64_t j = std::max(jb, i + 1); j < std::min(jb + B, n); ++j) o[j * n + i] = o[i * n + j];
        }
      }
    };
    if (nth == 1) mirror_blocks(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nth; ++t) th.emplace_back(mirror_blocks, t);
      for (auto& t : th) t.join();
    }
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    for (int b = 0; b < G; ++b) {
      stats->kernel_ms = std::max(stats->kernel_ms, sts[b].kernel_ms);
      stats->total_ms = std::max(stats->total_ms, sts[b].total_ms);
      stats->cells += sts[b].cells; stats->pairs += sts[b].pairs; stats->launches += sts[b].launches;
      stats->lb_kim_pruned += sts[b].lb_kim_pruned; stats->lb_keogh_pruned += sts[b].lb_keogh_pruned;
      stats->ambiguous += sts[b].ambiguous;
      stats->engine = std::max(stats->engine, sts[b].engine);
      if (sts[b].strip_w) { stats->strip_w = sts[b].strip_w; stats->strip_nr = sts[b].strip_nr; stats->strip_warps = sts[b].strip_warps; stats->strip_gring = sts[b].strip_gring; }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Lower-bound matrices (wildboar.distance.lb transformers), one device.
// ------------------------------------------------------------------------------------------
// distance/dtw.py:38-40 + LB:367-369: envelope half-width max(floor(T r), 1), T - 1 when that equals T
static int lb_warp_size(int64_t T, double r) {
  int64_t w = (int64_t)std::floor((double)T * r);
  if (w < 1) w = 1;
  if (w == T) w -= 1;
  return (int)w;
}

// op 0: LB_Keogh (kind 0 both / 1 left / 2 right), op 1: LB_Kim
static int run_lb(int op, const double* q, int64_t nq, int64_t qs, const double* x, int64_t nx, int64_t xs, int64_t T,
                  double r, int kind, double* out, int device, wb_stats* stats) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_err("no CUDA device available: wildboar_b200 has no CPU fallback"); return 1; }
  if (device < 0 || device >= ndev) { set_err("invalid device ordinal"); return 1; }
  WB_CK(cudaSetDevice(device));
  DeviceInfo di;
  if (device_info(&di)) return 1;
  cudaStream_t st;
  WB_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int rc = 0;
  wb_stats local; memset(&local, 0, sizeof local);
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    do {
      double *dq = nullptr, *dx = nullptr;
      if ((rc = ws.alloc(&dq, (size_t)nq * T)) || (rc = ws.alloc(&dx, (size_t)nx * T))) break;
      if ((rc = h2d_rows(dq, q, nq, T, qs, st)) || (rc = h2d_rows(dx, x, nx, T, xs, st))) break;
      // result slabs of <= 256 MB, double buffered like the pairwise driver
      const int64_t chunk = std::max<int64_t>(kLbQB, std::min<int64_t>(nq, (((int64_t)256 << 20) / (8 * std::max<int64_t>(nx, 1))) / kLbQB * kLbQB));
      double* dout = nullptr;
      if ((rc = ws.alloc(&dout, (size_t)std::min<int64_t>(chunk, nq) * nx))) break;
      double *qlo = nullptr, *qhi = nullptr, *xT = nullptr, *xloT = nullptr, *xhiT = nullptr;
      kt.start();
      if (op == 0) {
        const int w = lb_warp_size(T, r);
        if ((rc = ws.alloc(&qlo, (size_t)nq * T)) || (rc = ws.alloc(&qhi, (size_t)nq * T)) || (rc = ws.alloc(&xT, (size_t)nx * T)) ||
            (rc = ws.alloc(&xloT, (size_t)nx * T)) || (rc = ws.alloc(&xhiT, (size_t)nx * T))) break;
        k_envelope_rows<<<1024, 256, 0, st>>>(dq, nq, (int)T, w, 1, nullptr, qlo, qhi);  // LB:417-418 (natural time order)
        k_envelope_T<<<2048, 256, 0, st>>>(dx, nx, (int)T, w, 1, xT, xloT, xhiT);     // fit, LB:371-374
        WB_CK(cudaGetLastError());
        local.launches += 2;
      }
      for (int64_t i0 = 0; i0 < nq && !rc; i0 += chunk) {
        const int64_t ni = std::min(chunk, nq - i0);
        if (op == 0) {
          LbMatArgs a;
          a.q = dq + i0 * T; a.qlo = qlo + i0 * T; a.qhi = qhi + i0 * T; a.xT = xT; a.xloT = xloT; a.xhiT = xhiT;
          a.nq = ni; a.nx = nx; a.T = (int)T; a.out = dout; a.ld = nx;
          const long long tasks = ((nx + kLbNT - 1) / kLbNT) * ((ni + kLbQB - 1) / kLbQB);
          const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(tasks, (long long)di.sms * 8));
          if (kind == 1) k_lb_keogh_matrix<true, false><<<grid, kLbNT, 0, st>>>(a);
          else if (kind == 2) k_lb_keogh_matrix<false, true><<<grid, kLbNT, 0, st>>>(a);
          else k_lb_keogh_matrix<true, true><<<grid, kLbNT, 0, st>>>(a);
        } else {
          k_lb_kim_matrix<<<(unsigned)std::min<long long>((ni * nx + 255) / 256, (long long)di.sms * 16), 256, 0, st>>>(dq + i0 * T, ni, dx, nx, (int)T, dout, nx);
        }
        if (cudaGetLastError() != cudaSuccess) { set_err("lower-bound kernel launch failed"); rc = 1; break; }
        local.launches += 1;
        if (cudaMemcpyAsync(out + i0 * nx, dout, sizeof(double) * ni * nx, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the lower-bound slab failed"); rc = 1; break; }
      }
      kt.stop();
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); local.total_ms = total.ms(); local.kernel_ms = kt.ms(); local.pairs = nq * nx; local.cells = nq * nx * T; }
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (stats) *stats = local;
  return rc;
}

// dtw_envelop / dtw_lb_keogh of wildboar.distance.dtw (dtw.py:155-243), batched over n series of one device call.
// op 0: lower / upper envelopes (EL:1076-1092, half-width w); op 1: per-time-step LB_Keogh terms + their root sum.
static int run_lb_series(int op, const double* x, int64_t n, int64_t T, int64_t xs, int64_t w, const double* lower,
                         const double* upper, double* out_a, double* out_b, int device, wb_stats* stats) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_err("no CUDA device available: wildboar_b200 has no CPU fallback"); return 1; }
  if (device < 0 || device >= ndev) { set_err("invalid device ordinal"); return 1; }
  CtxLease lease(device);
  if (!lease.c) return 1;
  cudaStream_t st = lease.c->st;
  wb_stats local; memset(&local, 0, sizeof local);
  int rc = 0;
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    do {
      double *dx = nullptr, *da = nullptr, *db = nullptr, *dlo = nullptr, *dhi = nullptr;
      const size_t ne = (size_t)n * (size_t)T;
      if ((rc = ws.alloc(&dx, ne)) || (rc = h2d_rows(dx, x, n, T, xs, st))) break;
      if (op == 0) {
        if ((rc = ws.alloc(&da, ne)) || (rc = ws.alloc(&db, ne))) break;
        kt.start();
        k_envelope_rows<<<(unsigned)std::min<size_t>((ne + 255) / 256, 4096), 256, 0, st>>>(dx, n, (int)T, (int)w, 1, nullptr, da, db);
        kt.stop();
        WB_CK(cudaGetLastError());
        WB_CK(cudaMemcpyAsync(out_a, da, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
        WB_CK(cudaMemcpyAsync(out_b, db, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
      } else {
        if ((rc = ws.alloc(&dlo, ne)) || (rc = ws.alloc(&dhi, ne)) || (rc = ws.alloc(&da, (size_t)n)) || (rc = ws.alloc(&db, ne))) break;
        WB_CK(cudaMemcpyAsync(dlo, lower, sizeof(double) * ne, cudaMemcpyHostToDevice, st));
        WB_CK(cudaMemcpyAsync(dhi, upper, sizeof(double) * ne, cudaMemcpyHostToDevice, st));
        kt.start();
        k_lb_keogh_terms<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(dx, dlo, dhi, n, (int)T, da, db);
        kt.stop();
        WB_CK(cudaGetLastError());
        WB_CK(cudaMemcpyAsync(out_a, da, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
        WB_CK(cudaMemcpyAsync(out_b, db, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
      }
      local.launches = 1;
      WB_CK(cudaStreamSynchronize(st));
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); local.total_ms = total.ms(); local.kernel_ms = kt.ms(); local.pairs = n; local.cells = n * T; }
  }
  cudaStreamSynchronize(st);
  if (stats) *stats = local;
  return rc;
}

// ------------------------------------------------------------------------------------------
// DTW alignment paths + DBA (kernels in dba.cuh).
// ------------------------------------------------------------------------------------------
// dtw.py:38-40 `_compute_warp_size`: max(floor(max(Ta, Tb) r), 1)
static int warp_size_max(int64_t Ta, int64_t Tb, double r) {
  return (int)std::max<int64_t>((int64_t)std::floor((double)std::max(Ta, Tb) * r), 1);
}

// Device-level: paths of n_pairs (a[ia[p]], b[ib[p]]) pairs.  d_w: centre of the signed weight table or nullptr.
static int run_paths_dev(Workspace& ws, int di_smem_cap, int di_sms, const double* d_a, int64_t Ta, const double* d_b, int64_t Tb, const int* d_ia,
                         const int* d_ib, int64_t n_pairs, int R, const double* d_w, int* d_lo, int* d_hi, double* d_cost,
                         double* d_D, wb_stats* stats) {
  cudaStream_t st = ws.stream;
  PathArgs p;
  memset(&p, 0, sizeof p);
  p.a = d_a; p.b = d_b; p.g = make_geom((int)Ta, (int)Tb, R); p.w = d_w;
  p.HB = (p.g.H + 3) / 4;
  // threads per CTA: the band row (H + 1 slots per thread) lives in shared memory when it fits
  const size_t row_bytes = (size_t)(p.g.H + 1) * sizeof(double);
  // few pairs: small CTAs spread the warps over all SMs (each warp then has an L1 and a scheduler to itself)
  int nt = 128;
  while (nt > 32 && (row_bytes * nt > (size_t)200 * 1024 || (n_pairs + nt - 1) / nt < di_sms)) nt >>= 1;
  const bool smem_ok = row_bytes * nt <= (size_t)di_smem_cap;
  const size_t smem = smem_ok ? row_bytes * nt : 0;
  if (!smem_ok) nt = 128;
  if (smem_ok) WB_CK(cudaFuncSetAttribute(k_dtw_paths<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // pairs per launch: bound the move buffer (<= 2 GB) and the fallback scratch
  const size_t per_task_moves = (size_t)Ta * p.HB * 32;
  int64_t chunk = std::max<int64_t>(32, (int64_t)(((size_t)2 << 30) / per_task_moves) * 32);
  chunk = std::min<int64_t>(chunk, 1 << 16);
  chunk = std::min<int64_t>(chunk, (n_pairs + 31) / 32 * 32);
  const int64_t threads = (chunk + nt - 1) / nt * nt;
  unsigned char* moves = nullptr; double* scratch = nullptr;
  if (ws.alloc(&moves, (size_t)(threads / 32) * per_task_moves)) return 1;
  if (!smem_ok && ws.alloc(&scratch, (size_t)(p.g.H + 1) * threads)) return 1;
  p.moves = moves; p.scratch = scratch; p.sstride = threads;
  for (int64_t p0 = 0; p0 < n_pairs; p0 += chunk) {
    const int64_t np = std::min(chunk, n_pairs - p0);
    p.n_pairs = np;
    p.ia = d_ia ? d_ia + p0 : nullptr; p.ib = d_ib ? d_ib + p0 : nullptr;
    if (!d_ia) p.a = d_a + p0 * Ta;
    if (!d_ib) p.b = d_b + p0 * Tb;
    p.lo = d_lo + p0 * Ta; p.hi = d_hi + p0 * Ta;
    p.cost = d_cost ? d_cost + p0 : nullptr;
    p.D = d_D ? d_D + p0 * Ta * Tb : nullptr;
    const unsigned grid = (unsigned)((np + nt - 1) / nt);
    if (smem_ok) k_dtw_paths<true><<<grid, nt, smem, st>>>(p);
    else k_dtw_paths<false><<<grid, nt, 0, st>>>(p);
    WB_CK(cudaGetLastError());
    if (stats) { stats->launches += 1; stats->pairs += np; stats->cells += np * cells_per_pair((int)Ta, (int)Tb, R); }
  }
  return 0;
}

// signed table t[centre + d] = weights[|d|] (prep.hpp layout) from a caller-provided weight vector
static int upload_weight_table(Workspace& ws, const double* weights, int64_t n, const double** d_center) {
  *d_center = nullptr;
  if (!weights) return 0;
  const int64_t c = table_center(n);
  ws.host_keep.emplace_back((size_t)(2 * c + 1), 0.0);
  std::vector<double>& h = ws.host_keep.back();
  for (int64_t i = 0; i < n; ++i) { h[(size_t)(c + i)] = weights[i]; h[(size_t)(c - i)] = weights[i]; }
  double* d = nullptr;
  if (ws.alloc(&d, h.size())) return 1;
  WB_CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ws.stream));
  *d_center = d + c;
  return 0;
}

static int begin_single_device(int device, DeviceInfo* di, cudaStream_t* st) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_err("no CUDA device available: wildboar_b200 has no CPU fallback"); return 1; }
  if (device < 0 || device >= ndev) { set_err("invalid device ordinal"); return 1; }
  WB_CK(cudaSetDevice(device));
  if (device_info(di)) return 1;
  WB_CK(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
  return 0;
}

static int upload_index(Workspace& ws, const int64_t* h, int64_t n, int64_t limit, int** d) {
  *d = nullptr;
  if (!h) return 0;
  ws.host_keep_i.emplace_back((size_t)n);
  std::vector<int>& v = ws.host_keep_i.back();
  for (int64_t k = 0; k < n; ++k) {
    if (h[k] < 0 || h[k] >= limit) { set_err("index out of range"); return 1; }
    v[(size_t)k] = (int)h[k];
  }
  if (ws.alloc(d, (size_t)n)) return 1;
  WB_CK(cudaMemcpyAsync(*d, v.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ws.stream));
  return 0;
}

static int run_dtw_paths_host(const double* a, int64_t na, int64_t Ta, int64_t as, const double* b, int64_t nb, int64_t Tb,
                              int64_t bs, const int64_t* ia, const int64_t* ib, int64_t n_pairs, double r,
                              const double* weights, int32_t* lo, int32_t* hi, double* cost, double* D, int device,
                              wb_stats* stats) {
  DeviceInfo di; cudaStream_t st;
  if (begin_single_device(device, &di, &st)) return 1;
  int rc = 0;
  wb_stats local; memset(&local, 0, sizeof local);
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    do {
      double *da = nullptr, *db = nullptr, *dcost = nullptr, *dD = nullptr;
      int *dia = nullptr, *dib = nullptr, *dlo = nullptr, *dhi = nullptr;
      const double* dw = nullptr;
      if ((rc = ws.alloc(&da, (size_t)na * Ta)) || (rc = ws.alloc(&db, (size_t)nb * Tb))) break;
      if ((rc = h2d_rows(da, a, na, Ta, as, st)) || (rc = h2d_rows(db, b, nb, Tb, bs, st))) break;
      if ((rc = upload_index(ws, ia, n_pairs, na, &dia)) || (rc = upload_index(ws, ib, n_pairs, nb, &dib))) break;
      if ((rc = upload_weight_table(ws, weights, std::max(Ta, Tb), &dw))) break;
      if ((rc = ws.alloc(&dlo, (size_t)n_pairs * Ta)) || (rc = ws.alloc(&dhi, (size_t)n_pairs * Ta))) break;
      if (cost && (rc = ws.alloc(&dcost, (size_t)n_pairs))) break;
      if (D && (rc = ws.alloc(&dD, (size_t)n_pairs * Ta * Tb))) break;
      kt.start();
      if ((rc = run_paths_dev(ws, di.max_smem_optin, di.sms, da, Ta, db, Tb, dia, dib, n_pairs, warp_size_max(Ta, Tb, r), dw, dlo, dhi, dcost, dD, &local))) break;
      kt.stop();
      if (cudaMemcpyAsync(lo, dlo, sizeof(int) * n_pairs * Ta, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(hi, dhi, sizeof(int) * n_pairs * Ta, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          (cost && cudaMemcpyAsync(cost, dcost, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, st) != cudaSuccess) ||
          (D && cudaMemcpyAsync(D, dD, sizeof(double) * n_pairs * Ta * Tb, cudaMemcpyDeviceToHost, st) != cudaSuccess) ||
          cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the warping paths failed"); rc = 1; break; }
      local.kernel_ms = kt.ms();
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); local.total_ms = total.ms(); }
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (stats) *stats = local;
  return rc;
}

// One DBA step for K clusters against a resident sample set (see include/wb_cuda.h, wb_cuda_dba_epoch).
static int run_dba_epoch(const wb_fitted* fit, int metric, const wb_params& prm, const double* means_in, int64_t K, int64_t Tm,
                         const int64_t* off, const int64_t* member, const double* sample_weight, const double* weights,
                         int do_update, double* means_out, double* dist_out, wb_stats* stats) {
  DeviceInfo di; cudaStream_t st;
  if (begin_single_device(fit->devs[0], &di, &st)) return 1;
  int rc = 0;
  wb_stats local; memset(&local, 0, sizeof local);
  const int64_t n_m = off[K], T = fit->T;
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    do {
      const double* dX = fit->ptr[0];
      double *dmeans = nullptr, *dnew = nullptr, *ddist = nullptr, *dsw = nullptr;
      int *dmem = nullptr, *dlo = nullptr, *dhi = nullptr, *dlen = nullptr, *dia = nullptr;
      long long* doff = nullptr; int2* dlist = nullptr;
      const double* dw = nullptr;
      if ((rc = ws.alloc(&dmeans, (size_t)K * Tm)) || (rc = ws.alloc(&dnew, (size_t)K * Tm)) || (rc = ws.alloc(&ddist, (size_t)n_m)) ||
          (rc = ws.alloc(&doff, (size_t)K + 1)) || (rc = ws.alloc(&dlist, (size_t)n_m)) || (rc = ws.alloc(&dlen, 1))) break;
      WB_CK(cudaMemcpyAsync(dmeans, means_in, sizeof(double) * K * Tm, cudaMemcpyHostToDevice, st));
      if ((rc = upload_index(ws, member, n_m, fit->n, &dmem))) break;
      // per member: its cluster (row of the means) -- the first operand of every alignment / distance
      ws.host_keep_i.emplace_back((size_t)n_m);
      std::vector<int>& hia = ws.host_keep_i.back();
      ws.host_keep_i.emplace_back((size_t)2 * n_m + 1);
      std::vector<int>& hlist = ws.host_keep_i.back();
      for (int64_t c = 0; c < K; ++c) {
        if (off[c + 1] < off[c]) { set_err("member offsets must be non-decreasing"); rc = 1; break; }
        for (int64_t q = off[c]; q < off[c + 1]; ++q) { hia[(size_t)q] = (int)c; hlist[(size_t)2 * q] = (int)c; hlist[(size_t)2 * q + 1] = (int)member[q]; }
      }
      if (rc) break;
      hlist[(size_t)2 * n_m] = (int)n_m;
      if ((rc = ws.alloc(&dia, (size_t)n_m))) break;
      WB_CK(cudaMemcpyAsync(dia, hia.data(), sizeof(int) * n_m, cudaMemcpyHostToDevice, st));
      WB_CK(cudaMemcpyAsync(dlist, hlist.data(), sizeof(int) * 2 * n_m, cudaMemcpyHostToDevice, st));
      WB_CK(cudaMemcpyAsync(dlen, hlist.data() + 2 * n_m, sizeof(int), cudaMemcpyHostToDevice, st));
      WB_CK(cudaMemcpyAsync(doff, off, sizeof(long long) * (K + 1), cudaMemcpyHostToDevice, st));
      if (sample_weight) {
        if ((rc = ws.alloc(&dsw, (size_t)fit->n))) break;
        WB_CK(cudaMemcpyAsync(dsw, sample_weight, sizeof(double) * fit->n, cudaMemcpyHostToDevice, st));
      }
      kt.start();
      const double* dcur = dmeans;
      if (do_update) {
        if ((rc = upload_weight_table(ws, weights, std::max(Tm, T), &dw))) break;
        if ((rc = ws.alloc(&dlo, (size_t)n_m * Tm)) || (rc = ws.alloc(&dhi, (size_t)n_m * Tm))) break;
        if ((rc = run_paths_dev(ws, di.max_smem_optin, di.sms, dmeans, Tm, dX, T, dia, dmem, n_m, warp_size_max(Tm, T, prm.r), dw, dlo, dhi, nullptr, nullptr, &local))) break;
        DbaArgs u;
        u.X = dX; u.T = (int)T; u.member = dmem; u.off = doff; u.sw = dsw; u.lo = dlo; u.hi = dhi; u.K = (int)K; u.Tm = (int)Tm;
        u.mean_out = dnew;
        k_dba_update<<<(unsigned)((K * Tm + 31) / 32), 32, 0, st>>>(u);
        WB_CK(cudaGetLastError());
        local.launches += 1;
        dcur = dnew;
      }
      // distance of every member to its (new) centre: metric(centre, sample), as `pairwise_distance(mean, X)` (dtw.py:590-600)
      DpCall c; memset(&c, 0, sizeof c);
      c.metric = metric; c.p = prm; c.x = dcur; c.nx = K; c.Tx = (int)Tm; c.y = dX; c.ny = fit->n; c.Ty = (int)T;
      c.mode = PM_LISTP; c.list = dlist; c.list_len = dlen; c.list_n = n_m;
      if ((rc = prepare_operands(ws, c))) break;
      if ((rc = launch_dp(ws, di, c, 0, K, 0, fit->n, ddist, fit->n, nullptr, nullptr, &local))) break;
      kt.stop();
      if (cudaMemcpyAsync(means_out, dcur, sizeof(double) * K * Tm, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(dist_out, ddist, sizeof(double) * n_m, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the DBA step failed"); rc = 1; break; }
      local.kernel_ms = kt.ms();
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); local.total_ms = total.ms(); }
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (stats) *stats = local;
  return rc;
}

// ------------------------------------------------------------------------------------------
// Subsequence search, DTW family: min over the sliding windows of every sample.
// ------------------------------------------------------------------------------------------
struct SubseqJob {
  int metric; wb_params p;
  const double* s; const int64_t* soff; int64_t ns;   // subsequences, concatenated
  const double* x; int64_t nx, T, xs;
  int paired;
  int scaled;                                          // 1: scaled_<metric>: s is already z-normalised (dtw: UCR suite)
  const double* s_eps;                                 // edr, unscaled: epsilon per subsequence (or nullptr)
  double* out_dist; int64_t* out_idx;                  // (nx, ns) or, paired, (nx)
};

// EL:1917-1921 `_compute_warp_width` (the UCR-suite band |i - j| <= width; 0 = diagonal only)
static int64_t compute_warp_width(int64_t length, double r) {
  return r == 1.0 ? length - 1 : (int64_t)std::floor((double)length * r);
}

// samples [lo, hi) on device `dev`
static int subseq_worker(const SubseqJob& J, int dev, int64_t lo, int64_t hi, wb_stats* st_out) {
  DeviceInfo di; cudaStream_t st;
  if (begin_single_device(dev, &di, &st)) return 1;
  int rc = 0;
  wb_stats stats; memset(&stats, 0, sizeof stats);
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    const int64_t rows = hi - lo;
    const bool deriv = is_derivative(J.metric);
    const bool weighted = J.metric == M_WDTW || J.metric == M_WDDTW;
    const int64_t Tp = deriv ? J.T - 2 : J.T;  // length of the (prepared) samples
    do {
      double *dx = nullptr, *dxp = nullptr, *ds = nullptr, *draw = nullptr, *ddist = nullptr;
      long long* didx = nullptr;
      if ((rc = ws.alloc(&dx, (size_t)rows * J.T)) || (rc = h2d_rows(dx, J.x + lo * J.xs, rows, J.T, J.xs, st))) break;
      dxp = dx;
      if (deriv && Tp >= 1) {
        // the derivative of a window is the window of the derivative (average_slope is local, EL:3220-3225)
        if ((rc = ws.alloc(&dxp, (size_t)rows * Tp))) break;
        k_slope<<<1024, 256, 0, st>>>(dx, rows, (int)J.T, dxp);
        WB_CK(cudaGetLastError());
      }
      // subsequences (derivative metrics: their slopes), concatenated
      const int64_t stot = J.soff[J.ns];
      ws.host_keep.emplace_back((size_t)std::max<int64_t>(stot, 1));
      std::vector<double>& hs = ws.host_keep.back();
      std::vector<int64_t> poff((size_t)J.ns + 1, 0);  // offsets of the prepared subsequences
      for (int64_t k = 0; k < J.ns; ++k) {
        const int64_t m = J.soff[k + 1] - J.soff[k];
        const int64_t mp = deriv ? std::max<int64_t>(m - 2, 0) : m;
        if (deriv) { if (m >= 3) average_slope(J.s + J.soff[k], m, hs.data() + poff[(size_t)k]); }
        else memcpy(hs.data() + poff[(size_t)k], J.s + J.soff[k], sizeof(double) * m);
        poff[(size_t)k + 1] = poff[(size_t)k] + mp;
      }
      if ((rc = ws.alloc(&ds, hs.size()))) break;
      WB_CK(cudaMemcpyAsync(ds, hs.data(), sizeof(double) * hs.size(), cudaMemcpyHostToDevice, st));
      // weights: one table over the SERIES length (EL:2370-2372 wdtw: T; EL:2540-2543 wddtw: T - 2), libm exp
      const double* dw = nullptr;
      if (weighted) {
        const int64_t wn = deriv ? J.T - 2 : J.T;
        ws.host_keep.push_back(make_weights(J.p.g, wn));
        std::vector<double>& h = ws.host_keep.back();
        double* d = nullptr;
        if ((rc = ws.alloc(&d, h.size()))) break;
        WB_CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        dw = d + table_center(wn);
      }
      // scaled_dtw: running mean / std of every window, recomputed when the subsequence length changes
      double *dmean = nullptr, *dstd = nullptr, *dkim = nullptr, *tau = nullptr, *hval = nullptr;
      long long* hidx = nullptr; int* hn = nullptr;
      int64_t stats_m = -1;
      if (J.scaled && ((rc = ws.alloc(&dmean, (size_t)rows * Tp)) || (rc = ws.alloc(&dstd, (size_t)rows * Tp)) ||
                       (rc = ws.alloc(&dkim, (size_t)rows * Tp)) || (rc = ws.alloc(&tau, (size_t)rows)) ||
                       (rc = ws.alloc(&hval, (size_t)rows)) || (rc = ws.alloc(&hidx, (size_t)rows)) || (rc = ws.alloc(&hn, (size_t)rows)))) break;
      const int64_t nout = J.paired ? rows : rows * J.ns;
      if ((rc = ws.alloc(&draw, (size_t)rows * std::max<int64_t>(Tp, 1))) || (rc = ws.alloc(&ddist, (size_t)nout)) ||
          (rc = ws.alloc(&didx, (size_t)nout))) break;
      WB_CK(cudaMemsetAsync(ddist, 0, sizeof(double) * nout, st));
      WB_CK(cudaMemsetAsync(didx, 0, sizeof(long long) * nout, st));
      // head-then-abandon (unscaled scans): pair list / distances of the first 32 windows of every sample, their minimum
      const bool head_ok = !J.scaled && !getenv("WILDBOAR_CUDA_SCAN_NO_ABANDON");
      int2* hlist = nullptr; int* hdlen = nullptr; double *hd1 = nullptr, *hthr = nullptr; long long* hidx_tmp = nullptr;
      if (head_ok && ((rc = ws.alloc(&hlist, (size_t)rows * 32)) || (rc = ws.alloc(&hdlen, 1)) || (rc = ws.alloc(&hd1, (size_t)rows * 32)) ||
                      (rc = ws.alloc(&hthr, (size_t)rows)) || (rc = ws.alloc(&hidx_tmp, (size_t)rows)))) break;
      kt.start();
      // the DP policy on prepared data: ddtw -> dtw, wddtw -> wdtw
      const int dp_metric = J.scaled ? (int)M_SCALED_DTW : (J.metric == M_DDTW ? M_DTW : (J.metric == M_WDDTW ? M_WDTW : J.metric));
      // Unpaired, unscaled: subsequences of one length are the x rows of ONE launch (they share every block of 32 windows, the
      // head thresholds are per (subsequence, sample)); the per-subsequence loop below serves paired calls and scaled_dtw.
      const bool grouped = !J.scaled && !J.paired;
      if (grouped) {
        std::map<int64_t, std::vector<int64_t>> groups;
        for (int64_t k = 0; k < J.ns; ++k) {
          const int64_t m = J.soff[k + 1] - J.soff[k];
          if (deriv && m < 3) continue;  // EL:793-794: distance 0 (index unspecified in the reference; 0 here)
          groups[m].push_back(k);
        }
        for (auto& kv : groups) {
          if (rc) break;
          const int64_t m = kv.first;
          const std::vector<int64_t>& ks = kv.second;
          const int64_t G = (int64_t)ks.size(), mp = deriv ? m - 2 : m, nw = Tp - mp + 1, nr = rows;
          const long long ldd = nr * Tp;
          const int64_t gstep = std::max<int64_t>(1, std::min<int64_t>(G, ((int64_t)1 << 26) / std::max<long long>(ldd, 1)));
          for (int64_t g0 = 0; g0 < G && !rc; g0 += gstep) {
            const int64_t gc = std::min(gstep, G - g0);
            Workspace it(st);
            double *dsg = nullptr, *dgraw = nullptr; int* dks = nullptr;
            if ((rc = it.alloc(&dsg, (size_t)(gc * mp))) || (rc = it.alloc(&dks, (size_t)gc)) || (rc = it.alloc(&dgraw, (size_t)(gc * ldd)))) break;
            ws.host_keep.emplace_back((size_t)(gc * mp));
            std::vector<double>& hg = ws.host_keep.back();
            ws.host_keep_i.emplace_back((size_t)gc);
            std::vector<int>& hk = ws.host_keep_i.back();
            for (int64_t g = 0; g < gc; ++g) {
              const int64_t k = ks[(size_t)(g0 + g)];
              memcpy(hg.data() + g * mp, hs.data() + poff[(size_t)k], sizeof(double) * mp);
              hk[(size_t)g] = (int)k;
            }
            WB_CK(cudaMemcpyAsync(dsg, hg.data(), sizeof(double) * gc * mp, cudaMemcpyHostToDevice, st));
            WB_CK(cudaMemcpyAsync(dks, hk.data(), sizeof(int) * gc, cudaMemcpyHostToDevice, st));
            DpCall c; memset(&c, 0, sizeof c);
            c.metric = dp_metric; c.p = J.p; c.mode = PM_PAIRWISE;
            c.px = dsg; c.nx = gc; c.ptx = (int)mp;
            c.py = dxp; c.pty = (int)mp; c.ys = 1; c.ny = nr * Tp - mp + 1;
            c.R = (int)compute_r(m, J.p.r);  // from the ORIGINAL subsequence length (EL:2253, 2480)
            c.raw = 1;
            c.tab.weights = dw;
            const double* thr = nullptr;
            if (head_ok && nw >= 128) {
              // head-then-abandon, see the per-subsequence loop below
              const long long nq = gc * nr, n1 = nq * 32;
              int2* list = nullptr; int* dlen = nullptr; double *d1 = nullptr, *thr_w = nullptr; long long* tmp = nullptr;
              if ((rc = it.alloc(&list, (size_t)n1)) || (rc = it.alloc(&dlen, 1)) || (rc = it.alloc(&d1, (size_t)n1)) ||
                  (rc = it.alloc(&thr_w, (size_t)nq)) || (rc = it.alloc(&tmp, (size_t)nq))) break;
              const int n32 = (int)n1;
              WB_CK(cudaMemcpyAsync(dlen, &n32, sizeof(int), cudaMemcpyHostToDevice, st));
              k_scan_head_list<<<148 * 4, 256, 0, st>>>(list, n1, nr, 32, Tp);
              WB_CK(cudaGetLastError());
              DpCall ch = c;
              ch.mode = PM_LISTP; ch.list = list; ch.list_len = dlen; ch.list_n = n1;
              if ((rc = launch_dp(it, di, ch, 0, gc, 0, c.ny, d1, 0, nullptr, nullptr, &stats))) break;
              k_window_min<<<(unsigned)((nq * 32 + 127) / 128), 128, 0, st>>>(d1, nq, 32, 32, thr_w, tmp, 1, 0);
              WB_CK(cudaGetLastError());
              stats.launches += 2;
              thr = thr_w; c.thr_ld = nr; c.thr_div = Tp;
            }
            if ((rc = launch_dp(it, di, c, 0, gc, 0, c.ny, dgraw, ldd, nullptr, thr, &stats))) break;
            k_window_min<<<(unsigned)((gc * nr * 32 + 127) / 128), 128, 0, st>>>(dgraw, gc * nr, (int)Tp, (int)nw, ddist, didx, J.ns, 1, dks, nr);
            WB_CK(cudaGetLastError());
            stats.launches += 1;
          }
        }
      }
      // Unpaired scaled_dtw: the same grouping -- one DP launch per length group, LB_Kim per subsequence into a buffer laid out
      // like the distances, ONE replay over all (subsequence, sample) queries.
      const bool grouped_ucr = J.scaled && !J.paired;
      if (grouped_ucr) {
        std::map<int64_t, std::vector<int64_t>> groups;
        for (int64_t k = 0; k < J.ns; ++k) groups[J.soff[k + 1] - J.soff[k]].push_back(k);
        for (auto& kv : groups) {
          if (rc) break;
          const int64_t m = kv.first;
          const std::vector<int64_t>& ks = kv.second;
          const int64_t G = (int64_t)ks.size(), nw = Tp - m + 1, nr = rows;
          const long long ldd = nr * Tp;
          k_window_stats<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(dxp, rows, (int)Tp, (int)m, dmean, dstd);
          WB_CK(cudaGetLastError());
          stats.launches += 1;
          const int64_t gstep = std::max<int64_t>(1, std::min<int64_t>(G, ((int64_t)1 << 25) / std::max<long long>(ldd, 1)));
          for (int64_t g0 = 0; g0 < G && !rc; g0 += gstep) {
            const int64_t gc = std::min(gstep, G - g0);
            const long long nq = gc * nr;
            Workspace it(st);
            double *dsg = nullptr, *dgraw = nullptr, *gkim = nullptr, *gtau = nullptr, *ghval = nullptr;
            long long* ghidx = nullptr; int *ghn = nullptr, *dks = nullptr;
            if ((rc = it.alloc(&dsg, (size_t)(gc * m))) || (rc = it.alloc(&dks, (size_t)gc)) || (rc = it.alloc(&dgraw, (size_t)(gc * ldd))) ||
                (rc = it.alloc(&gkim, (size_t)(gc * ldd))) || (rc = it.alloc(&gtau, (size_t)nq)) || (rc = it.alloc(&ghval, (size_t)nq)) ||
                (rc = it.alloc(&ghidx, (size_t)nq)) || (rc = it.alloc(&ghn, (size_t)nq))) break;
            ws.host_keep.emplace_back((size_t)(gc * m));
            std::vector<double>& hg = ws.host_keep.back();
            ws.host_keep_i.emplace_back((size_t)gc);
            std::vector<int>& hk = ws.host_keep_i.back();
            for (int64_t g = 0; g < gc; ++g) {
              const int64_t k = ks[(size_t)(g0 + g)];
              memcpy(hg.data() + g * m, hs.data() + poff[(size_t)k], sizeof(double) * m);
              hk[(size_t)g] = (int)k;
            }
            WB_CK(cudaMemcpyAsync(dsg, hg.data(), sizeof(double) * gc * m, cudaMemcpyHostToDevice, st));
            WB_CK(cudaMemcpyAsync(dks, hk.data(), sizeof(int) * gc, cudaMemcpyHostToDevice, st));
            DpCall c; memset(&c, 0, sizeof c);
            c.metric = dp_metric; c.p = J.p; c.mode = PM_PAIRWISE;
            c.px = dsg; c.nx = gc; c.ptx = (int)m;
            c.py = dxp; c.pty = (int)m; c.ys = 1; c.ny = nr * Tp - m + 1;
            c.R = (int)compute_warp_width(m, J.p.r) + 1;  // band |i - j| <= warp width (EL:2033, 283-292)
            c.sy = dmean; c.sy2 = dstd;
            c.raw = 1;
            if ((rc = launch_dp(it, di, c, 0, gc, 0, c.ny, dgraw, ldd, nullptr, nullptr, &stats))) break;
            const long long npair = nr * nw;
            for (int64_t g = 0; g < gc; ++g)
              k_ucr_kim<<<(unsigned)((npair + 255) / 256), 256, 0, st>>>(dxp, nr, (int)Tp, (int)m, dsg + g * m, dmean, dstd, gkim + g * ldd);
            k_fill<<<64, 256, 0, st>>>(gtau, nq, WB_INF);
            k_fill<<<64, 256, 0, st>>>(ghval, nq, WB_INF);
            WB_CK(cudaMemsetAsync(ghidx, 0, sizeof(long long) * nq, st));
            WB_CK(cudaMemsetAsync(ghn, 0, sizeof(int) * nq, st));
            ReplayArgs ra;
            ra.d = dgraw; ra.m = nullptr; ra.lb = gkim; ra.ld = Tp; ra.nq = nq; ra.c0 = 0; ra.ncols = nw;
            ra.k = 1; ra.kind = TK_NONE; ra.scale = 1.0; ra.tau = gtau; ra.hidx = ghidx; ra.hval = ghval; ra.hn = ghn;
            k_replay<<<(unsigned)std::max<long long>(1, std::min<long long>((nq + 3) / 4, 148 * 16)), 128, 0, st>>>(ra);
            k_finish_scan_group<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(ghval, ghidx, nr, gc, dks, ddist, didx, J.ns, 1);
            WB_CK(cudaGetLastError());
            stats.launches += (int)gc + 4;
          }
        }
      }
      for (int64_t k = 0; k < J.ns && !rc && !grouped && !grouped_ucr; ++k) {
        const int64_t m = J.soff[k + 1] - J.soff[k];
        const int64_t mp = poff[(size_t)k + 1] - poff[(size_t)k];
        if (deriv && m < 3) continue;  // EL:793-794: distance 0 (index unspecified in the reference; 0 here)
        const int64_t nw = Tp - mp + 1;  // windows per sample
        // paired: subsequence k against sample k only; else against every sample of the block
        const int64_t r0 = J.paired ? k - lo : 0, nr = J.paired ? 1 : rows;
        if (J.paired && (k < lo || k >= hi)) continue;
        if (J.scaled && m != stats_m) {
          k_window_stats<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(dxp, rows, (int)Tp, (int)m, dmean, dstd);
          WB_CK(cudaGetLastError());
          stats.launches += 1;
          stats_m = m;
        }
        DpCall c; memset(&c, 0, sizeof c);
        c.metric = dp_metric; c.p = J.p; c.mode = PM_PAIRWISE;
        c.px = ds + poff[(size_t)k]; c.nx = 1; c.ptx = (int)mp;
        c.py = dxp + r0 * Tp; c.pty = (int)mp; c.ys = 1; c.ny = nr * Tp - mp + 1;
        c.R = (int)compute_r(m, J.p.r);  // from the ORIGINAL subsequence length (EL:2253, 2480)
        if (J.scaled) {
          c.R = (int)compute_warp_width(m, J.p.r) + 1;  // band |i - j| <= warp width (EL:2033, 283-292)
          c.sy = dmean + r0 * Tp; c.sy2 = dstd + r0 * Tp;
        }
        c.raw = 1;
        c.tab.weights = dw;
        // Early abandoning as in the reference's scan (EL:622-660 hands the running minimum to dtw_distance): the first 32
        // windows of every sample are evaluated first (pair list), their minimum t1 can only be undercut, so the main launch
        // abandons a window once a column minimum of its band exceeds t1 (strip engine, EA) -- such a window cannot be the
        // (first) minimum; ties with t1 are not abandoned (strict >), so the first-minimum rule still sees them.
        const double* thr = nullptr;
        if (head_ok && nw >= 128) {
          const long long n1 = nr * 32;
          const int n32 = (int)n1;
          WB_CK(cudaMemcpyAsync(hdlen, &n32, sizeof(int), cudaMemcpyHostToDevice, st));
          k_scan_head_list<<<148 * 4, 256, 0, st>>>(hlist, n1, nr, 32, Tp);
          WB_CK(cudaGetLastError());
          DpCall ch = c;
          ch.mode = PM_LISTP; ch.list = hlist; ch.list_len = hdlen; ch.list_n = n1;
          if ((rc = launch_dp(ws, di, ch, 0, 1, 0, c.ny, hd1, 0, nullptr, nullptr, &stats))) break;
          k_window_min<<<(unsigned)((nr * 32 + 127) / 128), 128, 0, st>>>(hd1, nr, 32, 32, hthr, hidx_tmp, 1, 0);
          WB_CK(cudaGetLastError());
          stats.launches += 2;
          thr = hthr; c.thr_ld = 0; c.thr_div = Tp;
        }
        if ((rc = launch_dp(ws, di, c, 0, 1, 0, c.ny, draw, c.ny, nullptr, thr, &stats))) break;
        double* od = J.paired ? ddist + r0 : ddist + k;
        long long* oi = J.paired ? didx + r0 : didx + k;
        if (J.scaled) {
          // the reference's scan, replayed exactly: windows in order, running minimum t, a window is skipped when its
          // LB_Kim value is >= t (EL:413) and accepted iff its distance is < t (EL:471) -- all in the squared-cost domain
          const long long npair = nr * nw;
          k_ucr_kim<<<(unsigned)((npair + 255) / 256), 256, 0, st>>>(dxp + r0 * Tp, nr, (int)Tp, (int)m, ds + poff[(size_t)k],
                                                                      dmean + r0 * Tp, dstd + r0 * Tp, dkim);
          k_fill<<<64, 256, 0, st>>>(tau, nr, WB_INF);
          WB_CK(cudaMemsetAsync(hn, 0, sizeof(int) * nr, st));
          ReplayArgs ra;
          ra.d = draw; ra.m = nullptr; ra.lb = dkim; ra.ld = Tp; ra.nq = nr; ra.c0 = 0; ra.ncols = nw;
          ra.k = 1; ra.kind = TK_NONE; ra.scale = 1.0; ra.tau = tau; ra.hidx = hidx; ra.hval = hval; ra.hn = hn;
          k_replay<<<(unsigned)std::max<long long>(1, std::min<long long>((nr + 3) / 4, 148 * 16)), 128, 0, st>>>(ra);
          k_finish_scan<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(hval, hidx, nr, od, oi, J.paired ? 1 : J.ns, 1);
          WB_CK(cudaGetLastError());
          stats.launches += 4;
          continue;
        }
        k_window_min<<<(unsigned)((nr * 32 + 127) / 128), 128, 0, st>>>(draw, nr, (int)Tp, (int)nw, od, oi, J.paired ? 1 : J.ns, 1);
        WB_CK(cudaGetLastError());
        stats.launches += 1;
      }
      if (rc) break;
      kt.stop();
      // gather: pairwise rows are contiguous (rows, ns); paired entries are contiguous (rows)
      double* hd = J.paired ? J.out_dist + lo : J.out_dist + lo * J.ns;
      int64_t* hi_ = J.paired ? J.out_idx + lo : J.out_idx + lo * J.ns;
      if (cudaMemcpyAsync(hd, ddist, sizeof(double) * nout, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(hi_, didx, sizeof(long long) * nout, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the subsequence distances failed"); rc = 1; break; }
      stats.kernel_ms = kt.ms();
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); stats.total_ms = total.ms(); }
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (st_out) *st_out = stats;
  return rc;
}

// ------------------------------------------------------------------------------------------
// Subsequence search as an exact replay of the reference's scan (second half):
//   * lcss / erp / edr / msm / twe SubsequenceMetric (EL:2616-3124; *_subsequence_distance EL:1186, 1350, 1500, 1650, 1832):
//     every window's *_distance() is early-abandoned against the running minimum, and for these metrics the abandoning
//     decides which windows are accepted (row minima are not monotone);
//   * scaled_<metric> = ScaledSubsequenceMetricWrap(Metric) (CD:470-551) for adtw, wdtw, ddtw, wddtw, lcss, erp, edr, msm,
//     twe: windows z-normalised with the running IncStats, then Metric._eadistance() against the running minimum.
// Device scheme (the one argmin uses, argmin.cuh): ONE DP launch per subsequence evaluates all windows of all samples of
// the block without abandoning and records, per window, the distance d and M = max over the checked rows of the row
// minimum; k_replay (one warp per sample) then walks the windows in order with the exact rule "accept iff d < t and
// not M > T(t)" (T = the metric's threshold transform).  Unscaled windows are addressed in place (stride 1); scaled
// windows are materialised z-normalised as dense rows, a bounded number of samples at a time.
// ------------------------------------------------------------------------------------------
static bool subseq_uses_scan(const SubseqJob& J) {
  if (J.scaled) return J.metric != M_DTW;
  return !is_dtw_family(J.metric) || (J.metric == M_ADTW && J.p.p < 0);
}

static int subseq_scan_worker(const SubseqJob& J, int dev, int64_t lo, int64_t hi, wb_stats* st_out) {
  DeviceInfo di; cudaStream_t st;
  if (begin_single_device(dev, &di, &st)) return 1;
  int rc = 0;
  wb_stats stats; memset(&stats, 0, sizeof stats);
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    const int64_t rows = hi - lo, T = J.T;
    const bool dtwfam = is_dtw_family(J.metric);
    const bool deriv = is_derivative(J.metric);
    const bool want_m = !dtwfam || (J.metric == M_ADTW && J.p.p < 0);
    do {
      double *dx = nullptr, *ddist = nullptr;
      long long* didx = nullptr;
      if ((rc = ws.alloc(&dx, (size_t)rows * T)) || (rc = h2d_rows(dx, J.x + lo * J.xs, rows, T, J.xs, st))) break;
      // tables over the SERIES length: wdtw / wddtw weights (wrap.reset(X, X): EL:3334-3341 T, EL:3415-3428 T - 2, libm exp);
      // twe's 2 * stiffness * |i - j| does not depend on the length
      const double *dw = nullptr, *dtw = nullptr;
      if (J.metric == M_WDTW || J.metric == M_WDDTW || J.metric == M_TWE) {
        const int64_t tn = J.metric == M_TWE ? T + 1 : (deriv ? T - 2 : T);
        ws.host_keep.push_back(J.metric == M_TWE ? make_tw(J.p.stiffness, tn) : make_weights(J.p.g, tn));
        std::vector<double>& h = ws.host_keep.back();
        double* d = nullptr;
        if ((rc = ws.alloc(&d, h.size()))) break;
        WB_CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        (J.metric == M_TWE ? dtw : dw) = d + table_center(tn);
      }
      const int64_t nout = J.paired ? rows : rows * J.ns;
      if ((rc = ws.alloc(&ddist, (size_t)nout)) || (rc = ws.alloc(&didx, (size_t)nout))) break;
      WB_CK(cudaMemsetAsync(ddist, 0, sizeof(double) * nout, st));
      WB_CK(cudaMemsetAsync(didx, 0, sizeof(long long) * nout, st));
      // Work units: (subsequences of ONE length, sample range).  Unpaired: every length group against all samples of the block
      // -- the group is the x operand of one DP launch (its rows share every block of 32 windows, the window statistics and
      // the z-normalised windows are computed once per length, not once per subsequence); paired: subsequence k against
      // sample k only.
      struct Unit { std::vector<int64_t> ks; int64_t b0, bn; };
      std::vector<Unit> units;
      if (J.paired) {
        for (int64_t k = lo; k < hi; ++k) units.push_back(Unit{{k}, k - lo, 1});
      } else {
        std::map<int64_t, size_t> by_len;
        for (int64_t k = 0; k < J.ns; ++k) {
          const int64_t m = J.soff[k + 1] - J.soff[k];
          auto itl = by_len.find(m);
          if (itl == by_len.end()) { by_len[m] = units.size(); units.push_back(Unit{{k}, 0, rows}); }
          else units[itl->second].ks.push_back(k);
        }
      }
      kt.start();
      for (size_t u = 0; u < units.size() && !rc; ++u) {
        const Unit& U = units[u];
        const int64_t m = J.soff[U.ks[0] + 1] - J.soff[U.ks[0]];
        const int64_t nw = T - m + 1;  // windows per sample
        const int64_t G = (int64_t)U.ks.size();
        // replay rule of the metric: T(t) handed to the DP by *_subsequence_distance (unscaled) / _eadistance (scaled)
        int kind = TK_IDENT; double scale = 1.0;
        if (dtwfam) kind = J.scaled ? TK_SQUARE : TK_IDENT;  // unscaled adtw scans in the squared-cost domain (EL:701-740)
        else if (J.metric == M_LCSS) { kind = TK_LCSS; scale = (double)m; }                      // EL:1213-1215, 3526-3528
        else if (J.metric == M_EDR) { kind = TK_SCALE; scale = (double)(J.scaled ? m : T); }     // EL:3875 / EL:1523 (series length)
        const long long ld = J.scaled ? nw : T;  // distance entries per sample (unscaled: incl. the windows that straddle two samples)
        // samples per pass: the materialised windows of the scaled metrics stay below ~256 MB; subsequences per launch: at
        // most 2^26 (sample, window, subsequence) entries
        int64_t step = U.bn;
        if (J.scaled) {
          int64_t budget = (int64_t)32 << 20;  // doubles
          if (const char* e = getenv("WILDBOAR_CUDA_SCAN_WINDOW_BUDGET")) { const long long v = atoll(e); if (v > 0) budget = v; }  // test knob
          step = std::max<int64_t>(1, std::min<int64_t>(U.bn, budget / std::max<int64_t>(nw * m, 1)));
        }
        // the group's subsequences as dense rows (G, m)
        Workspace uw(st);
        double* dsg = nullptr;
        if ((rc = uw.alloc(&dsg, (size_t)(G * m)))) break;
        {
          ws.host_keep.emplace_back((size_t)(G * m));
          std::vector<double>& h = ws.host_keep.back();
          for (int64_t g = 0; g < G; ++g) memcpy(h.data() + g * m, J.s + J.soff[U.ks[(size_t)g]], sizeof(double) * m);
          WB_CK(cudaMemcpyAsync(dsg, h.data(), sizeof(double) * G * m, cudaMemcpyHostToDevice, st));
        }
        // unscaled edr with per-subsequence epsilons: the policy takes max(sx, sy) / 4 with sx = 4 * epsilon, sy = 0
        double* edr_sx = nullptr;
        const bool edr_eps = J.metric == M_EDR && !J.scaled && J.s_eps && std::isnan(J.p.epsilon);
        if (edr_eps) {
          ws.host_keep.emplace_back((size_t)G);
          std::vector<double>& h = ws.host_keep.back();
          for (int64_t g = 0; g < G; ++g) h[(size_t)g] = J.s_eps[U.ks[(size_t)g]] * 4.0;
          if ((rc = uw.alloc(&edr_sx, (size_t)G))) break;
          WB_CK(cudaMemcpyAsync(edr_sx, h.data(), sizeof(double) * G, cudaMemcpyHostToDevice, st));
        }
        // where query (g, i) of a launch goes: out[(r0 + i) * ns + ks[g]] (paired: out[r0 + i])
        int* dks = nullptr;
        {
          ws.host_keep_i.emplace_back((size_t)G);
          std::vector<int>& h = ws.host_keep_i.back();
          for (int64_t g = 0; g < G; ++g) h[(size_t)g] = J.paired ? 0 : (int)U.ks[(size_t)g];
          if ((rc = uw.alloc(&dks, (size_t)G))) break;
          WB_CK(cudaMemcpyAsync(dks, h.data(), sizeof(int) * G, cudaMemcpyHostToDevice, st));
        }
        const long long ldo = J.paired ? 1 : J.ns;
        for (int64_t q0 = 0; q0 < U.bn && !rc; q0 += step) {
          const int64_t r0 = U.b0 + q0, nr = std::min(step, U.bn - q0);
          const int64_t gstep = std::max<int64_t>(1, std::min<int64_t>(G, ((int64_t)1 << 26) / std::max<int64_t>(nr * ld, 1)));
          Workspace pw(st);  // buffers of this pass
          double *mean = nullptr, *stdv = nullptr, *wn = nullptr;
          if (J.scaled && !(deriv && m < 3)) {
            if ((rc = pw.alloc(&mean, (size_t)(nr * nw))) || (rc = pw.alloc(&stdv, (size_t)(nr * nw))) ||
                (rc = pw.alloc(&wn, (size_t)(nr * nw * m)))) break;
            k_inc_window_stats<<<(unsigned)((nr + 63) / 64), 64, 0, st>>>(dx + r0 * T, nr, (int)T, (int)m, mean, stdv);
            k_normalise_windows<<<148 * 8, 256, 0, st>>>(dx + r0 * T, nr, (int)T, (int)m, mean, stdv, wn);
            WB_CK(cudaGetLastError());
            stats.launches += 2;
          }
          for (int64_t g0 = 0; g0 < G && !rc; g0 += gstep) {
            const int64_t gc = std::min(gstep, G - g0);
            const long long nq = gc * nr;  // replay queries: (subsequence g, sample i) -> row g * nr + i of the distance buffer
            Workspace it(st);
            double *tau = nullptr, *hval = nullptr; long long* hidx = nullptr; int* hn = nullptr;
            if ((rc = it.alloc(&tau, (size_t)nq)) || (rc = it.alloc(&hval, (size_t)nq)) || (rc = it.alloc(&hidx, (size_t)nq)) ||
                (rc = it.alloc(&hn, (size_t)nq))) break;
            k_fill<<<64, 256, 0, st>>>(tau, nq, WB_INF);
            k_fill<<<64, 256, 0, st>>>(hval, nq, WB_INF);
            WB_CK(cudaMemsetAsync(hidx, 0, sizeof(long long) * nq, st));
            WB_CK(cudaMemsetAsync(hn, 0, sizeof(int) * nq, st));
            double* od = J.paired ? ddist + r0 : ddist + r0 * J.ns;
            long long* oi = J.paired ? didx + r0 : didx + r0 * J.ns;
            if (deriv && m < 3) {
              // EL:3297-3298: _eadistance() accepts nothing -> the minimum stays +inf (index left at 0)
              k_finish_scan_group<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(hval, hidx, nr, gc, dks + g0, od, oi, ldo, 0);
              WB_CK(cudaGetLastError());
              continue;
            }
            DpCall c; memset(&c, 0, sizeof c);
            c.metric = J.metric; c.p = J.p; c.mode = PM_PAIRWISE;
            if (edr_eps) c.p.epsilon = std::nan("");
            if (J.scaled) {
              c.x = dsg + g0 * m; c.nx = gc; c.Tx = (int)m;
              c.y = wn; c.ny = nr * nw; c.Ty = (int)m;
              c.ea = 1;  // _eadistance: ddtw band from the derivative length (EL:3308)
              if ((rc = prepare_operands(it, c))) break;
              if (dw) c.tab.weights = dw;
              if (want_m && (rc = interleave_y(it, c))) break;
            } else {
              c.px = dsg + g0 * m; c.nx = gc; c.ptx = (int)m;
              c.py = dx + r0 * T; c.pty = (int)m; c.ys = 1; c.ny = nr * T - m + 1;
              c.R = (int)compute_r(m, J.p.r);
              c.tab.tw = dtw;
              c.raw = dtwfam ? 1 : 0;
              if (J.metric == M_ERP) {
                double *sx = nullptr, *sy = nullptr;
                if ((rc = it.alloc(&sx, (size_t)gc)) || (rc = it.alloc(&sy, (size_t)c.ny))) break;
                k_series_stat<<<(unsigned)((gc + 127) / 128), 128, 0, st>>>(c.px, gc, (int)m, 0, J.p.g, sx, m);
                k_series_stat<<<(unsigned)((c.ny + 127) / 128), 128, 0, st>>>(c.py, c.ny, (int)m, 0, J.p.g, sy, 1);
                WB_CK(cudaGetLastError());
                stats.launches += 2;
                c.sx = sx; c.sy = sy;
              }
              if (edr_eps) c.sx = edr_sx + g0;
            }
            const long long ldd = nr * ld;  // distance entries per subsequence
            double *draw = nullptr, *mraw = nullptr;
            if ((rc = it.alloc(&draw, (size_t)(gc * ldd))) || (want_m && (rc = it.alloc(&mraw, (size_t)(gc * ldd))))) break;
            ReplayArgs ra;
            ra.lb = nullptr; ra.nq = nq; ra.k = 1; ra.kind = kind; ra.scale = scale;
            ra.tau = tau; ra.hidx = hidx; ra.hval = hval; ra.hn = hn;
            const unsigned rgrid = (unsigned)std::max<long long>(1, std::min<long long>((nq + 3) / 4, 148 * 16));
            // Early abandoning, as the reference's scan has it: the first `head` windows of every query are evaluated and
            // replayed first; their running minimum t1 bounds every later running minimum from above, so the main launch
            // may abandon a window as soon as a row minimum exceeds T(t1) -- the reference (bound T(t) <= T(t1)) abandons
            // that window too.  Needs T monotone in t (not lcss) and row minima (not the strip path); used where it pays:
            // the unscaled metrics with T(t) = t (msm, twe, erp: 2x fewer cells on random walks).  edr's bound t * T is far
            // above any row minimum of an m-row window, and z-normalised windows are too alike for the head to bound much
            // (measured: +6 % time), so those run the single launch.  WILDBOAR_CUDA_SCAN_ABANDON=0 / 1 forces it off / on.
            long long head = 0;
            const double* thr = nullptr;
            bool abandon = !J.scaled && kind == TK_IDENT;
            if (const char* e = getenv("WILDBOAR_CUDA_SCAN_ABANDON")) abandon = atoi(e) != 0;
            if (abandon && want_m && kind != TK_LCSS && kind != TK_NONE && nw >= 128) {
              head = 32;
              const long long n1 = nq * head;
              int2* list = nullptr; int* dlen = nullptr; double *d1 = nullptr, *m1 = nullptr, *thr_w = nullptr;
              if ((rc = it.alloc(&list, (size_t)n1)) || (rc = it.alloc(&dlen, 1)) || (rc = it.alloc(&d1, (size_t)n1)) ||
                  (rc = it.alloc(&m1, (size_t)n1)) || (rc = it.alloc(&thr_w, (size_t)nq))) break;
              const int n32 = (int)n1;
              WB_CK(cudaMemcpyAsync(dlen, &n32, sizeof(int), cudaMemcpyHostToDevice, st));
              k_scan_head_list<<<148 * 4, 256, 0, st>>>(list, n1, nr, (int)head, ld);
              WB_CK(cudaGetLastError());
              DpCall ch = c;
              ch.mode = PM_LISTP; ch.list = list; ch.list_len = dlen; ch.list_n = n1;
              if ((rc = launch_dp(it, di, ch, 0, gc, 0, c.ny, d1, 0, m1, nullptr, &stats))) break;
              ra.d = d1; ra.m = m1; ra.ld = head; ra.c0 = 0; ra.ncols = head;
              k_replay<<<rgrid, 128, 0, st>>>(ra);
              k_thr_raw<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(tau, nq, kind, scale, thr_w);
              WB_CK(cudaGetLastError());
              stats.launches += 3;
              thr = thr_w;
              c.thr_ld = nr; c.thr_div = ld;
            }
            if ((rc = launch_dp(it, di, c, 0, gc, 0, c.ny, draw, ldd, mraw, thr, &stats))) break;
            ra.d = draw + head; ra.m = mraw ? mraw + head : nullptr; ra.ld = ld; ra.c0 = head; ra.ncols = nw - head;
            k_replay<<<rgrid, 128, 0, st>>>(ra);
            k_finish_scan_group<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(hval, hidx, nr, gc, dks + g0, od, oi, ldo,
                                                                               (!J.scaled && dtwfam) ? 1 : 0);
            WB_CK(cudaGetLastError());
            stats.launches += 2;
            for (auto& v : it.host_keep) ws.host_keep.push_back(std::move(v));  // staging of async copies outlives the pass
          }
        }
      }
      if (rc) break;
      kt.stop();
      double* hd = J.paired ? J.out_dist + lo : J.out_dist + lo * J.ns;
      int64_t* hi_ = J.paired ? J.out_idx + lo : J.out_idx + lo * J.ns;
      if (cudaMemcpyAsync(hd, ddist, sizeof(double) * nout, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(hi_, didx, sizeof(long long) * nout, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the subsequence distances failed"); rc = 1; break; }
      stats.kernel_ms = kt.ms();
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); stats.total_ms = total.ms(); }
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (st_out) *st_out = stats;
  return rc;
}

// ------------------------------------------------------------------------------------------
// Subsequence matches / distance profile: the dense form of SubsequenceMetric._matches -- for ONE
// subsequence against every sample (subsequence_match) or subsequence i against sample i (paired_subsequence_match,
// distance_profile), out[i][w] = distance of window w where the reference reports it under `threshold`, NaN elsewhere.
// Every (sample, window) pair of a pass is one entry of a pair list (PM_LISTP): one DP launch, one selection kernel.
// ------------------------------------------------------------------------------------------
struct ProfileJob {
  int metric; wb_params p;
  const double* s; int64_t ns, m;      // ns == 1 or ns == nx (paired), dense (ns, m); scaled: z-normalised by the caller
  const double* x; int64_t nx, T, xs;
  int scaled; const double* s_eps;     // edr, unscaled, default epsilon: std / 4 per subsequence
  double threshold;
  double* out;                         // (nx, T - m + 1)
  // argmin_subsequence_distance (argmin_k > 0, paired): the k closest windows under the sequential scan with
  // Metric._eadistance (CD:1380-1548) instead of the dense profile; windows raw (scaled == 0) or z-normalised
  int64_t argmin_k; int64_t* out_idx; double* out_dist;  // (nx, k), heap order
  int64_t weight_len;  // > 0: wdtw / wddtw weight tables for a series of this many points instead of T (dilated profile)
};

static int subseq_profile_worker(const ProfileJob& J, int dev, int64_t lo, int64_t hi, wb_stats* st_out) {
  DeviceInfo di; cudaStream_t st;
  if (begin_single_device(dev, &di, &st)) return 1;
  int rc = 0;
  wb_stats stats; memset(&stats, 0, sizeof stats);
  {
    Workspace ws(st);
    Timer total(st), kt(st);
    total.start();
    const int64_t rows = hi - lo, T = J.T, m = J.m, nw = T - m + 1;
    const bool paired = J.ns > 1;
    const int64_t nsub = paired ? rows : 1;
    const bool dtwfam = is_dtw_family(J.metric);
    const bool deriv = is_derivative(J.metric);
    const int64_t K = J.argmin_k;
    const bool ucr = J.scaled && J.metric == M_DTW && K == 0;
    const bool wrap = (J.scaled && !ucr) || K > 0;   // Metric._eadistance on materialised windows
    const bool identity = K > 0 && !J.scaled;        // ... which are the raw windows (mean 0, std 1: (x - 0) / 1 == x)
    const bool want_m = !dtwfam || (J.metric == M_ADTW && J.p.p < 0);
    const double thr = J.threshold;
    do {
      double *dx = nullptr, *ds = nullptr, *dout = nullptr;
      if ((rc = ws.alloc(&dx, (size_t)rows * T)) || (rc = h2d_rows(dx, J.x + lo * J.xs, rows, T, J.xs, st))) break;
      if ((rc = ws.alloc(&ds, (size_t)(nsub * m))) || (K == 0 && (rc = ws.alloc(&dout, (size_t)(rows * nw))))) break;
      WB_CK(cudaMemcpyAsync(ds, J.s + (paired ? lo * m : 0), sizeof(double) * nsub * m, cudaMemcpyHostToDevice, st));
      double *tau = nullptr, *hval = nullptr; long long* hidx = nullptr; int* hn = nullptr;
      if (K > 0) {
        if ((rc = ws.alloc(&tau, (size_t)rows)) || (rc = ws.alloc(&hn, (size_t)rows)) || (rc = ws.alloc(&hval, (size_t)(rows * K))) ||
            (rc = ws.alloc(&hidx, (size_t)(rows * K)))) break;
        k_fill<<<64, 256, 0, st>>>(tau, rows, WB_INF);
        WB_CK(cudaMemsetAsync(hn, 0, sizeof(int) * rows, st));
        WB_CK(cudaMemsetAsync(hval, 0, sizeof(double) * rows * K, st));
        WB_CK(cudaMemsetAsync(hidx, 0, sizeof(long long) * rows * K, st));
      }
      kt.start();
      if (deriv && m < 3) {
        if (K > 0) { /* EL:3297: nothing is ever accepted; the reference returns uninitialised memory, we return zeros */ } else
        // EL:843-844 (ddtw_subsequence_matches returns no match), EL:3297 (the wrap's _eadistance accepts nothing)
        k_fill<<<256, 256, 0, st>>>(dout, rows * nw, __builtin_nan(""));
        WB_CK(cudaGetLastError());
      } else {
        // unscaled derivative metrics: the derivative of a window is the window of the derivative (EL:3220-3225)
        const int64_t Tp = (deriv && !wrap) ? T - 2 : T, mp = (deriv && !wrap) ? m - 2 : m;
        double *dxp = dx, *dsp = ds;
        if (deriv && !wrap) {
          if ((rc = ws.alloc(&dxp, (size_t)rows * Tp)) || (rc = ws.alloc(&dsp, (size_t)(nsub * mp)))) break;
          k_slope<<<1024, 256, 0, st>>>(dx, rows, (int)T, dxp);
          k_slope<<<256, 256, 0, st>>>(ds, nsub, (int)m, dsp);
          WB_CK(cudaGetLastError());
        }
        const double *dw = nullptr, *dtw = nullptr;
        if (J.metric == M_WDTW || J.metric == M_WDDTW || J.metric == M_WLCSS || J.metric == M_TWE) {  // wlcss: argmin mode only
          const int64_t Tw = J.weight_len > 0 ? J.weight_len : T;  // series the reference reset() the metric with
          const int64_t tn = J.metric == M_TWE ? T + 1 : (deriv ? Tw - 2 : Tw);
          ws.host_keep.push_back(J.metric == M_TWE ? make_tw(J.p.stiffness, tn) : make_weights(J.p.g, tn));
          std::vector<double>& h = ws.host_keep.back();
          double* d = nullptr;
          if ((rc = ws.alloc(&d, h.size()))) break;
          WB_CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
          (J.metric == M_TWE ? dtw : dw) = d + table_center(tn);
        }
        // thresholds: what the DP is abandoned against (thr_m, compared with M) and what the distance is compared with
        double thr_d = thr, thr_m = thr; int strict = wrap ? 1 : 0, apply_sqrt = 0;
        if (dtwfam) {
          thr_m = thr * thr;
          if (!wrap) { thr_d = thr * thr; apply_sqrt = 1; }  // unscaled + scaled_dtw: squared-cost domain, sqrt of a match
        } else if (J.metric == M_LCSS) thr_m = std::isinf(thr) ? thr : (double)m - thr * (double)m;
        else if (J.metric == M_EDR) thr_m = thr * (double)(wrap ? m : T);
        // unscaled edr with the default epsilon: the policy takes max(sx, sy) / 4 with sx = 4 * (std / 4), sy = 0
        double* edr_sx = nullptr;
        if (J.metric == M_EDR && !wrap && std::isnan(J.p.epsilon)) {
          ws.host_keep.emplace_back((size_t)nsub);
          std::vector<double>& h = ws.host_keep.back();
          for (int64_t k = 0; k < nsub; ++k) h[(size_t)k] = J.s_eps[paired ? lo + k : 0] * 4.0;
          if ((rc = ws.alloc(&edr_sx, (size_t)nsub))) break;
          WB_CK(cudaMemcpyAsync(edr_sx, h.data(), sizeof(double) * nsub, cudaMemcpyHostToDevice, st));
        }
        // pair-list entries are int2: at most 2^28 windows per pass and window offsets (i * T + w) below 2^31
        int64_t step = std::max<int64_t>(1, std::min<int64_t>(rows, std::min<int64_t>(((int64_t)1 << 28) / std::max<int64_t>(nw, 1), ((int64_t)1 << 31) / T - 1)));
        if (wrap) {
          int64_t budget = (int64_t)32 << 20;  // doubles of materialised windows per pass
          if (const char* e = getenv("WILDBOAR_CUDA_SCAN_WINDOW_BUDGET")) { const long long v = atoll(e); if (v > 0) budget = v; }
          step = std::max<int64_t>(1, std::min<int64_t>(step, budget / std::max<int64_t>(nw * m, 1)));
        }
        for (int64_t r0 = 0; r0 < rows && !rc; r0 += step) {
          const int64_t nr = std::min(step, rows - r0);
          const long long n = nr * nw;
          Workspace it(st);
          int2* list = nullptr; int* dlen = nullptr;
          double *draw = nullptr, *mraw = nullptr, *dkim = nullptr;
          if ((rc = it.alloc(&list, (size_t)n)) || (rc = it.alloc(&dlen, 1)) || (rc = it.alloc(&draw, (size_t)n)) ||
              (want_m && (rc = it.alloc(&mraw, (size_t)n)))) break;
          const int n32 = (int)n;
          WB_CK(cudaMemcpyAsync(dlen, &n32, sizeof(int), cudaMemcpyHostToDevice, st));
          DpCall c; memset(&c, 0, sizeof c);
          c.metric = J.metric; c.p = J.p; c.mode = PM_LISTP; c.list = list; c.list_len = dlen; c.list_n = n;
          long long ystride, ldk = 0;
          if (wrap) {
            double *mean = nullptr, *stdv = nullptr, *wn = nullptr;
            if ((rc = it.alloc(&mean, (size_t)n)) || (rc = it.alloc(&stdv, (size_t)n)) || (rc = it.alloc(&wn, (size_t)(n * m)))) break;
            if (identity) { k_fill<<<256, 256, 0, st>>>(mean, n, 0.0); k_fill<<<256, 256, 0, st>>>(stdv, n, 1.0); }
            else k_inc_window_stats<<<(unsigned)((nr + 63) / 64), 64, 0, st>>>(dx + r0 * T, nr, (int)T, (int)m, mean, stdv);
            k_normalise_windows<<<148 * 8, 256, 0, st>>>(dx + r0 * T, nr, (int)T, (int)m, mean, stdv, wn);
            WB_CK(cudaGetLastError());
            stats.launches += 2;
            c.x = ds; c.nx = nsub; c.Tx = (int)m; c.y = wn; c.ny = n; c.Ty = (int)m; c.ea = 1;
            if ((rc = prepare_operands(it, c))) break;
            if (dw) c.tab.weights = dw;
            if (want_m && (rc = interleave_y(it, c))) break;
            ystride = nw;
          } else if (ucr) {
            double *mean = nullptr, *stdv = nullptr;
            if ((rc = it.alloc(&mean, (size_t)(nr * T))) || (rc = it.alloc(&stdv, (size_t)(nr * T))) || (rc = it.alloc(&dkim, (size_t)(nr * T)))) break;
            k_window_stats<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(dx + r0 * T, nr, (int)T, (int)m, mean, stdv);
            k_ucr_kim<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dx + r0 * T, nr, (int)T, (int)m, ds + (paired ? r0 * m : 0), mean, stdv, dkim,
                                                                   paired ? m : 0);
            WB_CK(cudaGetLastError());
            stats.launches += 2;
            c.metric = M_SCALED_DTW;
            c.px = ds; c.nx = nsub; c.ptx = (int)m; c.py = dx + r0 * T; c.pty = (int)m; c.ys = 1; c.ny = nr * T - m + 1;
            c.R = (int)compute_warp_width(m, J.p.r) + 1; c.sy = mean; c.sy2 = stdv; c.raw = 1;
            ystride = T; ldk = T;
          } else {
            c.metric = J.metric == M_DDTW ? M_DTW : (J.metric == M_WDDTW ? M_WDTW : J.metric);  // the DP on prepared data
            c.px = dsp; c.nx = nsub; c.ptx = (int)mp; c.py = dxp + r0 * Tp; c.pty = (int)mp; c.ys = 1; c.ny = nr * Tp - mp + 1;
            c.R = (int)compute_r(m, J.p.r);  // from the ORIGINAL subsequence length (EL:2253, 2480)
            c.tab.weights = dw; c.tab.tw = dtw; c.raw = dtwfam ? 1 : 0;
            if (J.metric == M_ERP) {
              double *sx = nullptr, *sy = nullptr;
              if ((rc = it.alloc(&sx, (size_t)nsub)) || (rc = it.alloc(&sy, (size_t)c.ny))) break;
              k_series_stat<<<(unsigned)((nsub + 127) / 128), 128, 0, st>>>(c.px, nsub, (int)m, 0, J.p.g, sx, m);
              k_series_stat<<<(unsigned)((c.ny + 127) / 128), 128, 0, st>>>(c.py, c.ny, (int)m, 0, J.p.g, sy, 1);
              WB_CK(cudaGetLastError());
              stats.launches += 2;
              c.sx = sx; c.sy = sy;
            }
            if (edr_sx) c.sx = edr_sx;
            ystride = Tp;
          }
          k_profile_list<<<148 * 4, 256, 0, st>>>(list, n, (int)nw, ystride, (int)r0, paired ? 1 : 0);
          WB_CK(cudaGetLastError());
          if ((rc = launch_dp(it, di, c, 0, c.nx, 0, c.ny, draw, 0, mraw, nullptr, &stats))) break;
          if (K > 0) {
            ReplayArgs ra;
            ra.d = draw; ra.m = mraw; ra.lb = nullptr; ra.ld = nw; ra.nq = nr; ra.c0 = 0; ra.ncols = nw; ra.k = (int)K;
            const bool lcss = J.metric == M_LCSS || J.metric == M_WLCSS;
            ra.kind = dtwfam ? TK_SQUARE : (lcss ? TK_LCSS : (J.metric == M_EDR ? TK_SCALE : TK_IDENT));
            ra.scale = (lcss || J.metric == M_EDR) ? (double)m : 1.0;
            ra.tau = tau + r0; ra.hidx = hidx + r0 * K; ra.hval = hval + r0 * K; ra.hn = hn + r0;
            k_replay<<<(unsigned)std::max<long long>(1, std::min<long long>((nr + 3) / 4, 148 * 16)), 128, 0, st>>>(ra);
          } else
            k_profile_select<<<148 * 4, 256, 0, st>>>(draw, mraw, dkim, n, (int)nw, ldk, thr_d, thr_m, strict, apply_sqrt, dout + r0 * nw);
          WB_CK(cudaGetLastError());
          stats.launches += 2;
          for (auto& v : it.host_keep) ws.host_keep.push_back(std::move(v));
        }
        if (rc) break;
      }
      kt.stop();
      if (K > 0) {
        if (cudaMemcpyAsync(J.out_idx + lo * K, hidx, sizeof(long long) * rows * K, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(J.out_dist + lo * K, hval, sizeof(double) * rows * K, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the closest windows failed"); rc = 1; break; }
      } else if (cudaMemcpyAsync(J.out + lo * nw, dout, sizeof(double) * rows * nw, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess) { set_err("device-to-host copy of the distance profile failed"); rc = 1; break; }
      stats.kernel_ms = kt.ms();
    } while (0);
    total.stop();
    if (!rc) { cudaStreamSynchronize(st); stats.total_ms = total.ms(); }
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (st_out) *st_out = stats;
  return rc;
}

// rows of x in contiguous blocks over the devices, one host thread per device
template <class Worker>
static int run_row_sharded(int64_t nx, const int* devices, int n_devices, wb_stats* stats, Worker worker) {
  int ndev_avail = 0;
  if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail < 1) {
    set_err("no CUDA device available: wildboar_b200 has no CPU fallback");
    return 1;
  }
  std::vector<int> devs;
  if (devices && n_devices > 0) devs.assign(devices, devices + n_devices); else devs.push_back(0);
  for (int d : devs) if (d < 0 || d >= ndev_avail) { set_err("invalid device ordinal"); return 1; }
  const int G = (int)std::min<int64_t>((int64_t)devs.size(), std::max<int64_t>(nx, 1));
  std::vector<int64_t> off;
  row_blocks(nx, G, off);
  std::vector<wb_stats> sts((size_t)G);
  std::vector<int> rcs((size_t)G, 0);
  std::vector<std::string> errs((size_t)G);
  if (G == 1) { rcs[0] = worker(devs[0], off[0], off[1], &sts[0]); errs[0] = g_err; }
  else {
    std::vector<std::thread> th;
    for (int b = 0; b < G; ++b)
      th.emplace_back([&, b]() { rcs[(size_t)b] = worker(devs[(size_t)b], off[(size_t)b], off[(size_t)b + 1], &sts[(size_t)b]); errs[(size_t)b] = g_err; });
    for (auto& t : th) t.join();
  }
  for (int b = 0; b < G; ++b) if (rcs[(size_t)b]) { set_err(errs[(size_t)b]); return rcs[(size_t)b]; }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    for (int b = 0; b < G; ++b) {
      stats->kernel_ms = std::max(stats->kernel_ms, sts[(size_t)b].kernel_ms);
      stats->total_ms = std::max(stats->total_ms, sts[(size_t)b].total_ms);
      stats->cells += sts[(size_t)b].cells; stats->pairs += sts[(size_t)b].pairs; stats->launches += sts[(size_t)b].launches;
      stats->engine = std::max(stats->engine, sts[(size_t)b].engine);
    }
  }
  return 0;
}

static int run_subsequence(const SubseqJob& J, const int* devices, int n_devices, wb_stats* stats) {
  int ndev_avail = 0;
  if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail < 1) {
    set_err("no CUDA device available: wildboar_b200 has no CPU fallback");
    return 1;
  }
  std::vector<int> devs;
  if (devices && n_devices > 0) devs.assign(devices, devices + n_devices); else devs.push_back(0);
  for (int d : devs) if (d < 0 || d >= ndev_avail) { set_err("invalid device ordinal"); return 1; }
  const int G = (int)std::min<int64_t>((int64_t)devs.size(), std::max<int64_t>(J.nx, 1));
  std::vector<int64_t> off;
  row_blocks(J.nx, G, off);
  std::vector<wb_stats> sts((size_t)G);
  std::vector<int> rcs((size_t)G, 0);
  std::vector<std::string> errs((size_t)G);
  const auto worker = subseq_uses_scan(J) ? subseq_scan_worker : subseq_worker;
  if (G == 1) { rcs[0] = worker(J, devs[0], off[0], off[1], &sts[0]); errs[0] = g_err; }
  else {
    std::vector<std::thread> th;
    for (int b = 0; b < G; ++b)
      th.emplace_back([&, b]() { rcs[(size_t)b] = worker(J, devs[(size_t)b], off[(size_t)b], off[(size_t)b + 1], &sts[(size_t)b]); errs[(size_t)b] = g_err; });
    for (auto& t : th) t.join();
  }
  for (int b = 0; b < G; ++b) if (rcs[(size_t)b]) { set_err(errs[(size_t)b]); return rcs[(size_t)b]; }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    for (int b = 0; b < G; ++b) {
      stats->kernel_ms = std::max(stats->kernel_ms, sts[(size_t)b].kernel_ms);
      stats->total_ms = std::max(stats->total_ms, sts[(size_t)b].total_ms);
      stats->cells += sts[(size_t)b].cells; stats->pairs += sts[(size_t)b].pairs; stats->launches += sts[(size_t)b].launches;
      stats->engine = std::max(stats->engine, sts[(size_t)b].engine);
    }
  }
  return 0;
}

static int check_common(int metric, const wb_params* p, const void* x, int64_t n, int64_t T) {
  if (!p || !x) { set_err("null argument"); return 1; }
  if (metric < 0 || metric >= M_COUNT) { set_err("unknown metric id"); return 1; }
  if (n < 1 || T < 1) { set_err("empty input"); return 1; }
  if (T > (1 << 24)) { set_err("series too long"); return 1; }
  if (!(p->r >= 0.0 && p->r <= 1.0)) { set_err("r must be in [0, 1]"); return 1; }
  if (p->precision < 0 || p->precision > 2) { set_err("precision must be 0 (fp64, bit-exact), 1 (fp32) or 2 (fp64 with fused multiply-add)"); return 1; }
  return 0;
}

}  // namespace wb

using namespace wb;

extern "C" {

int wb_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* wb_cuda_last_error(void) { return g_err.c_str(); }

int wb_cuda_pairwise(int metric, const wb_params* params, const double* x, int64_t nx, int64_t Tx, int64_t x_stride,
                     const double* y, int64_t ny, int64_t Ty, int64_t y_stride, double* out, const int* devices,
                     int n_devices, wb_stats* stats) {
  if (check_common(metric, params, x, nx, Tx) || check_common(metric, params, y, ny, Ty)) return 1;
  if (!out) { set_err("null output"); return 1; }
  if (metric == M_WDDTW && Tx > Ty) { set_err("wddtw requires len(x) <= len(y) (the reference overflows a buffer otherwise)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 0; J.metric = metric; J.p = *params; J.x = x; J.nx = nx; J.Tx = Tx; J.xs = x_stride;
  J.y = y; J.ny = ny; J.Ty = Ty; J.ys = y_stride; J.out = out;
  return run_host_job(J, devices, n_devices, stats);
}

int wb_cuda_pairwise_self(int metric, const wb_params* params, const double* x, int64_t n, int64_t T,
                          int64_t x_stride, double* out, const int* devices, int n_devices, wb_stats* stats) {
  if (check_common(metric, params, x, n, T)) return 1;
  if (!out) { set_err("null output"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 1; J.metric = metric; J.p = *params; J.x = x; J.nx = n; J.Tx = T; J.xs = x_stride;
  J.y = x; J.ny = n; J.Ty = T; J.ys = x_stride; J.out = out;
  return run_host_job(J, devices, n_devices, stats);
}

int wb_cuda_paired(int metric, const wb_params* params, const double* x, int64_t n, int64_t Tx, int64_t x_stride,
                   const double* y, int64_t Ty, int64_t y_stride, double* out, const int* devices, int n_devices,
                   wb_stats* stats) {
  if (check_common(metric, params, x, n, Tx) || check_common(metric, params, y, n, Ty)) return 1;
  if (!out) { set_err("null output"); return 1; }
  if (metric == M_WDDTW && Ty > Tx) { set_err("wddtw (paired, operands swapped) requires len(y) <= len(x)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 2; J.metric = metric; J.p = *params; J.x = x; J.nx = n; J.Tx = Tx; J.xs = x_stride;
  J.y = y; J.ny = n; J.Ty = Ty; J.ys = y_stride; J.out = out;
  return run_host_job(J, devices, n_devices, stats);
}

static int check_nd(int64_t n_dims, int combine) {
  if (n_dims < 1) { set_err("n_dims must be >= 1"); return 1; }
  if (combine != 0 && combine != 1) { set_err("combine must be 0 (mean) or 1 (full)"); return 1; }
  return 0;
}

int wb_cuda_pairwise_nd(int metric, const wb_params* params, const double* x, int64_t nx, int64_t n_dims, int64_t Tx,
                        int64_t x_stride, int64_t x_dim_stride, const double* y, int64_t ny, int64_t Ty, int64_t y_stride,
                        int64_t y_dim_stride, int combine, double* out, const int* devices, int n_devices, wb_stats* stats) {
  if (check_common(metric, params, x, nx, Tx) || check_nd(n_dims, combine)) return 1;
  if (y && check_common(metric, params, y, ny, Ty)) return 1;
  if (!out) { set_err("null output"); return 1; }
  if (y && metric == M_WDDTW && Tx > Ty) { set_err("wddtw requires len(x) <= len(y) (the reference overflows a buffer otherwise)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.metric = metric; J.p = *params; J.x = x; J.nx = nx; J.Tx = Tx; J.xs = x_stride; J.xds = x_dim_stride;
  J.nd = n_dims; J.combine = combine; J.out = out;
  if (y) { J.kind = 0; J.y = y; J.ny = ny; J.Ty = Ty; J.ys = y_stride; J.yds = y_dim_stride; }
  else { J.kind = 1; J.y = x; J.ny = nx; J.Ty = Tx; J.ys = x_stride; J.yds = x_dim_stride; }
  return run_host_job(J, devices, n_devices, stats);
}

int wb_cuda_paired_nd(int metric, const wb_params* params, const double* x, int64_t n, int64_t n_dims, int64_t Tx,
                      int64_t x_stride, int64_t x_dim_stride, const double* y, int64_t Ty, int64_t y_stride,
                      int64_t y_dim_stride, int combine, double* out, const int* devices, int n_devices, wb_stats* stats) {
  if (check_common(metric, params, x, n, Tx) || check_common(metric, params, y, n, Ty) || check_nd(n_dims, combine)) return 1;
  if (!out) { set_err("null output"); return 1; }
  if (metric == M_WDDTW && Ty > Tx) { set_err("wddtw (paired, operands swapped) requires len(y) <= len(x)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 2; J.metric = metric; J.p = *params; J.x = x; J.nx = n; J.Tx = Tx; J.xs = x_stride; J.xds = x_dim_stride;
  J.y = y; J.ny = n; J.Ty = Ty; J.ys = y_stride; J.yds = y_dim_stride; J.nd = n_dims; J.combine = combine; J.out = out;
  return run_host_job(J, devices, n_devices, stats);
}

int wb_cuda_argmin(int metric, const wb_params* params, const double* x, int64_t nx, int64_t Tx, int64_t x_stride,
                   const double* y, int64_t ny, int64_t Ty, int64_t y_stride, int64_t k, const double* lower_bound,
                   int use_device_lb, int64_t* out_idx, double* out_dist, const int* devices, int n_devices,
                   wb_stats* stats) {
  if (check_common(metric, params, x, nx, Tx) || check_common(metric, params, y, ny, Ty)) return 1;
  if (!out_idx || !out_dist) { set_err("null output"); return 1; }
  if (k < 1 || k > ny) { set_err("k must satisfy 1 <= k <= n_y"); return 1; }
  if (metric == M_WDDTW && Tx > Ty) { set_err("wddtw requires len(x) <= len(y)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 3; J.metric = metric; J.p = *params; J.x = x; J.nx = nx; J.Tx = Tx; J.xs = x_stride;
  J.y = y; J.ny = ny; J.Ty = Ty; J.ys = y_stride; J.out = out_dist; J.out_idx = out_idx; J.k = k;
  J.lower_bound = lower_bound; J.use_device_lb = use_device_lb;
  return run_host_job(J, devices, n_devices, stats);
}

int wb_cuda_fit(const double* y, int64_t ny, int64_t n_dims, int64_t Ty, int64_t y_stride, int64_t y_dim_stride,
                const int* devices, int n_devices, wb_fitted** out) {
  if (!y || !out) { set_err("null argument"); return 1; }
  if (ny < 1 || Ty < 1 || n_dims < 1) { set_err("empty input"); return 1; }
  int ndev_avail = 0;
  if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail < 1) {
    set_err("no CUDA device available: wildboar_b200 has no CPU fallback");
    return 1;
  }
  wb_fitted* f = new wb_fitted;
  f->n = ny; f->nd = n_dims; f->T = Ty;
  if (devices && n_devices > 0) f->devs.assign(devices, devices + n_devices); else f->devs.push_back(0);
  int rc = 0;
  for (int d : f->devs) {
    if (d < 0 || d >= ndev_avail) { set_err("invalid device ordinal"); rc = 1; break; }
    DeviceInfo di;
    double* p = nullptr;
    if (cudaSetDevice(d) != cudaSuccess || device_info(&di)) { rc = 1; if (g_err.empty()) set_err("cudaSetDevice failed"); break; }
    // stream-ordered allocation from the device's retained pool (device_info): cudaMalloc / cudaFree cost up to
    // hundreds of ms next to a large retained pool, which would dominate short estimator calls
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { set_err("cudaStreamCreate failed"); rc = 1; break; }
    if (cudaMallocAsync((void**)&p, sizeof(double) * (size_t)n_dims * ny * Ty, st) != cudaSuccess) {
      cudaGetLastError(); cudaStreamDestroy(st); set_err("out of device memory for the fitted set"); rc = 1; break;
    }
    f->ptr.push_back(p);
    f->casc.emplace_back();
    for (int64_t k = 0; k < n_dims && !rc; ++k) rc = h2d_rows_staged(p + k * ny * Ty, y + k * y_dim_stride, ny, Ty, y_stride, st);
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { set_err("upload of the fitted set failed"); rc = 1; }
    cudaStreamDestroy(st);
    if (rc) break;
  }
  if (rc) { wb_cuda_fit_free(f); return rc; }
  *out = f;
  return 0;
}

void wb_cuda_fit_free(wb_fitted* f) {
  if (!f) return;
  for (size_t q = 0; q < f->ptr.size(); ++q) {
    // every library call synchronises before it returns, so no work is pending on the set
    if (cudaSetDevice(f->devs[q]) == cudaSuccess) {
      cudaFreeAsync(f->ptr[q], 0);
      if (q < f->casc.size() && f->casc[q].envT) {
        cudaFreeAsync(f->casc[q].envT, 0); cudaFreeAsync(f->casc[q].yvT, 0); cudaFreeAsync(f->casc[q].y0, 0); cudaFreeAsync(f->casc[q].yL, 0);
      }
    }
  }
  delete f;
}

int wb_cuda_pairwise_fitted(int metric, const wb_params* params, const double* x, int64_t nx, int64_t n_dims, int64_t Tx,
                            int64_t x_stride, int64_t x_dim_stride, const wb_fitted* fit, int combine, double* out,
                            wb_stats* stats) {
  if (!fit) { set_err("null fitted set"); return 1; }
  if (check_common(metric, params, x, nx, Tx) || check_nd(n_dims, combine)) return 1;
  if (n_dims != fit->nd) { set_err("x and the fitted set must have the same number of dimensions"); return 1; }
  if (!out) { set_err("null output"); return 1; }
  if (metric == M_WDDTW && Tx > fit->T) { set_err("wddtw requires len(x) <= len(y)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 0; J.metric = metric; J.p = *params; J.x = x; J.nx = nx; J.Tx = Tx; J.xs = x_stride; J.xds = x_dim_stride;
  J.ny = fit->n; J.Ty = fit->T; J.nd = n_dims; J.combine = combine; J.out = out; J.fit = fit;
  return run_host_job(J, fit->devs.data(), (int)fit->devs.size(), stats);
}

int wb_cuda_argmin_fitted(int metric, const wb_params* params, const double* x, int64_t nx, int64_t Tx, int64_t x_stride,
                          const wb_fitted* fit, int64_t k, const double* lower_bound, int use_device_lb, int64_t* out_idx,
                          double* out_dist, wb_stats* stats) {
  if (!fit) { set_err("null fitted set"); return 1; }
  if (check_common(metric, params, x, nx, Tx)) return 1;
  if (fit->nd != 1) { set_err("argmin needs a univariate fitted set"); return 1; }
  if (!out_idx || !out_dist) { set_err("null output"); return 1; }
  if (k < 1 || k > fit->n) { set_err("k must satisfy 1 <= k <= n_y"); return 1; }
  if (metric == M_WDDTW && Tx > fit->T) { set_err("wddtw requires len(x) <= len(y)"); return 1; }
  HostJob J; memset(&J, 0, sizeof J);
  J.kind = 3; J.metric = metric; J.p = *params; J.x = x; J.nx = nx; J.Tx = Tx; J.xs = x_stride;
  J.ny = fit->n; J.Ty = fit->T; J.out = out_dist; J.out_idx = out_idx; J.k = k;
  J.lower_bound = lower_bound; J.use_device_lb = use_device_lb; J.fit = fit;
  return run_host_job(J, fit->devs.data(), (int)fit->devs.size(), stats);
}

int wb_cuda_dtw_paths(const double* a, int64_t na, int64_t Ta, int64_t a_stride, const double* b, int64_t nb, int64_t Tb,
                      int64_t b_stride, const int64_t* ia, const int64_t* ib, int64_t n_pairs, double r,
                      const double* weights, int32_t* path_lo, int32_t* path_hi, double* cost, double* out_matrix,
                      int device, wb_stats* stats) {
  if (!a || !b || !path_lo || !path_hi) { set_err("null argument"); return 1; }
  if (na < 1 || nb < 1 || Ta < 1 || Tb < 1 || n_pairs < 1) { set_err("empty input"); return 1; }
  if (Ta > (1 << 24) || Tb > (1 << 24)) { set_err("series too long"); return 1; }
  if (!(r >= 0.0 && r <= 1.0)) { set_err("r must be in [0, 1]"); return 1; }
  if ((!ia && n_pairs > na) || (!ib && n_pairs > nb)) { set_err("without an index array, n_pairs must not exceed the number of series"); return 1; }
  return run_dtw_paths_host(a, na, Ta, a_stride, b, nb, Tb, b_stride, ia, ib, n_pairs, r, weights, path_lo, path_hi, cost,
                            out_matrix, device, stats);
}

int wb_cuda_dba_epoch(const wb_fitted* fit, int metric, const wb_params* params, const double* means_in, int64_t K, int64_t Tm,
                      const int64_t* member_offsets, const int64_t* members, const double* sample_weight,
                      const double* weights, int do_update, double* means_out, double* dist_out, wb_stats* stats) {
  if (!fit || !params || !means_in || !member_offsets || !members || !means_out || !dist_out) { set_err("null argument"); return 1; }
  if (fit->nd != 1) { set_err("DBA needs a univariate fitted set"); return 1; }
  if (metric != M_DTW && metric != M_WDTW) { set_err("DBA is defined for dtw and wdtw"); return 1; }
  if (K < 1 || Tm < 1 || member_offsets[0] != 0 || member_offsets[K] < 1) { set_err("empty input"); return 1; }
  if (!(params->r >= 0.0 && params->r <= 1.0)) { set_err("r must be in [0, 1]"); return 1; }
  if (metric == M_WDTW && do_update && !weights) { set_err("wdtw alignment needs the weight vector"); return 1; }
  return run_dba_epoch(fit, metric, *params, means_in, K, Tm, member_offsets, members, sample_weight, weights, do_update,
                       means_out, dist_out, stats);
}

int wb_cuda_subsequence(int metric, const wb_params* params, const double* s, const int64_t* s_offsets, int64_t n_s,
                        const double* x, int64_t nx, int64_t T, int64_t x_stride, int paired, int scaled,
                        const double* s_epsilon, double* out_dist, int64_t* out_idx, const int* devices, int n_devices,
                        wb_stats* stats) {
  if (check_common(metric, params, x, nx, T)) return 1;
  if (!s || !s_offsets || !out_dist || !out_idx) { set_err("null argument"); return 1; }
  if (metric == M_WLCSS) { set_err("wlcss has no subsequence metric in the reference (_distance.py:143-178)"); return 1; }
  if (params->precision != 0) { set_err("subsequence search runs in fp64"); return 1; }
  if (n_s < 1 || s_offsets[0] != 0) { set_err("empty input"); return 1; }
  if (paired && n_s != nx) { set_err("paired subsequence search needs one subsequence per sample"); return 1; }
  if (scaled && metric == M_DTW) for (int64_t k = 0; k < n_s; ++k)
    if (s_offsets[k + 1] - s_offsets[k] < 3) { set_err("scaled_dtw needs subsequences of at least 3 samples (the reference's LB_Kim reads S[1], S[2])"); return 1; }
  for (int64_t k = 0; k < n_s; ++k) {
    const int64_t m = s_offsets[k + 1] - s_offsets[k];
    if (m < 1 || m > T) { set_err("every subsequence needs 1 <= length <= n_timestep"); return 1; }
  }
  if (metric == M_EDR && !scaled) {
    // EdrSubsequenceMetric._distance (EL:2762-2765): the default epsilon is s_std / 4 of each subsequence, which the caller
    // computes with numpy exactly as ScaledSubsequenceMetric.from_array does (CD:453-467)
    if (std::isnan(params->epsilon) && !s_epsilon) { set_err("edr subsequence search with the default epsilon needs s_epsilon (std / 4 per subsequence)"); return 1; }
    if (s_epsilon) for (int64_t k = 0; k < n_s; ++k)
      if (!(s_epsilon[k] > 0.0)) { set_err("s_epsilon must be positive"); return 1; }
  }
  SubseqJob J;
  J.metric = metric; J.p = *params; J.s = s; J.soff = s_offsets; J.ns = n_s; J.x = x; J.nx = nx; J.T = T; J.xs = x_stride;
  J.paired = paired ? 1 : 0; J.scaled = scaled ? 1 : 0; J.s_eps = (metric == M_EDR && !scaled) ? s_epsilon : nullptr;
  J.out_dist = out_dist; J.out_idx = out_idx;
  return run_subsequence(J, devices, n_devices, stats);
}

int wb_cuda_subsequence_profile(int metric, const wb_params* params, const double* s, int64_t n_s, int64_t m,
                                const double* x, int64_t nx, int64_t T, int64_t x_stride, int scaled, const double* s_epsilon,
                                double threshold, double* out, const int* devices, int n_devices, wb_stats* stats) {
  if (check_common(metric, params, x, nx, T)) return 1;
  if (!s || !out) { set_err("null argument"); return 1; }
  if (metric == M_WLCSS) { set_err("wlcss has no subsequence metric in the reference (_distance.py:143-178)"); return 1; }
  if (params->precision != 0) { set_err("subsequence search runs in fp64"); return 1; }
  if (n_s != 1 && n_s != nx) { set_err("the profile needs one subsequence, or one per sample"); return 1; }
  if (m < 1 || m > T) { set_err("the subsequence needs 1 <= length <= n_timestep"); return 1; }
  if (scaled && metric == M_DTW && m < 3) { set_err("scaled_dtw needs subsequences of at least 3 samples (the reference's LB_Kim reads S[1], S[2])"); return 1; }
  if (std::isnan(threshold)) { set_err("threshold must not be NaN"); return 1; }
  if (metric == M_EDR && !scaled && std::isnan(params->epsilon)) {
    if (!s_epsilon) { set_err("edr with the default epsilon needs s_epsilon (std / 4 per subsequence)"); return 1; }
    for (int64_t k = 0; k < n_s; ++k) if (!(s_epsilon[k] >= 0.0)) { set_err("s_epsilon must not be negative"); return 1; }
  }
  wb::ProfileJob J;
  J.metric = metric; J.p = *params; J.s = s; J.ns = n_s; J.m = m; J.x = x; J.nx = nx; J.T = T; J.xs = x_stride;
  J.scaled = scaled ? 1 : 0; J.s_eps = s_epsilon; J.threshold = threshold; J.out = out;
  J.argmin_k = 0; J.out_idx = nullptr; J.out_dist = nullptr; J.weight_len = 0;
  return run_row_sharded(nx, devices, n_devices, stats,
                         [&](int dev, int64_t lo, int64_t hi, wb_stats* st) { return subseq_profile_worker(J, dev, lo, hi, st); });
}

int wb_cuda_subsequence_argmin(int metric, const wb_params* params, const double* s, int64_t n_s, int64_t m,
                               const double* x, int64_t nx, int64_t T, int64_t x_stride, int scaled, int64_t k,
                               int64_t weight_len, int64_t* out_idx, double* out_dist, const int* devices, int n_devices,
                               wb_stats* stats) {
  if (check_common(metric, params, x, nx, T)) return 1;
  if (!s || !out_idx || !out_dist) { set_err("null argument"); return 1; }
  if (params->precision != 0) { set_err("subsequence search runs in fp64"); return 1; }
  if (metric == M_WLCSS) { set_err("wlcss is not a subsequence metric of the reference (_distance.py:143-178, 1602-1614)"); return 1; }
  if (n_s != nx) { set_err("argmin_subsequence_distance pairs subsequence i with sample i"); return 1; }
  if (m < 1 || m > T) { set_err("the subsequence needs 1 <= length <= n_timestep"); return 1; }
  if (k < 1 || k > T - m + 1) { set_err("k must be in [1, n_timestep - m + 1]"); return 1; }
  if (weight_len != 0 && weight_len < T) { set_err("weight_len must be 0 or >= n_timestep"); return 1; }
  wb::ProfileJob J;
  J.metric = metric; J.p = *params; J.s = s; J.ns = n_s; J.m = m; J.x = x; J.nx = nx; J.T = T; J.xs = x_stride;
  J.scaled = scaled ? 1 : 0; J.s_eps = nullptr; J.threshold = WB_INF; J.out = nullptr;
  J.argmin_k = k; J.out_idx = out_idx; J.out_dist = out_dist; J.weight_len = weight_len;
  return run_row_sharded(nx, devices, n_devices, stats,
                         [&](int dev, int64_t lo, int64_t hi, wb_stats* st) { return subseq_profile_worker(J, dev, lo, hi, st); });
}

int wb_cuda_pairwise_dev(int metric, const wb_params* params, const double* d_x, int64_t nx, int64_t Tx,
                         const double* d_y, int64_t ny, int64_t Ty, double* d_out, void* stream, wb_stats* stats) {
  if (check_common(metric, params, d_x, nx, Tx) || check_common(metric, params, d_y, ny, Ty)) return 1;
  if (!d_out) { set_err("null output"); return 1; }
  if (metric == M_WDDTW && Tx > Ty) { set_err("wddtw requires len(x) <= len(y)"); return 1; }
  DeviceInfo di;
  if (device_info(&di)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  wb_stats local; memset(&local, 0, sizeof local);
  int rc = 0;
  double kms = 0;
  {
    Workspace ws(st);
    DpCall c; memset(&c, 0, sizeof c);
    c.metric = metric; c.p = *params; c.x = d_x; c.nx = nx; c.Tx = (int)Tx; c.y = d_y; c.ny = ny; c.Ty = (int)Ty;
    c.mode = PM_PAIRWISE;
    Timer kt(st);
    rc = prepare_operands(ws, c);
    if (!rc) {
      kt.start();
      rc = launch_dp(ws, di, c, 0, nx, 0, ny, d_out, ny, nullptr, nullptr, &local);
      kt.stop();
    }
    if (!rc && stats) kms = kt.ms();
  }
  if (rc) return rc;
  if (stats) { *stats = local; stats->kernel_ms = kms; stats->total_ms = kms; }
  return 0;
}

int wb_cuda_lb_keogh(const double* q, int64_t nq, int64_t q_stride, const double* x, int64_t nx, int64_t x_stride,
                     int64_t T, double r, int kind, double* out, int device, wb_stats* stats) {
  if (!q || !x || !out) { set_err("null argument"); return 1; }
  if (nq < 1 || nx < 1 || T < 1) { set_err("empty input"); return 1; }
  if (!(r >= 0.0 && r <= 1.0)) { set_err("r must be in [0, 1]"); return 1; }
  if (kind < 0 || kind > 2) { set_err("kind must be 0 (both), 1 (left) or 2 (right)"); return 1; }
  return run_lb(0, q, nq, q_stride, x, nx, x_stride, T, r, kind, out, device, stats);
}

int wb_cuda_lb_kim(const double* q, int64_t nq, int64_t q_stride, const double* x, int64_t nx, int64_t x_stride,
                   int64_t T, double* out, int device, wb_stats* stats) {
  if (!q || !x || !out) { set_err("null argument"); return 1; }
  if (nq < 1 || nx < 1 || T < 1) { set_err("empty input"); return 1; }
  return run_lb(1, q, nq, q_stride, x, nx, x_stride, T, 0.0, 0, out, device, stats);
}

int wb_cuda_dtw_envelope(const double* x, int64_t n, int64_t T, int64_t x_stride, int64_t w, double* lower, double* upper,
                         int device, wb_stats* stats) {
  if (!x || !lower || !upper) { set_err("null argument"); return 1; }
  if (n < 1 || T < 1) { set_err("empty input"); return 1; }
  if (T > (1 << 24)) { set_err("series too long"); return 1; }
  if (w < 0 || w >= T) { set_err("invalid r"); return 1; }  // EL:1077-1078
  return run_lb_series(0, x, n, T, x_stride, w, nullptr, nullptr, lower, upper, device, stats);
}

int wb_cuda_dtw_lb_keogh_terms(const double* x, const double* lower, const double* upper, int64_t n, int64_t T,
                               double* min_dist, double* cb, int device, wb_stats* stats) {
  if (!x || !lower || !upper || !min_dist || !cb) { set_err("null argument"); return 1; }
  if (n < 1 || T < 1) { set_err("empty input"); return 1; }
  if (T > (1 << 24)) { set_err("series too long"); return 1; }
  return run_lb_series(1, x, n, T, T, 0, lower, upper, min_dist, cb, device, stats);
}

void* wb_cuda_host_alloc(size_t bytes) { return pinned_alloc(bytes); }
void wb_cuda_host_free(void* p) { pinned_free(p); }

int wb_cuda_fp64_peak(int mix, double* inst_per_s, double* sm_mhz_est) {
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int threads = 256, per_sm = 4, iters = 1 << 16;  // 32 warps per SM: one resident wave
  const long long grid = (long long)di.sms * per_sm;
  double* out = nullptr; unsigned long long* cyc = nullptr;
  WB_CK(cudaMalloc(&out, sizeof(double) * grid * threads));
  WB_CK(cudaMalloc(&cyc, sizeof(unsigned long long)));
  cudaEvent_t a, b;
  WB_CK(cudaEventCreate(&a)); WB_CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    WB_CK(cudaEventRecord(a));
    k_fp64_peak<<<(unsigned)grid, threads>>>(mix, iters, 1.0 + rep, out, cyc);
    WB_CK(cudaEventRecord(b));
    WB_CK(cudaEventSynchronize(b));
    float ms = 0; WB_CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0) best = std::min(best, ms);
  }
  unsigned long long hc = 0;
  WB_CK(cudaMemcpy(&hc, cyc, sizeof hc, cudaMemcpyDeviceToHost));
  // FP64-pipe instructions per thread-iteration: mix 0: 8 DADD; mix 1: 4 cells x (DADD, DMUL, DADD, 2 DSETP)
  const double per_iter = mix == 0 ? 8.0 : 20.0;
  const double lane_inst = (double)grid * threads * (double)iters * per_iter;
  if (inst_per_s) *inst_per_s = lane_inst / (best * 1e-3);
  if (sm_mhz_est) *sm_mhz_est = (double)hc / (best * 1e-3) / 1e6;
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out); cudaFree(cyc);
  return 0;
}

}  // extern "C"
