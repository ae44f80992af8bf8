// Host-side helper of the pipelined reference upload (argmin.cuh, ensure_refs): rows of a PAGEABLE host array are copied into
// a page-locked staging buffer by a few threads at once, from where they reach the device by asynchronous DMA.  The driver's own
// pageable path stages through one thread (about 10 GB/s on the GPU boxes: 40 ms for the 410 MB of cfg4's references, more than
// the kernels of the whole call); four copy threads reach the DMA's pace.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace wb {

class HostCopyPool {
 public:
  // leaked singleton: the workers are detached daemons that sleep on a condition variable (nothing to join at exit)
  static HostCopyPool& get() { static HostCopyPool* p = new HostCopyPool; return *p; }
  int threads() const { return nworkers_ + 1; }

  // dst (dense rows of row_bytes) <- rows of src that start src_stride bytes apart; returns when every byte is in place
  void copy_rows(char* dst, const char* src, size_t rows, size_t row_bytes, size_t src_stride) {
    const size_t nparts = (size_t)std::min<size_t>((size_t)nworkers_ + 1, std::max<size_t>(1, rows * row_bytes >> 20));  // >= 1 MB per part
    if (nparts <= 1) { run(Job{dst, src, rows, row_bytes, src_stride, nullptr}); return; }
    std::atomic<int> left((int)nparts - 1);
    const size_t per = (rows + nparts - 1) / nparts;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (size_t p = 1; p < nparts; ++p) {
        const size_t r0 = std::min(rows, p * per), r1 = std::min(rows, (p + 1) * per);
        q_.push_back(Job{dst + r0 * row_bytes, src + r0 * src_stride, r1 - r0, row_bytes, src_stride, &left});
      }
    }
    cv_.notify_all();
    run(Job{dst, src, std::min(rows, per), row_bytes, src_stride, nullptr});
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return left.load() == 0; });
  }

 private:
  struct Job { char* dst; const char* src; size_t rows, row_bytes, src_stride; std::atomic<int>* left; };
  static void run(const Job& j) {
    if (j.rows == 0) return;
    if (j.src_stride == j.row_bytes) { memcpy(j.dst, j.src, j.rows * j.row_bytes); return; }
    for (size_t r = 0; r < j.rows; ++r) memcpy(j.dst + r * j.row_bytes, j.src + r * j.src_stride, j.row_bytes);
  }
  HostCopyPool() {
    int n = 3;  // + the calling thread
    if (const char* e = getenv("WILDBOAR_CUDA_COPY_THREADS")) n = std::max(0, std::min(15, atoi(e) - 1));
    const unsigned hw = std::thread::hardware_concurrency();
    if (hw && (int)hw / 2 - 1 < n) n = std::max(0, (int)hw / 2 - 1);
    nworkers_ = n;
    for (int w = 0; w < n; ++w)
      std::thread([this] {
        for (;;) {
          Job j;
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return !q_.empty(); });
            j = q_.front();
            q_.pop_front();
          }
          run(j);
          if (j.left && j.left->fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(mu_);  // the waiter checks the counter under this lock: no lost wake-up
            done_cv_.notify_all();
          }
        }
      }).detach();
  }
  int nworkers_ = 0;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::deque<Job> q_;
};

}  // namespace wb
