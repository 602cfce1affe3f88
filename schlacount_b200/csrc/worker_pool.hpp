// Fixed pool of host threads with one operation: run fn(i) for i in [0, n) and wait. Used by the host driver to parse
// BAM records in parallel (BAM ingest off the critical path). Product code.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace hlac {

class WorkerPool {
 public:
  explicit WorkerPool(unsigned n_threads) {
    for (unsigned i = 1; i < n_threads; i++) workers_.emplace_back([this] { run(); });   // the caller is worker 0
  }
  ~WorkerPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  unsigned size() const { return (unsigned)workers_.size() + 1; }
  // fn(begin, end) over [0, n) in chunks of `grain`; rethrows the first exception a chunk raised
  void for_ranges(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn) {
    if (n == 0) return;
    if (workers_.empty() || n <= grain) { fn(0, n); return; }
    { std::lock_guard<std::mutex> l(m_); fn_ = &fn; n_ = n; grain_ = grain; next_.store(0); active_ = (unsigned)workers_.size(); err_ = nullptr; gen_++; }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> l(m_);
    done_cv_.wait(l, [this] { return active_ == 0; });
    fn_ = nullptr;
    if (err_) std::rethrow_exception(err_);
  }

 private:
  void work() {
    for (;;) {
      const size_t b = next_.fetch_add(grain_);
      if (b >= n_) break;
      try { (*fn_)(b, std::min(n_, b + grain_)); }
      catch (...) { std::lock_guard<std::mutex> l(m_); if (!err_) err_ = std::current_exception(); }
    }
  }
  void run() {
    unsigned long long seen = 0;
    for (;;) {
      { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return gen_ != seen; }); seen = gen_; if (stop_) return; }
      work();
      { std::lock_guard<std::mutex> l(m_); if (--active_ == 0) done_cv_.notify_one(); }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t, size_t)>* fn_ = nullptr;
  size_t n_ = 0, grain_ = 1;
  std::atomic<size_t> next_{0};
  unsigned active_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  std::exception_ptr err_;
};

}  // namespace hlac
