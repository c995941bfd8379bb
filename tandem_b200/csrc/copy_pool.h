// Tiny persistent thread pool for the host-side staging copies of the DrMvsnet boundary: the API contract hands us
// pageable caller memory that is only valid during the call (inputs) or must be filled before returning (outputs), so a
// CPU copy into / out of pinned memory is unavoidable; at 6.45 MB in + 4.9 MB out per keyframe a single memcpy thread
// (~10 GB/s) would cost more than the whole FeatureNet.  Jobs are (dst, src, bytes, optional callback run after the copy).
#pragma once
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace tdm {

class CopyPool {
 public:
  struct Job {
    void* dst;
    const void* src;
    size_t bytes;
    std::function<void()> after;   // runs on the copying thread once the bytes are in place (may be empty)
  };

  explicit CopyPool(int nthreads) {
    for (int i = 0; i < nthreads; ++i) threads_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }

  // Runs all jobs (the calling thread helps) and returns when every job and its callback has finished.
  void run(std::vector<Job>& jobs) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      jobs_ = &jobs;
      next_ = 0;
      pending_ = (int)jobs.size();
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    jobs_ = nullptr;
  }

 private:
  bool take(Job*& j) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!jobs_ || next_ >= (int)jobs_->size()) return false;
    j = &(*jobs_)[next_++];
    return true;
  }
  void work() {
    Job* j;
    while (take(j)) {
      std::memcpy(j->dst, j->src, j->bytes);
      if (j->after) j->after();
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || (jobs_ && next_ < (int)jobs_->size()); });
        if (stop_) return;
      }
      work();
    }
  }

  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::vector<Job>* jobs_ = nullptr;
  int next_ = 0, pending_ = 0;
  bool stop_ = false;
};

}  // namespace tdm
