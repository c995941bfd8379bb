// Tiny persistent thread pool for the host-side staging copies of the DrMvsnet boundary: the API contract hands us
// caller memory that is only valid during the call (inputs) or must be filled before returning (outputs); when that
// memory is pageable a CPU copy into / out of pinned memory is unavoidable, and at 6.45 MB in + 4.9 MB out per keyframe a
// single memcpy thread (~10 GB/s) would cost more than the whole FeatureNet.  (Page-locked caller buffers skip the pool
// altogether: mvsnet.cu DMA's straight from / into them.)
// ONE pool per process, shared by every handle (round 1 gave each handle three threads of its own: 8 ranks x 4 handles x 4
// threads fought over the host cores and the end-to-end scaling dropped to 0.55 at 8 GPUs).  Size: TDM_COPY_THREADS, default 3.
// Several handles may call run() concurrently: every call is a batch; the calling thread always works on its own batch,
// pool threads help whichever batch still has unclaimed jobs.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
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

  static CopyPool& shared() {
    static CopyPool pool([] {
      const char* e = std::getenv("TDM_COPY_THREADS");
      const int n = e ? std::atoi(e) : 3;
      return n < 0 ? 0 : (n > 64 ? 64 : n);
    }());
    return pool;
  }

  // Runs all jobs (the calling thread helps) and returns when every job and its callback has finished.
  void run(std::vector<Job>& jobs) {
    if (jobs.empty()) return;
    Batch b;
    b.jobs = &jobs;
    b.pending = (int)jobs.size();
    {
      std::lock_guard<std::mutex> lk(mu_);
      batches_.push_back(&b);
    }
    cv_.notify_all();
    work(b);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&b] { return b.pending == 0; });
    batches_.remove(&b);
  }

 private:
  struct Batch {
    std::vector<Job>* jobs = nullptr;
    int next = 0, pending = 0;   // guarded by mu_
  };
  bool take(Batch& b, Job*& j) {
    std::lock_guard<std::mutex> lk(mu_);
    if (b.next >= (int)b.jobs->size()) return false;
    j = &(*b.jobs)[b.next++];
    return true;
  }
  void work(Batch& b) {
    Job* j;
    while (take(b, j)) {
      std::memcpy(j->dst, j->src, j->bytes);
      if (j->after) j->after();
      std::lock_guard<std::mutex> lk(mu_);
      if (--b.pending == 0) done_.notify_all();
    }
  }
  Batch* open_batch() {   // mu_ held
    for (Batch* b : batches_)
      if (b->next < (int)b->jobs->size()) return b;
    return nullptr;
  }
  void loop() {
    for (;;) {
      Job* j = nullptr;
      Batch* b = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || open_batch() != nullptr; });
        if (stop_) return;
        b = open_batch();
        j = &(*b->jobs)[b->next++];   // claimed under the lock: the batch cannot retire before pending reaches 0
      }
      std::memcpy(j->dst, j->src, j->bytes);
      if (j->after) j->after();
      std::lock_guard<std::mutex> lk(mu_);
      if (--b->pending == 0) done_.notify_all();
    }
  }

  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::list<Batch*> batches_;
  bool stop_ = false;
};

}  // namespace tdm
