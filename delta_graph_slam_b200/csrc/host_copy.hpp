// b200reg — large host-to-host copies between a caller's PAGEABLE cloud and the library's page-locked staging buffers.
//
// pcl::PointCloud storage is pageable, so the literal drop-in stages every cloud it takes or returns: a 2 MB raw scan, the
// 0.8 MB filtered cloud twice (out of the filter, into the registration) and the 0.8 MB aligned cloud — ~4.3 MB of memcpy
// per frame on the calling thread, which at ~10 GB/s of one core was 430 us of a 512 us frame (bench e2e_pageable): the
// host, not the GPU, bound that leg.  Copies of 256 KB and more are therefore split over a small pool of helper threads
// (created on first use, parked on a condition variable in between, never destroyed: a process-lifetime singleton).
// A caller that finds the pool busy (another handle's thread is copying) copies alone, as before.
#pragma once
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

namespace b200 {

class HostCopyPool {
 public:
  static constexpr int kHelpers = 3;
  static constexpr size_t kMinBytes = 256 * 1024;

  static void copy(void* dst, const void* src, size_t bytes) {
    if (bytes < kMinBytes) { memcpy(dst, src, bytes); return; }
    HostCopyPool* p = instance();
    if (!p || !p->busy_.try_lock()) { memcpy(dst, src, bytes); return; }
    p->run(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    p->busy_.unlock();
  }

 private:
  struct Part { char* d; const char* s; size_t n; };
  std::mutex busy_;  // one split copy at a time
  std::mutex m_;
  std::condition_variable cv_job_, cv_done_;
  Part parts_[kHelpers];
  unsigned long long generation_ = 0;
  int pending_ = 0;
  bool ok_ = false;

  static HostCopyPool* instance() {
    static HostCopyPool* p = create();  // leaked on purpose: helper threads may outlive static destruction
    return p;
  }
  static HostCopyPool* create() {
    HostCopyPool* p = new (std::nothrow) HostCopyPool();
    if (!p) return nullptr;
    try {
      for (int w = 0; w < kHelpers; ++w) std::thread(&HostCopyPool::worker, p, w).detach();
      p->ok_ = true;
    } catch (...) {
      p->ok_ = false;  // no helpers: copy() falls back to memcpy through run()'s own-part path only
    }
    return p->ok_ ? p : nullptr;
  }
  void worker(int w) {
    unsigned long long seen = 0;
    for (;;) {
      Part job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_job_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        job = parts_[w];
      }
      if (job.n) memcpy(job.d, job.s, job.n);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) cv_done_.notify_one();
      }
    }
  }
  void run(char* d, const char* s, size_t bytes) {
    // kHelpers + 1 parts, 4 KB granular; the caller takes the last (and any remainder)
    const size_t part = (bytes / (kHelpers + 1)) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(m_);
      for (int w = 0; w < kHelpers; ++w) parts_[w] = Part{d + (size_t)w * part, s + (size_t)w * part, part};
      pending_ = kHelpers;
      ++generation_;
    }
    cv_job_.notify_all();
    memcpy(d + (size_t)kHelpers * part, s + (size_t)kHelpers * part, bytes - (size_t)kHelpers * part);
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
  }
};

inline void host_copy(void* dst, const void* src, size_t bytes) { HostCopyPool::copy(dst, src, bytes); }

}  // namespace b200
