// eventgrad_b200 -- native host-side batch prefetcher (C++17, no CUDA, no Python in the loop).
//
// Replaces the host half of the reference's LibTorch DataLoader (sampler order -> per-sample fetch ->
// Stack, /root/reference/dcifar10/event/event.cpp:93-105): a worker thread walks the epoch's index
// order and gathers each batch (raw uint8 sample rows + int64 labels) into a ring of caller-provided
// staging slots -- pinned host memory on GPU runs, so the trainer only has to issue the async H2D copy.
// The consumer side is a bounded producer/consumer queue: next() blocks until the batch is staged,
// release() hands the slot back.  No GIL is held while gathering.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace egb {

class HostPrefetcher {
 public:
  HostPrefetcher(const uint8_t* images, const int64_t* labels, int64_t n_samples, int64_t sample_bytes,
                 int64_t batch, std::vector<uint8_t*> slot_images, std::vector<int64_t*> slot_labels)
      : images_(images), labels_(labels), n_(n_samples), sb_(sample_bytes), batch_(batch),
        simg_(std::move(slot_images)), slab_(std::move(slot_labels)),
        state_(simg_.size(), kFree), count_(simg_.size(), 0) {}

  ~HostPrefetcher() { stop(); }

  // Start gathering the batches of one epoch; `order` (length n_order) is copied.
  void start_epoch(const int64_t* order, int64_t n_order) {
    stop();
    order_.assign(order, order + n_order);
    n_batches_ = (n_order + batch_ - 1) / batch_;
    produced_ = consumed_ = 0;
    for (auto& s : state_) s = kFree;
    quit_ = false;
    error_ = false;
    worker_ = std::thread([this] { run(); });
  }

  int64_t num_batches() const { return n_batches_; }

  // Blocks until the next batch is staged. Returns {slot, count}; slot = -1 at end of epoch / on error.
  std::pair<int, int64_t> next() {
    std::unique_lock<std::mutex> lk(mu_);
    if (consumed_ >= n_batches_) return {-1, 0};
    const int slot = (int)(consumed_ % (int64_t)simg_.size());
    cv_.wait(lk, [&] { return state_[slot] == kReady || error_; });
    if (error_) return {-1, -1};
    state_[slot] = kInUse;
    ++consumed_;
    return {slot, count_[slot]};
  }

  void release(int slot) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (slot >= 0 && slot < (int)state_.size() && state_[slot] == kInUse) state_[slot] = kFree;
    }
    cv_.notify_all();
  }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
  }

 private:
  enum : int { kFree = 0, kReady = 1, kInUse = 2 };

  void run() {
    for (int64_t b = 0; b < n_batches_; ++b) {
      const int slot = (int)(b % (int64_t)simg_.size());
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return state_[slot] == kFree || quit_; });
        if (quit_) return;
      }
      const int64_t beg = b * batch_;
      const int64_t end = std::min<int64_t>(beg + batch_, (int64_t)order_.size());
      uint8_t* di = simg_[slot];
      int64_t* dl = slab_[slot];
      bool ok = true;
      for (int64_t j = beg; j < end; ++j) {
        const int64_t idx = order_[j];
        if (idx < 0 || idx >= n_) {
          ok = false;
          break;
        }
        std::memcpy(di + (j - beg) * sb_, images_ + idx * sb_, (size_t)sb_);
        dl[j - beg] = labels_[idx];
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!ok) error_ = true;
        count_[slot] = end - beg;
        state_[slot] = kReady;
        ++produced_;
      }
      cv_.notify_all();
      if (!ok) return;
    }
  }

  const uint8_t* images_;
  const int64_t* labels_;
  int64_t n_, sb_, batch_;
  std::vector<uint8_t*> simg_;
  std::vector<int64_t*> slab_;
  std::vector<int> state_;
  std::vector<int64_t> count_;
  std::vector<int64_t> order_;
  int64_t n_batches_ = 0, produced_ = 0, consumed_ = 0;
  bool quit_ = false, error_ = false;
  std::mutex mu_;
  std::condition_variable cv_;
  std::thread worker_;
};

}  // namespace egb
