// oracle_common.h — shared helpers of the CPU oracle (TEST INFRASTRUCTURE, see oracle.h).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "oracle.h"

namespace oracle {

// Dynamic 1-item-per-task scheduling: mirrors the reference's execution shape on this path
// (FLAGS_vector_read_batch_size_per_task = 1, src/vector/vector_index.cc:54,:256; pool of N
// std::threads, src/server/server.cc:868-873).
inline void parallel_for(int64_t n, int nthreads, const std::function<void(int64_t)>& fn) {
  if (nthreads <= 1 || n <= 1) {
    for (int64_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> ts;
  int nt = (int)std::min<int64_t>(nthreads, n);
  ts.reserve(nt);
  for (int t = 0; t < nt; ++t)
    ts.emplace_back([&] {
      for (;;) {
        int64_t i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto& t : ts) t.join();
}

// Block-partitioned variant for bulk work (assignment, encoding).
inline void parallel_blocks(int64_t n, int nthreads, int64_t block,
                            const std::function<void(int64_t, int64_t)>& fn) {
  int64_t nb = (n + block - 1) / block;
  parallel_for(nb, nthreads, [&](int64_t b) { fn(b * block, std::min(n, (b + 1) * block)); });
}

inline bool filter_pass(const oracle_filter* f, int64_t id) {
  if (!f) return true;
  if (f->has_range && !(id >= f->range_min && id < f->range_max)) return false;  // vector_index.h:79
  if (f->sorted_ids) {
    bool exist = std::binary_search(f->sorted_ids, f->sorted_ids + f->n_ids, id);  // vector_index.h:125-141
    if (f->negate ? exist : !exist) return false;
  }
  return true;
}

// Result collector.  "key" is the quantity being MINIMISED (L2: distance; IP: -ip is NOT used —
// for IP we keep ip and maximise, see below).  Total order = (better value first, then smaller id):
// the oracle's documented tie rule (faiss keeps an arbitrary heap order on exact ties; ids are the
// only layout-independent tie-break — DESIGN.md §Ties).
struct TopK {
  int k;
  bool maximize;  // true for inner product
  std::vector<float> v;
  std::vector<int64_t> id;
  int n = 0;
  TopK(int k_, bool maximize_) : k(k_), maximize(maximize_), v(k_), id(k_) {}
  // strictly-better ordering
  inline bool better(float a, int64_t ia, float b, int64_t ib) const {
    if (a != b) return maximize ? a > b : a < b;
    return ia < ib;
  }
  // heap keeps the WORST element at index 0
  inline void sift_down(int i) {
    for (;;) {
      int l = 2 * i + 1, r = l + 1, w = i;
      if (l < n && better(v[w], id[w], v[l], id[l])) w = l;
      if (r < n && better(v[w], id[w], v[r], id[r])) w = r;
      if (w == i) break;
      std::swap(v[i], v[w]); std::swap(id[i], id[w]); i = w;
    }
  }
  inline void sift_up(int i) {
    while (i > 0) {
      int p = (i - 1) / 2;
      if (better(v[p], id[p], v[i], id[i])) { std::swap(v[i], v[p]); std::swap(id[i], id[p]); i = p; }
      else break;
    }
  }
  inline void push(float val, int64_t vid) {
    if (n < k) { v[n] = val; id[n] = vid; ++n; sift_up(n - 1); }
    else if (better(val, vid, v[0], id[0])) { v[0] = val; id[0] = vid; sift_down(0); }
  }
  inline float worst() const { return v[0]; }
  inline bool full() const { return n == k; }
  // write out best-first; pad with id -1
  void finish(float* out_v, int64_t* out_id) {
    std::vector<int> ord(n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return better(v[a], id[a], v[b], id[b]); });
    for (int i = 0; i < k; ++i) {
      if (i < n) { out_v[i] = v[ord[i]]; out_id[i] = id[ord[i]]; }
      else { out_v[i] = 0.0f; out_id[i] = -1; }
    }
  }
};

}  // namespace oracle
