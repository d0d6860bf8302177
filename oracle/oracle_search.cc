// oracle_search.cc — CPU ORACLE (test infrastructure): Flat, k-means, IVF-Flat.
//
// Restates the reference plugins' read path on top of the hooked distance kernels:
//   VectorIndexFlat::Search            src/vector/vector_index_flat.cc:205-264
//   VectorIndexIvfFlat::Search/Train   src/vector/vector_index_ivf_flat.cc:191-275, :644-712, :805-837
//   FillSearchResult (1 - ip mapping)  src/vector/vector_index_utils.cc:611-655
// and, FROM PUBLISHED ALGORITHMS (the fork dingodb/faiss@c50158c8 is not vendored — parity unpinned):
//   faiss exhaustive_{L2sqr,inner_product}_seq (always taken by the service path: one query per task,
//   src/vector/vector_index.cc:54,:256), IndexIVF::search_preassigned + IVFFlatScanner, faiss::Clustering.
// Tie rule (documented oracle choice): better value first, then smaller id.
#include <random>

#include "oracle_common.h"

using namespace oracle;

namespace {

inline bool is_ip(int metric) { return metric == ORACLE_IP || metric == ORACLE_COSINE; }

// one query against a contiguous run of rows
inline void scan_rows(bool ip, int32_t d, const float* q, const float* xb, const int64_t* ids, int64_t n0,
                      int64_t n1, const oracle_filter* filt, TopK& heap) {
  for (int64_t i = n0; i < n1; ++i) {
    const int64_t id = ids[i];
    if (id < 0) continue;  // removed slot
    if (!filter_pass(filt, id)) continue;
    const float v = ip ? oracle_fvec_inner_product(q, xb + i * (int64_t)d, d) : oracle_fvec_L2sqr(q, xb + i * (int64_t)d, d);
    heap.push(v, id);
  }
}

inline void emit(bool ip, int32_t k, TopK& heap, float* od, int64_t* oi) {
  heap.finish(od, oi);
  if (ip)
    for (int i = 0; i < k; ++i)
      if (oi[i] >= 0) od[i] = 1.0F - od[i];  // vector_index_utils.cc:632-634
}

// ---- faiss RandomGenerator / rand_perm (faiss/utils/random.cpp, public algorithm) ----
struct FaissRng {
  std::mt19937 mt;
  explicit FaissRng(int64_t seed) : mt((unsigned)seed) {}
  int rand_int(int max) { return mt() % max; }
  float rand_float() { return mt() / float(mt.max()); }
};
void rand_perm(std::vector<int64_t>& perm, int64_t n, int64_t seed) {
  perm.resize(n);
  for (int64_t i = 0; i < n; i++) perm[i] = i;
  FaissRng rng(seed);
  for (int64_t i = 0; i + 1 < n; i++) {
    int64_t i2 = i + rng.rand_int((int)(n - i));
    std::swap(perm[i], perm[i2]);
  }
}

}  // namespace

extern "C" {

// first-touch copy by `nthreads` workers (pages spread over the NUMA nodes the worker threads run on), so the timed
// CPU scans are not throttled by one memory controller; harness utility, not part of the restated path
void oracle_parallel_copy(void* dst, const void* src, size_t bytes, int nthreads) {
  const size_t block = 4u << 20;
  const int64_t nb = (int64_t)((bytes + block - 1) / block);
  parallel_for(nb, nthreads, [&](int64_t b) {
    const size_t off = (size_t)b * block;
    memcpy((char*)dst + off, (const char*)src + off, std::min(block, bytes - off));
  });
}

// test/unit_test/vector/test_vector_index_flat.cc:491-500
void oracle_fixture_mt19937(int64_t n, int32_t d, float* out) {
  std::mt19937 rng;
  std::uniform_real_distribution<> distrib;
  for (int64_t i = 0; i < n; i++) {
    for (int j = 0; j < d; j++) out[d * i + j] = distrib(rng);
    out[d * i] += i / 1000.;
  }
}

int oracle_flat_search(int metric, int32_t d, int64_t n, const float* xb, const int64_t* ids, int64_t nq,
                       const float* xq, int32_t k, const oracle_filter* filt, int nthreads, float* out_dist,
                       int64_t* out_ids) {
  if (k <= 0 || nq <= 0) return 0;
  const bool ip = is_ip(metric);
  parallel_for(nq, nthreads, [&](int64_t qi) {
    std::vector<float> qbuf(xq + qi * (int64_t)d, xq + (qi + 1) * (int64_t)d);
    if (metric == ORACLE_COSINE) oracle_normalize_faiss(qbuf.data(), d);  // vector_index_flat.cc:243
    TopK heap(k, ip);
    scan_rows(ip, d, qbuf.data(), xb, ids, 0, n, filt, heap);
    emit(ip, k, heap, out_dist + qi * (int64_t)k, out_ids + qi * (int64_t)k);
  });
  return 0;
}

int oracle_assign(int metric, int32_t d, int64_t n, const float* x, int32_t nlist, const float* centroids,
                  int nthreads, int32_t* out_assign) {
  const bool ip = is_ip(metric);
  parallel_blocks(n, nthreads, 256, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      const float* xi = x + i * (int64_t)d;
      int best = 0;
      float bv = ip ? oracle_fvec_inner_product(xi, centroids, d) : oracle_fvec_L2sqr(xi, centroids, d);
      for (int c = 1; c < nlist; ++c) {
        float v = ip ? oracle_fvec_inner_product(xi, centroids + c * (int64_t)d, d)
                     : oracle_fvec_L2sqr(xi, centroids + c * (int64_t)d, d);
        if (ip ? v > bv : v < bv) { bv = v; best = c; }
      }
      out_assign[i] = best;
    }
  });
  return 0;
}

// faiss::Clustering::train restatement (public algorithm; defaults mirrored by
// src/vector/vector_index_ivf_flat.cc:654-664: min 39 / max 256 points per centroid, seed 1234;
// niter: 10 for the IVF level-1 quantiser, 25 for PQ sub-quantisers).
int oracle_kmeans(int metric, int32_t d, int64_t n, const float* x_in, int32_t k, int32_t niter,
                  int32_t max_points_per_centroid, int64_t seed, int nthreads, float* centroids) {
  if (n < k) return -1;
  const float* x = x_in;
  std::vector<float> sub;
  if (n > (int64_t)k * max_points_per_centroid) {  // subsample_training_set
    std::vector<int64_t> perm;
    rand_perm(perm, n, seed);
    n = (int64_t)k * max_points_per_centroid;
    sub.resize(n * (int64_t)d);
    for (int64_t i = 0; i < n; ++i) memcpy(&sub[i * (int64_t)d], x_in + perm[i] * (int64_t)d, sizeof(float) * d);
    x = sub.data();
  }
  if (n == k) { memcpy(centroids, x, sizeof(float) * (size_t)n * d); return 0; }
  {
    std::vector<int64_t> perm;
    rand_perm(perm, n, seed + 1);  // redo = 0
    for (int i = 0; i < k; ++i) memcpy(centroids + i * (int64_t)d, x + perm[i] * (int64_t)d, sizeof(float) * d);
  }
  std::vector<int32_t> assign(n);
  std::vector<float> hassign(k);
  for (int it = 0; it < niter; ++it) {
    oracle_assign(metric, d, n, x, k, centroids, nthreads, assign.data());
    // compute_centroids: sums in point order, then scale by 1/count
    std::fill(hassign.begin(), hassign.end(), 0.0f);
    memset(centroids, 0, sizeof(float) * (size_t)k * d);
    // (faiss slices centroids across threads; per-centroid summation order is still point order)
    const int nt = std::max(1, nthreads);
    parallel_for(nt, nt, [&](int64_t t) {
      const int c0 = (int)((int64_t)k * t / nt), c1 = (int)((int64_t)k * (t + 1) / nt);
      for (int64_t i = 0; i < n; ++i) {
        const int ci = assign[i];
        if (ci >= c0 && ci < c1) {
          hassign[ci] += 1.0f;
          float* c = centroids + ci * (int64_t)d;
          const float* xi = x + i * (int64_t)d;
          for (int j = 0; j < d; ++j) c[j] += xi[j];
        }
      }
    });
    for (int ci = 0; ci < k; ++ci) {
      if (hassign[ci] == 0) continue;
      const float norm = 1 / hassign[ci];
      float* c = centroids + ci * (int64_t)d;
      for (int j = 0; j < d; ++j) c[j] *= norm;
    }
    // split_clusters
    const float EPS = 1 / 1024.;
    FaissRng rng(1234);
    for (int ci = 0; ci < k; ++ci) {
      if (hassign[ci] != 0) continue;
      int cj;
      for (cj = 0;; cj = (cj + 1) % k) {
        float p = (hassign[cj] - 1.0) / (float)(n - k);
        float r = rng.rand_float();
        if (r < p) break;
      }
      memcpy(centroids + ci * (int64_t)d, centroids + cj * (int64_t)d, sizeof(float) * d);
      for (int j = 0; j < d; ++j) {
        if (j % 2 == 0) { centroids[ci * (int64_t)d + j] *= 1 + EPS; centroids[cj * (int64_t)d + j] *= 1 - EPS; }
        else            { centroids[ci * (int64_t)d + j] *= 1 - EPS; centroids[cj * (int64_t)d + j] *= 1 + EPS; }
      }
      hassign[ci] = hassign[cj] / 2;
      hassign[cj] -= hassign[ci];
    }
  }
  return 0;
}

int oracle_ivfflat_search(int metric, int32_t d, int32_t nlist, const float* centroids, const int64_t* list_off,
                          const float* xb, const int64_t* ids, int64_t nq, const float* xq, int32_t k,
                          int32_t nprobe, const oracle_filter* filt, int nthreads, float* out_dist,
                          int64_t* out_ids) {
  if (k <= 0 || nq <= 0) return 0;
  const bool ip = is_ip(metric);
  if (nprobe <= 0) nprobe = 80;             // Constant::kSearchIvfFlatParamNprobe, src/common/constant.h:178
  nprobe = std::min(nprobe, nlist);         // vector_index_ivf_flat.cc:234
  parallel_for(nq, nthreads, [&](int64_t qi) {
    std::vector<float> qbuf(xq + qi * (int64_t)d, xq + (qi + 1) * (int64_t)d);
    if (metric == ORACLE_COSINE) oracle_normalize_faiss(qbuf.data(), d);
    const float* q = qbuf.data();
    // coarse quantiser: IndexFlatL2 / IndexFlatIP over the centroids, top-nprobe (ties -> smaller list id)
    TopK coarse(nprobe, ip);
    for (int c = 0; c < nlist; ++c)
      coarse.push(ip ? oracle_fvec_inner_product(q, centroids + c * (int64_t)d, d)
                     : oracle_fvec_L2sqr(q, centroids + c * (int64_t)d, d), c);
    std::vector<float> cd(nprobe);
    std::vector<int64_t> ci(nprobe);
    coarse.finish(cd.data(), ci.data());
    TopK heap(k, ip);
    for (int p = 0; p < nprobe; ++p) {
      if (ci[p] < 0) continue;
      scan_rows(ip, d, q, xb, ids, list_off[ci[p]], list_off[ci[p] + 1], filt, heap);
    }
    emit(ip, k, heap, out_dist + qi * (int64_t)k, out_ids + qi * (int64_t)k);
  });
  return 0;
}

// Pairwise distance matrix: CalcDistanceEntry -> CalcDistanceCore (src/vector/vector_index_utils.cc:48-124), one
// DoCalc*Distance call per (left, right) pair (:193-419).  faiss::VectorDistance<METRIC_L2 / INNER_PRODUCT> and
// the hnswlib space functions are the hooked fvec_L2sqr / fvec_inner_product (src/vector/vector_index.cc:153-185);
// hnswlib's InnerProductSpace already returns 1 - ip, the faiss flavour subtracts explicitly (:262) - same value.
int oracle_calc_distance(int algorithm, int metric, int32_t d, int64_t nl, const float* left, int64_t nr,
                         const float* right, float* out, float* left_out, float* right_out) {
  if ((algorithm != 1 && algorithm != 2) || (metric != ORACLE_L2 && metric != ORACLE_IP && metric != ORACLE_COSINE)) return -1;
  if (d <= 0 || nl < 0 || nr < 0) return -1;
  std::vector<float> ln((size_t)nl * d), rn((size_t)nr * d);
  auto prep = [&](const float* src, float* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
      const float* s = src + i * (int64_t)d;
      float* t = dst + i * (int64_t)d;
      if (metric != ORACLE_COSINE) { std::memcpy(t, s, (size_t)d * 4); continue; }
      if (algorithm == 1) { std::memcpy(t, s, (size_t)d * 4); oracle_normalize_faiss(t, d); }  // :283-284
      else oracle_normalize_hnsw(s, d, t);                                                        // :407-413
    }
  };
  prep(left, ln.data(), nl);
  prep(right, rn.data(), nr);
  for (int64_t i = 0; i < nl; ++i)
    for (int64_t j = 0; j < nr; ++j) {
      const float* a = ln.data() + i * (int64_t)d;
      const float* b = rn.data() + j * (int64_t)d;
      out[i * nr + j] = metric == ORACLE_L2 ? oracle_fvec_L2sqr(a, b, d) : 1.0f - oracle_fvec_inner_product(a, b, d);
    }
  if (left_out) std::memcpy(left_out, ln.data(), ln.size() * 4);
  if (right_out) std::memcpy(right_out, rn.data(), rn.size() * 4);
  return 0;
}

}  // extern "C"
