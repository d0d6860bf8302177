// oracle_pq.cc — CPU ORACLE (test infrastructure): product quantiser + IVF-PQ search.
//
// Reference call sites: faiss::IndexIVFPQ(quantizer, d, nlist, M, nbits, metric)
//   src/vector/vector_index_raw_ivf_pq.cc:551-571 (Init), :457-500 (Train), :157-210 (Search;
//   nprobe default 80 clamped to nlist :170,:190; cosine = normalise + IP).
// The arithmetic lives in dingodb/faiss@c50158c8 (NOT vendored) — restated here from the published
// IndexIVFPQ algorithm ("parity unpinned"):
//   * by_residual = true for every metric: codes quantise x - centroid[list].
//   * sub-quantiser training: one k-means per sub-space (ksub = 2^nbits centroids, 25 iterations,
//     seed 1234, <=256 points per centroid), L2 assignment.
//   * encode: per sub-vector argmin L2 (first minimum wins).
//   * IP search:  dis = <q, c_list> + sum_m <q_m, codeword[m][code_m]>     (larger is better)
//   * L2 search:  dis = ||q - c_list||^2 + sum_m (T[list][m][code_m] - 2 <q_m, codeword[m][code_m]>)
//                 with the precomputed table T = ||r||^2 + 2 <c_list_m, r>   (use_precomputed_table,
//                 whose memory the reference accounts for at raw_ivf_pq.cc:447-450).
// ORACLE CHOICES: sub-space distances/inner products use the plain sequential order
// (src/simd/distances_ref.cc:23-56 — faiss's own fvec_*_ny twins are un-vendored); LUT sums run
// sequentially over m.  IVF-PQ parity is a tolerance gate (1e-4 relative, recall@k 1e-3), not bit-exact.
#include "oracle_common.h"

using namespace oracle;

namespace {
inline bool is_ip(int metric) { return metric == ORACLE_IP || metric == ORACLE_COSINE; }
}

extern "C" {

int oracle_pq_train(int32_t d, int32_t M, int32_t nbits, int64_t n, const float* x, int32_t niter, int64_t seed,
                    int nthreads, float* codebooks) {
  if (d % M != 0) return -1;
  const int dsub = d / M, ksub = 1 << nbits;
  if (n < ksub) return -2;
  std::vector<float> slice((size_t)n * dsub);
  for (int m = 0; m < M; ++m) {
    for (int64_t i = 0; i < n; ++i) memcpy(&slice[i * (int64_t)dsub], x + i * (int64_t)d + m * dsub, sizeof(float) * dsub);
    // sub-space k-means with SEQUENTIAL-order L2 (dsub is tiny) — local Lloyd to keep the order explicit
    std::vector<float> cent((size_t)ksub * dsub);
    // reuse the generic k-means but with sequential distance: implement inline
    // (init/subsample identical to oracle_kmeans)
    int rc = oracle_kmeans(ORACLE_L2, dsub, n, slice.data(), ksub, niter, 256, seed, nthreads, cent.data());
    if (rc != 0) return rc;
    memcpy(codebooks + (size_t)m * ksub * dsub, cent.data(), sizeof(float) * (size_t)ksub * dsub);
  }
  return 0;
}

int oracle_pq_encode(int32_t d, int32_t M, int32_t nbits, const float* codebooks, int64_t n, const float* x,
                     int nthreads, uint8_t* codes) {
  if (nbits != 8) return -1;
  const int dsub = d / M, ksub = 1 << nbits;
  parallel_blocks(n, nthreads, 256, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i)
      for (int m = 0; m < M; ++m) {
        const float* xs = x + i * (int64_t)d + m * dsub;
        const float* cb = codebooks + (size_t)m * ksub * dsub;
        int best = 0;
        float bv = oracle_fvec_L2sqr_seq(xs, cb, dsub);
        for (int j = 1; j < ksub; ++j) {
          float v = oracle_fvec_L2sqr_seq(xs, cb + (size_t)j * dsub, dsub);
          if (v < bv) { bv = v; best = j; }
        }
        codes[i * (int64_t)M + m] = (uint8_t)best;
      }
  });
  return 0;
}

int oracle_ivfpq_encode(int32_t d, int32_t M, int32_t nbits, const float* codebooks, int32_t nlist,
                        const float* centroids, int64_t n, const float* x, const int32_t* assign, int nthreads,
                        uint8_t* codes) {
  (void)nlist;
  std::vector<float> res((size_t)n * d);
  for (int64_t i = 0; i < n; ++i) {
    const float* c = centroids + (int64_t)assign[i] * d;
    for (int j = 0; j < d; ++j) res[i * (int64_t)d + j] = x[i * (int64_t)d + j] - c[j];
  }
  return oracle_pq_encode(d, M, nbits, codebooks, n, res.data(), nthreads, codes);
}

int oracle_ivfpq_search(int metric, int32_t d, int32_t nlist, int32_t M, int32_t nbits, const float* centroids,
                        const float* codebooks, const int64_t* list_off, const uint8_t* codes,
                        const int64_t* ids, int64_t nq, const float* xq, int32_t k, int32_t nprobe,
                        const oracle_filter* filt, int nthreads, float* out_dist, int64_t* out_ids) {
  if (k <= 0 || nq <= 0) return 0;
  if (nbits != 8) return -1;
  const bool ip = is_ip(metric);
  const int dsub = d / M, ksub = 1 << nbits;
  if (nprobe <= 0) nprobe = 80;      // Constant::kSearchIvfPqParamNprobe, src/common/constant.h:188
  nprobe = std::min(nprobe, nlist);  // raw_ivf_pq.cc:190
  // L2: precomputed table T[list][m][j] = ||cw||^2 + 2 <c_list_m, cw>
  std::vector<float> pre;
  if (!ip) {
    pre.resize((size_t)nlist * M * ksub);
    parallel_for(nlist, nthreads, [&](int64_t l) {
      for (int m = 0; m < M; ++m)
        for (int j = 0; j < ksub; ++j) {
          const float* cw = codebooks + ((size_t)m * ksub + j) * dsub;
          const float* cs = centroids + l * (int64_t)d + m * dsub;
          float r2 = oracle_fvec_inner_product_seq(cw, cw, dsub);
          float cr = oracle_fvec_inner_product_seq(cs, cw, dsub);
          pre[((size_t)l * M + m) * ksub + j] = r2 + 2 * cr;
        }
    });
  }
  parallel_for(nq, nthreads, [&](int64_t qi) {
    std::vector<float> qbuf(xq + qi * (int64_t)d, xq + (qi + 1) * (int64_t)d);
    if (metric == ORACLE_COSINE) oracle_normalize_faiss(qbuf.data(), d);
    const float* q = qbuf.data();
    TopK coarse(nprobe, ip);
    for (int c = 0; c < nlist; ++c)
      coarse.push(ip ? oracle_fvec_inner_product(q, centroids + c * (int64_t)d, d)
                     : oracle_fvec_L2sqr(q, centroids + c * (int64_t)d, d), c);
    std::vector<float> cd(nprobe);
    std::vector<int64_t> ci(nprobe);
    coarse.finish(cd.data(), ci.data());
    // per-query inner-product table  sim[m][j] = <q_m, codeword[m][j]>
    std::vector<float> sim((size_t)M * ksub), tab((size_t)M * ksub);
    for (int m = 0; m < M; ++m)
      for (int j = 0; j < ksub; ++j)
        sim[(size_t)m * ksub + j] = oracle_fvec_inner_product_seq(q + m * dsub, codebooks + ((size_t)m * ksub + j) * dsub, dsub);
    TopK heap(k, ip);
    for (int p = 0; p < nprobe; ++p) {
      if (ci[p] < 0) continue;
      const int64_t l = ci[p];
      const float dis0 = cd[p];
      const float* t = sim.data();
      if (!ip) {  // fvec_madd(n, precomputed, -2, sim, tab)
        const float* pl = &pre[(size_t)l * M * ksub];
        for (size_t i = 0; i < (size_t)M * ksub; ++i) tab[i] = pl[i] + (-2.0f) * sim[i];
        t = tab.data();
      }
      for (int64_t r = list_off[l]; r < list_off[l + 1]; ++r) {
        const int64_t id = ids[r];
        if (id < 0 || !filter_pass(filt, id)) continue;
        const uint8_t* code = codes + r * (int64_t)M;
        float dis = dis0;
        for (int m = 0; m < M; ++m) dis += t[(size_t)m * ksub + code[m]];
        heap.push(dis, id);
      }
    }
    float* od = out_dist + qi * (int64_t)k;
    int64_t* oi = out_ids + qi * (int64_t)k;
    heap.finish(od, oi);
    if (ip)
      for (int i = 0; i < k; ++i)
        if (oi[i] >= 0) od[i] = 1.0F - od[i];
  });
  return 0;
}

}  // extern "C"
