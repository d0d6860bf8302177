/*
 * oracle.h — C ABI of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a CPU restatement of dingo-store's vector-search hot path
 * (reference: /root/reference, dingodb/dingo-store @ dc8c439c).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product library (libb200vs.so) never links or calls it.
 *
 * Parity pinning status:
 *   - distance arithmetic (fvec_L2sqr / fvec_inner_product): PINNED bit-for-bit against the
 *     reference's own src/simd compiled from /root/reference into oracle/_ref/ and against
 *     the known-answer values recorded in SURVEY.md §0 / tests/golden/simd_kat.json.
 *   - normalisers, result marshalling, defaults/clamps: restated from the src/vector plugins (cited
 *     per function) — no numeric goldens exist in the reference's tests ("parity unpinned"
 *     beyond the behavioural contract).
 *   - faiss / hnswlib algorithms (k-means, IVF scan, PQ, HNSW): the pinned forks
 *     dingodb/faiss@c50158c8 and dingodb/hnswlib@1964db3e are NOT vendored; restated from the
 *     published algorithms — "parity unpinned", anchored on the reference call sites.
 */
#ifndef B200VS_ORACLE_H_
#define B200VS_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* metric codes mirror pb::common::MetricType (L2=1, INNER_PRODUCT=2, COSINE=3) */
enum { ORACLE_L2 = 1, ORACLE_IP = 2, ORACLE_COSINE = 3 };

/* ---- distance primitives: AVX-512 accumulation order of src/simd/distances_avx512.cc:48-143 ---- */
float oracle_fvec_L2sqr(const float* x, const float* y, size_t d);
float oracle_fvec_inner_product(const float* x, const float* y, size_t d);
/* plain sequential order of src/simd/distances_ref.cc:23-53 (for cross-checks) */
float oracle_fvec_L2sqr_seq(const float* x, const float* y, size_t d);
float oracle_fvec_inner_product_seq(const float* x, const float* y, size_t d);

/* ---- normalisers: src/vector/vector_index_utils.cc:480-491 (faiss flavour), :493-500 (hnsw flavour) ---- */
void oracle_normalize_faiss(float* x, int32_t d);                       /* in place */
void oracle_normalize_hnsw(const float* x, int32_t d, float* out);

/* ---- fixtures: default-seeded std::mt19937 + uniform_real_distribution<> generator of
 *      test/unit_test/vector/test_vector_index_flat.cc:491-500 ---- */
void oracle_fixture_mt19937(int64_t n, int32_t d, float* out);

/* ---- filters (vector_index.h:67-146): AND of an optional id range and an optional sorted id list ---- */
typedef struct {
  int has_range; int64_t range_min, range_max;            /* RangeFilterFunctor: min <= id < max */
  const int64_t* sorted_ids; int64_t n_ids; int negate;   /* SortFilterFunctor */
} oracle_filter;

/* ---- Flat exact search: vector_index_flat.cc:205-264 + faiss IndexFlat/IndexIDMap2 semantics.
 * xb row-major [n,d] ALREADY normalised for cosine (as stored by Add); xq RAW (normalised inside
 * for cosine).  out_dist in API semantics (L2: squared L2; IP/cosine: 1-ip), ascending; out_ids -1 padded.
 * ids[i] < 0 marks a removed slot. */
int oracle_flat_search(int metric, int32_t d, int64_t n, const float* xb, const int64_t* ids,
                       int64_t nq, const float* xq, int32_t k, const oracle_filter* filt,
                       int nthreads, float* out_dist, int64_t* out_ids);

/* ---- k-means, faiss::Clustering restatement (ivf_flat.cc:644-712 calls index_->train) ---- */
int oracle_kmeans(int metric, int32_t d, int64_t n, const float* x, int32_t k, int32_t niter,
                  int32_t max_points_per_centroid, int64_t seed, int nthreads, float* centroids /*[k,d]*/);
/* coarse assignment used by add_with_ids: argmin L2 / argmax IP, ties -> smaller centroid index */
int oracle_assign(int metric, int32_t d, int64_t n, const float* x, int32_t nlist,
                  const float* centroids, int nthreads, int32_t* out_assign);

/* ---- IVF-Flat search: vector_index_ivf_flat.cc:191-275 + faiss IndexIVFFlat semantics.
 * Inverted lists given list-major: list l owns rows [list_off[l], list_off[l+1]) of xb/ids. */
int oracle_ivfflat_search(int metric, int32_t d, int32_t nlist, const float* centroids,
                          const int64_t* list_off, const float* xb, const int64_t* ids,
                          int64_t nq, const float* xq, int32_t k, int32_t nprobe,
                          const oracle_filter* filt, int nthreads, float* out_dist, int64_t* out_ids);

/* ---- Product quantiser (faiss::ProductQuantizer restatement) and IVF-PQ (raw_ivf_pq.cc:157-210) ---- */
int oracle_pq_train(int32_t d, int32_t M, int32_t nbits, int64_t n, const float* x, int32_t niter,
                    int64_t seed, int nthreads, float* codebooks /*[M, 2^nbits, d/M]*/);
int oracle_pq_encode(int32_t d, int32_t M, int32_t nbits, const float* codebooks, int64_t n,
                     const float* x, int nthreads, uint8_t* codes /*[n,M]*/);
/* by_residual encoding helper: residual = x - centroid[assign] then encode */
int oracle_ivfpq_encode(int32_t d, int32_t M, int32_t nbits, const float* codebooks, int32_t nlist,
                        const float* centroids, int64_t n, const float* x, const int32_t* assign,
                        int nthreads, uint8_t* codes);
int oracle_ivfpq_search(int metric, int32_t d, int32_t nlist, int32_t M, int32_t nbits,
                        const float* centroids, const float* codebooks, const int64_t* list_off,
                        const uint8_t* codes, const int64_t* ids, int64_t nq, const float* xq,
                        int32_t k, int32_t nprobe, const oracle_filter* filt, int nthreads,
                        float* out_dist, int64_t* out_ids);

/* ---- HNSW (hnswlib::HierarchicalNSW restatement; vector_index_hnsw.cc:135-184, :203-254, :318-485) ---- */
typedef struct oracle_hnsw oracle_hnsw;
oracle_hnsw* oracle_hnsw_create(int metric, int32_t d, int64_t max_elements, int32_t M,
                                int32_t ef_construction, int64_t seed);
void oracle_hnsw_destroy(oracle_hnsw*);
/* x RAW (hnsw-normalised inside for cosine); single-threaded insertion in the given order */
int oracle_hnsw_add(oracle_hnsw*, int64_t n, const float* x, const int64_t* labels);
int oracle_hnsw_search(oracle_hnsw*, int64_t nq, const float* xq, int32_t k, int32_t ef,
                       const oracle_filter* filt, int nthreads, float* out_dist, int64_t* out_ids,
                       int64_t* out_ndis /*[nq] or NULL*/, int64_t* out_hops /*[nq] or NULL*/);
/* export the graph in the flat layout the GPU index loads (see include/b200vs.h, B200VS_STATE_HNSW) */
int64_t oracle_hnsw_export_size(oracle_hnsw*);
int oracle_hnsw_export(oracle_hnsw*, void* blob, int64_t len);
int oracle_hnsw_import(oracle_hnsw*, const void* blob, int64_t len);  /* search a graph built elsewhere (same layout as export) */

/* harness utility: multi-threaded first-touch copy (NUMA-spread pages for the timed CPU baseline) */
void oracle_parallel_copy(void* dst, const void* src, size_t bytes, int nthreads);

/* ---- pairwise distance matrix: VectorIndexUtils::CalcDistanceEntry / CalcDistanceCore
 * (src/vector/vector_index_utils.cc:48-124) with the per-pair functions DoCalc{L2,Ip,Cosine}DistanceBy{Faiss,Hnswlib}
 * (:193-419).  algorithm 1 = ALGORITHM_FAISS, 2 = ALGORITHM_HNSWLIB.  out[i*nr + j] = distance(left i, right j):
 * L2 -> squared L2; IP -> 1 - ip; COSINE -> 1 - ip of the normalised copies (faiss flavour: NormalizeVectorForFaiss
 * :480-491; hnswlib flavour: NormalizeVectorForHnsw :493-500).  left_out / right_out (nullable, [nl,d] / [nr,d]) receive
 * what is_return_normlize returns: the normalised copies for COSINE, the inputs otherwise. */
int oracle_calc_distance(int algorithm, int metric, int32_t d, int64_t nl, const float* left, int64_t nr,
                         const float* right, float* out, float* left_out, float* right_out);

const char* oracle_version(void);

#ifdef __cplusplus
}
#endif
#endif
