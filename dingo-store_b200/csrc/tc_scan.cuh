// tc_scan.cuh — tensor-core candidate pass for Flat / IVF-Flat search (sm_100a: TMA + tcgen05 + TMEM).
//
// What it replaces: the per-query CPU scans IVFFlatScanner::scan_codes / exhaustive_*_seq behind
// VectorIndexIvfFlat::Search and VectorIndexFlat::Search (src/vector/vector_index_ivf_flat.cc:247-251,
// vector_index_flat.cc:249-252) and the IndexFlat coarse quantiser of IVF (vector_index_ivf_flat.cc:805-837).
// The reference streams every probed list once PER QUERY on one core; here each probed list chunk is streamed
// from HBM ONCE PER BATCH and multiplied against all queries that probe it — the one place on this path where
// the work is a dense contraction.
//
// Exactness: the MMA runs in TF32 (operands truncated to 10 mantissa bits), so its scores only SELECT
// candidates.  Per query the engine keeps every row whose approximate score is within 2*eps of the k-th best
// approximate score (eps = rigorous bound on the TF32 score error); all of those are re-scored by the exact
// FP32 kernel in the reference's AVX-512 order, so the returned ids and distances are identical to the exact
// path.  Queries whose window cannot be certified (candidate overflow / threshold too tight) are re-run on
// the exact scan.  See DESIGN.md §Tensor-core candidate pass.
#pragma once
#include <cuda.h>

#include "filter.cuh"
#include "index.h"

namespace b200vs {

constexpr int TC_BM = 128;        // database rows per MMA tile (UMMA M)
constexpr int TC_BK = 32;         // floats per K block = 128 B = one swizzle span
constexpr int TC_NQT = 128;       // queries per work item (UMMA N <= 128): a probed list chunk is re-streamed only beyond 128 queries
constexpr int TC_STAGES = 5;      // TMA->MMA ring depth (5 x 32 KB; + 24 KB of capture staging: ~40 KB per SM stay free for co-resident small kernels)
constexpr int TC_THREADS = 256;   // warp0 TMA + scheduler, warp1 MMA, warp2 TMEM alloc, warps4-7 epilogue
constexpr int TC_CHUNK = 512;     // rows per work item (list chunk): fine grain for dynamic load balance
constexpr int TC_SPAN = 2048;     // one sampled chunk per TC_SPAN rows of a list
constexpr int TC_HITS = 512;      // captured rows staged per epilogue warp between flushes
constexpr int TC_SAMPLE = 32;     // sampled rows per span for the threshold estimate
constexpr int TC_SQ = 4;          // scheduler queue depth (items the producer may run ahead)
constexpr uint32_t TC_A_BYTES = TC_BM * 128;
constexpr uint32_t TC_B_BYTES = TC_NQT * 128;
constexpr size_t TC_SMEM = (size_t)TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024;

struct TcItem {
  int list;
  int row_begin, row_end;  // rows within the list
  int pair_begin;          // first row of this item's query group in the gathered-query workspace
  int nq;                  // queries in the group (1..TC_NQT)
  int sample_slot;         // >= 0 when this item is a sampled chunk
  int pad[2];
};

struct TcParams {
  const long long* ids;
  const float* norms;
  const long long* list_off;
  int d;
  const TcItem* items;
  const int* totals;        // [0] n_items  [1] n_pairs  [2] n_sample_items
  const int* sample_list;   // [n_sample_items] item indices
  int* work_counter;        // dynamic scheduler: next work index
  const int* pair_query;    // [npairs] query index of each gathered row
  int mode;                 // 0 = sample pass, 1 = capture pass, 2 = dense pass (every score written)
  float* sample;            // mode 0: [sample_items, TC_NQT, sample_rows]
  int sample_rows;          // TC_SAMPLE (32-row box) or TC_BM (the whole first tile of each span)
  const float* tau;         // mode 1: [nq] capture threshold on the approximate score
  unsigned long long* cand; // mode 1: [nq, cap]  (ord(score) << 32 | arena row)
  int* cand_cnt;            // mode 1: [nq]
  int cap;
  float* dense;             // mode 2: [nq, dense_ld] approximate scores (row = column index)
  long long dense_ld;
  int dense_accum;          // mode 2: add to the stored score instead of overwriting (split-precision passes 2 and 3)
  int add_norm;             // include the ||x||^2 term (off for the correction passes)
  int b_gather;             // B tiles gathered straight from the rounded query matrix with TMA gather4 (no gathered copy)
  int split;                // error-compensated TF32: hi*hi + lo*hi + hi*lo into one accumulator (coarse quantiser)
  int l2;
  FilterDev filt;
};

// view of an index the tile scan can run on (lists contiguous in one arena)
struct TcView {
  const float* vecs = nullptr;
  const long long* ids = nullptr;
  const float* norms = nullptr;
  int64_t arena_rows = 0;               // rows addressable by the A tensor map
  const long long* list_off = nullptr;  // device [nlist]
  const int* list_len = nullptr;        // device [nlist]
  int nlist = 0;
  int64_t total_chunks = 0;  // sum over lists of ceil(len / TC_CHUNK)   (host bookkeeping)
  int max_chunks_per_list = 0;
  float max_norm = 0.f;      // max ||x|| over the index (host bookkeeping, monotone)
  const float* vecs_hi = nullptr;  // optional split of the rows for the error-compensated coarse pass:
  const float* vecs_lo = nullptr;  //   hi = x with the low 13 mantissa bits cleared, lo = x - hi (both exact)
  bool flat = false;         // single list covering rows [0, arena_rows)
  long long id_offset = 0;   // tc_coarse on a slice of the centroid table: added to the returned row indices
  float owned_frac = 1.f;    // share of the lists that hold rows here (a list-sharded rank owns 1 / world of them)
  bool api_scores = false;   // tc_coarse: return the ascending ranking score (L2 distance | -ip) instead of the raw metric value
};

// true when this search can use the tensor-core pass (otherwise the caller runs the exact scan)
bool tc_eligible(const IndexBase* ix, const TcView& v, int64_t nq, int k, int nprobe, const SearchCtx& sc);

// probes: device [nq, nprobe] list indices (for Flat: all zeros, nprobe = 1).  q: device queries, already
// normalised for cosine.  Writes API-semantics results; uncertified queries are re-run on the exact scan.
void tc_search(IndexBase* ix, const TcView& v, bool l2, int64_t nq, const float* q, int k, const long long* probes,
               int nprobe, const SearchCtx& sc, float* out_dist, long long* out_ids, cudaStream_t s);

// Exact top-nprobe of every query against a small row set (the IVF coarse quantiser): dense TF32 score matrix on the
// tensor cores, then window select + exact FP32 re-score.  Always certified (every score is available).
// out_probes [nq, nprobe] row indices; out_raw (optional) [nq, nprobe] raw metric values (L2 distance / ip).
bool tc_coarse_eligible(const IndexBase* ix, int64_t nq, int nrows, int nprobe);
void tc_coarse(IndexBase* ix, const TcView& v, bool l2, int64_t nq, const float* q, int nprobe, long long* out_probes,
               float* out_raw, cudaStream_t s);

float device_max_norm(IndexBase* ix, const float* norms_sq, int64_t n, cudaStream_t s);  // sqrt(max ||x||^2)
void launch_row_norms(const float* x, int64_t n, int d, float* out, cudaStream_t s);     // out[i] = <x_i, x_i>
void launch_split_rows(const float* x, int64_t n, int d, float* hi, float* lo, cudaStream_t s);

}  // namespace b200vs
