// tc_scan.cu — tensor-core candidate pass (see tc_scan.cuh for the contract, DESIGN.md for the proof sketch).
//
// Kernels:
//   tc_prep_queries_kernel   round queries to TF32 (RN), ||q||^2
//   tc_count_pairs_kernel    invert the probe table: how many queries probe each list
//   tc_plan_kernel           exclusive scans -> pair offsets, work-item offsets
//   tc_fill_pairs_kernel     gather the (TF32-rounded) queries of every list into one contiguous block
//   tc_items_kernel          emit work items (list chunk x query group)
//   tc_scan_kernel           *** the hot kernel: TMA -> smem ring -> tcgen05.mma (TF32) -> TMEM -> epilogue ***
//   tc_tau_kernel            per-query capture threshold from the sampled rows
//   tc_final_kernel          per-query window select + exact FP32 (AVX-512 order) rerank + certification
//   tc_compact_flags_kernel  list of uncertified queries for the exact re-run
#include <algorithm>
#include <cmath>

#include "scan_kernels.cuh"
#include "sm100_ptx.cuh"
#include "tc_scan.cuh"

namespace b200vs {

using namespace ptx;

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
__global__ void tc_prep_queries_kernel(const float* __restrict__ q, long long nq, int d, float* __restrict__ q32,
                                       float* __restrict__ qnorm) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nq) return;
  const float* src = q + (size_t)w * d;
  float* dst = q32 + (size_t)w * d;
  float acc = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = src[i];
    acc = fmaf(v, v, acc);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    dst[i] = __uint_as_float(r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) qnorm[w] = acc;
}

__global__ void tc_count_pairs_kernel(const long long* __restrict__ probes, long long n, int* cnt, int* pos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long l = probes[i];
  pos[i] = l >= 0 ? atomicAdd(cnt + l, 1) : -1;
}

// single block: exclusive scans over the lists
__global__ void tc_plan_kernel(const int* __restrict__ cnt, const int* __restrict__ list_len, int nlist, int* pair_off,
                               int* item_off, int* totals /*[0]=n_items [1]=n_pairs*/) {
  __shared__ int s_pairs[1024], s_items[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (nlist + T - 1) / T;
  const int b = t * per, e = min(nlist, b + per);
  int sp = 0, si = 0;
  for (int l = b; l < e; ++l) {
    const int c = cnt[l], len = list_len[l];
    sp += c;
    if (c > 0 && len > 0) si += ((c + TC_NQT - 1) / TC_NQT) * ((len + TC_CHUNK - 1) / TC_CHUNK);
  }
  s_pairs[t] = sp; s_items[t] = si;
  __syncthreads();
  if (t == 0) {
    int ap = 0, ai = 0;
    for (int i = 0; i < T; ++i) { const int p = s_pairs[i], q = s_items[i]; s_pairs[i] = ap; s_items[i] = ai; ap += p; ai += q; }
    totals[0] = ai; totals[1] = ap;
  }
  __syncthreads();
  sp = s_pairs[t]; si = s_items[t];
  for (int l = b; l < e; ++l) {
    const int c = cnt[l], len = list_len[l];
    pair_off[l] = sp; item_off[l] = si;
    sp += c;
    if (c > 0 && len > 0) si += ((c + TC_NQT - 1) / TC_NQT) * ((len + TC_CHUNK - 1) / TC_CHUNK);
  }
}

// one warp per (query, probe): place the pair and copy the rounded query row
__global__ void tc_fill_pairs_kernel(const long long* __restrict__ probes, const int* __restrict__ pos,
                                     const int* __restrict__ pair_off, long long n, int nprobe, int d,
                                     const float* __restrict__ q32, int* pair_query, int* pair_of, float* bws) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long l = probes[w];
  if (l < 0) { if (lane == 0) pair_of[w] = -1; return; }
  const int p = pair_off[l] + pos[w];
  const int q = (int)(w / nprobe);
  if (lane == 0) { pair_query[p] = q; pair_of[w] = p; }
  const float4* src = reinterpret_cast<const float4*>(q32 + (size_t)q * d);
  float4* dst = reinterpret_cast<float4*>(bws + (size_t)p * d);
  for (int i = lane; i < (d >> 2); i += 32) dst[i] = src[i];
}

__global__ void tc_items_kernel(const int* __restrict__ cnt, const int* __restrict__ list_len,
                                const int* __restrict__ pair_off, const int* __restrict__ item_off, int nlist, TcItem* items) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlist) return;
  const int c = cnt[l], len = list_len[l];
  if (c <= 0 || len <= 0) return;
  const int ng = (c + TC_NQT - 1) / TC_NQT, nc = (len + TC_CHUNK - 1) / TC_CHUNK;
  TcItem* out = items + item_off[l];
  for (int g = 0; g < ng; ++g)
    for (int ch = 0; ch < nc; ++ch) {
      TcItem it;
      it.list = l;
      it.row_begin = ch * TC_CHUNK;
      it.row_end = min(len, (ch + 1) * TC_CHUNK);
      it.pair_begin = pair_off[l] + g * TC_NQT;
      it.nq = min(TC_NQT, c - g * TC_NQT);
      it.pad[0] = it.pad[1] = it.pad[2] = 0;
      out[g * nc + ch] = it;
    }
}

// ---------------------------------------------------------------------------------------------
// THE HOT KERNEL.  Persistent, warp-specialised:
//   warp 0 (one lane): TMA producer  — per K block: A tile = 128 list rows x 32 floats (16 KB, EVICT_FIRST: streamed
//                      once), B tile = up to 64 gathered queries x 32 floats (8 KB, EVICT_LAST: re-read per tile)
//   warp 1 (one lane): tcgen05.mma issuer — 4 x (M=128, N=16..64, K=8) TF32 MMAs per K block into TMEM
//   warp 2           : TMEM allocation (2 accumulator buffers x 64 columns)
//   warps 4-7        : epilogue — tcgen05.ld the 128 x N scores, add ||x||^2, sample / capture
// Algorithmic HBM bytes per item: rows x (d*4 + 4 + 8)  (vector, norm, id), read exactly once.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)TC_STAGES * TC_A_BYTES;
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_tau[2][TC_NQT];
  __shared__ int s_q[2][TC_NQT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&s_tmem_base, 2 * TC_NQT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int n_items = *p.n_items;
  const int kblocks = (p.d + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer ----------------
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const TcItem I = p.items[it];
        const long long base_row = p.list_off[I.list] + I.row_begin;
        int rows = I.row_end - I.row_begin;
        if (p.mode == 0) rows = min(rows, TC_SAMPLE);
        const int ntiles = (rows + TC_BM - 1) / TC_BM;
        for (int t = 0; t < ntiles; ++t)
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], TC_A_BYTES + TC_B_BYTES);
            tma_load_2d(sA + (size_t)stage * TC_A_BYTES, &tmA, &full_bar[stage], kb * TC_BK, (int)(base_row + (long long)t * TC_BM), kEvictFirst);
            tma_load_2d(sB + (size_t)stage * TC_B_BYTES, &tmB, &full_bar[stage], kb * TC_BK, I.pair_begin, kEvictLast);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer ----------------
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tcount = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const TcItem I = p.items[it];
        int rows = I.row_end - I.row_begin;
        if (p.mode == 0) rows = min(rows, TC_SAMPLE);
        const int ntiles = (rows + TC_BM - 1) / TC_BM;
        const uint32_t npad = (uint32_t)max(16, (I.nq + 15) & ~15);
        const uint32_t idesc = make_idesc_tf32(TC_BM, npad);
        for (int t = 0; t < ntiles; ++t, ++tcount) {
          const uint32_t acc = tcount & 1, aphase = (tcount >> 1) & 1;
          mbar_wait(&tempty_bar[acc], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TC_NQT;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t adesc = make_desc_k128(smem_u32(sA + (size_t)stage * TC_A_BYTES));
            const uint64_t bdesc = make_desc_k128(smem_u32(sB + (size_t)stage * TC_B_BYTES));
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)  // UMMA K = 8 TF32 = 32 B: advance the start address by 2 (>>4 units)
              umma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: 4 warps, warp ew owns TMEM lanes [32*ew, 32*ew+32) ----------------
    const int ew = warp - 4;
    const int et = threadIdx.x - 128;
    uint32_t tcount = 0;
    int icount = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++icount) {
      const TcItem I = p.items[it];
      const int buf = icount & 1;
      if (et < I.nq) {
        const int q = p.pair_query[I.pair_begin + et];
        s_q[buf][et] = q;
        s_tau[buf][et] = p.mode ? p.tau[q] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long list_base = p.list_off[I.list];
      int rows = I.row_end - I.row_begin;
      if (p.mode == 0) rows = min(rows, TC_SAMPLE);
      const int ntiles = (rows + TC_BM - 1) / TC_BM;
      const int npad = max(16, (I.nq + 15) & ~15);
      for (int t = 0; t < ntiles; ++t, ++tcount) {
        const uint32_t acc = tcount & 1, aphase = (tcount >> 1) & 1;
        const int r = t * TC_BM + ew * 32 + lane;  // row within the item
        bool valid = r < rows;
        const long long arow = list_base + I.row_begin + r;
        float nrm = 0.f;
        if (valid) {
          const long long id = p.ids[arow];
          valid = id >= 0 && filter_pass(p.filt, id);
          if (valid && p.l2) nrm = p.norms[arow];
        }
        mbar_wait(&tfull_bar[acc], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * TC_NQT;
        for (int c0 = 0; c0 < npad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = c0 + j;
            if (n < I.nq) {
              const float dot = __uint_as_float(v[j]);
              const float score = p.l2 ? fmaf(-2.f, dot, nrm) : -dot;
              if (p.mode == 0) {
                if (ew == 0 && t == 0) p.sample[((size_t)it * TC_NQT + n) * TC_SAMPLE + lane] = valid ? score : __int_as_float(0x7f800000);
              } else if (valid && score <= s_tau[buf][n]) {
                const int q = s_q[buf][n];
                const int slot = atomicAdd(p.cand_cnt + q, 1);
                if (slot < p.cap) p.cand[(size_t)q * p.cap + slot] = ((unsigned long long)f2ord(score) << 32) | (unsigned long long)(uint32_t)arow;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * TC_NQT); }
}

// ---------------------------------------------------------------------------------------------
// per-query capture threshold: the k-th smallest sampled score (an upper bound of the k-th smallest score overall)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) tc_tau_kernel(const long long* __restrict__ probes, const int* __restrict__ pos,
                                                              const int* __restrict__ pair_of, const int* __restrict__ item_off,
                                                              const int* __restrict__ list_len, const int* __restrict__ pair_off,
                                                              int nprobe, const float* __restrict__ sample, int k, int pool_cap,
                                                              float* tau) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int q = blockIdx.x;
  BlockSelect sel;
  sel.init(smem, pool_cap, k);
  // enumerate (probe j, chunk c, sample row s); one thread per sample, SCAN_THREADS at a time
  for (int j = 0; j < nprobe; ++j) {
    const long long l = probes[(size_t)q * nprobe + j];
    if (l < 0) continue;  // uniform across the block
    const int len = list_len[l];
    if (len <= 0) continue;
    const int nc = (len + TC_CHUNK - 1) / TC_CHUNK;
    const int ps = pos[(size_t)q * nprobe + j];
    const int g = ps / TC_NQT, n = ps % TC_NQT;
    const int item0 = item_off[l] + g * nc;
    const int tot = nc * TC_SAMPLE;
    for (int base = 0; base < tot; base += blockDim.x) {
      sel.maybe_prune(blockDim.x);
      const int i = base + threadIdx.x;
      if (i < tot) {
        const int c = i / TC_SAMPLE, s = i % TC_SAMPLE;
        const float v = sample[((size_t)(item0 + c) * TC_NQT + n) * TC_SAMPLE + s];
        if (v < __int_as_float(0x7f800000)) {
          const uint32_t key = f2ord(v);
          const long long uid = ((long long)(item0 + c) << 8) | s;
          if (sel.passes(key, uid)) sel.push(key, uid);
        }
      }
    }
  }
  sel.prune();
  if (threadIdx.x == 0) tau[q] = (*sel.count >= k) ? ord2f(sel.kd[k - 1]) : __int_as_float(0x7f800000);
}

// ---------------------------------------------------------------------------------------------
// per query: window select on the approximate scores, exact rerank, certification
// ---------------------------------------------------------------------------------------------
template <bool L2>
__global__ void __launch_bounds__(SCAN_THREADS)
tc_final_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int cap,
                const float* __restrict__ tau, const float* __restrict__ qnorm, float max_norm, const float* __restrict__ q,
                const float* __restrict__ vecs, const long long* __restrict__ ids, int d, int k, int pool_cap,
                float* out_dist, long long* out_ids, int* flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int qi = blockIdx.x;
  float* qs = reinterpret_cast<float*>(smem);
  const size_t qbytes = ((size_t)d * 4 + 15) / 16 * 16;
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[(size_t)qi * d + i];
  const int total = cand_cnt[qi];
  const int n = min(total, cap);
  const unsigned long long* c = cand + (size_t)qi * cap;
  BlockSelect sel;
  sel.init(smem + qbytes, pool_cap, k);
  // pass 1: k-th smallest approximate score among the captured rows
  for (int base = 0; base < n; base += blockDim.x) {
    sel.maybe_prune(blockDim.x);
    const int i = base + threadIdx.x;
    if (i < n) {
      const unsigned long long e = c[i];
      const uint32_t key = (uint32_t)(e >> 32);
      const long long row = (long long)(e & 0xffffffffull);
      if (sel.passes(key, row)) sel.push(key, row);
    }
  }
  sel.prune();
  const int have1 = *sel.count;
  const float inf = __int_as_float(0x7f800000);
  const float a_k = have1 >= k ? ord2f(sel.kd[k - 1]) : inf;
  __syncthreads();
  // rigorous bound on |approx score - exact score| (TF32 truncation of the rows, RN rounding of the query,
  // FP32 accumulation), see DESIGN.md
  const float qn = sqrtf(qnorm[qi]);
  const float eps = (L2 ? 2.f : 1.f) * 0.001953125f /*2^-9*/ * qn * max_norm + (float)(d + 64) * 1.1920929e-7f * (qn + max_norm) * (qn + max_norm);
  const float window = a_k + 2.f * eps;  // +inf when fewer than k rows were captured
  const float tq = tau[qi];
  const bool certified = (total <= cap) && (tq == inf || window <= tq);
  // pass 2: exact distances (reference AVX-512 order) of every captured row inside the window
  sel.init(smem + qbytes, pool_cap, k);
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const bool vec = (d & 3) == 0;
  for (int base = 0; base < n; base += SCAN_QUADS) {
    sel.maybe_prune(SCAN_QUADS);
    const int i = base + quad;
    bool valid = i < n;
    long long row = 0;
    if (valid) {
      const unsigned long long e = c[i];
      valid = ord2f((uint32_t)(e >> 32)) <= window;
      row = (long long)(e & 0xffffffffull);
    }
    if (__ballot_sync(0xffffffffu, valid) == 0u) continue;
    if (!valid) row = (long long)(c[0] & 0xffffffffull);
    const float v = quad_distance<L2>(vecs + (size_t)row * d, qs, d, t, vec);
    if (valid && t == 0) {
      const uint32_t key = f2ord(L2 ? v : -v);
      const long long id = ids[row];
      if (sel.passes(key, id)) sel.push(key, id);
    }
  }
  sel.prune();
  const int have = *sel.count;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float api = 0.f;
    long long id = -1;
    if (i < have) {
      const float v = ord2f(sel.kd[i]);
      const float raw = L2 ? v : -v;
      api = L2 ? raw : __fsub_rn(1.0f, raw);
      id = sel.kid[i];
    }
    out_dist[(size_t)qi * k + i] = api;
    out_ids[(size_t)qi * k + i] = id;
  }
  if (threadIdx.x == 0) flags[qi] = certified ? 0 : 1;
}

__global__ void tc_compact_flags_kernel(const int* __restrict__ flags, int nq, int* qmap, int* count) {
  // single block, order-preserving
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int base = 0; base < nq; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int f = (i < nq) ? flags[i] : 0;
    if (f) qmap[atomicAdd(&s_cnt, 1)] = i;
  }
  __syncthreads();
  if (threadIdx.x == 0) *count = s_cnt;
}

__global__ void max_norm_kernel(const float* __restrict__ n2, long long n, float* out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, n2[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));  // non-negative floats order as ints
}

float device_max_norm(IndexBase* ix, const float* norms_sq, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0.f;
  float* d = ix->scratch.alloc<float>(1);
  B200VS_CUDA(cudaMemsetAsync(d, 0, 4, s));
  max_norm_kernel<<<(unsigned)std::min<int64_t>(1024, cdiv(n, 256)), 256, 0, s>>>(norms_sq, n, d);
  float h = 0.f;
  B200VS_CUDA(cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, s));
  B200VS_CUDA(cudaStreamSynchronize(s));
  return std::sqrt(h);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B200VS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
  if (!p || qres != cudaDriverEntryPointSuccess) fail(B200VS_EINTERNAL, "cuTensorMapEncodeTiled not available");
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// row-major float matrix [rows, d]; box = 32 floats (128 B) x box_rows, 128-byte swizzle
static CUtensorMap make_tmap(const float* base, int64_t rows, int d, int box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)std::max<int64_t>(rows, 1)};
  cuuint64_t gstride[1] = {(cuuint64_t)d * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(B200VS_EINTERNAL, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return m;
}

static int64_t tc_item_bound(const TcView& v, int64_t npairs) {
  return (npairs / TC_NQT + 1) * (int64_t)std::max(1, v.max_chunks_per_list) + v.total_chunks;
}
static int tc_cand_cap(int k) { return std::min(16384, std::max(2048, next_pow2(128 * k))); }

bool tc_eligible(const IndexBase* ix, const TcView& v, int64_t nq, int k, int nprobe, const SearchCtx& sc) {
  if (sc.exact_only) return false;
  if (ix->dim % 4 != 0 || ix->dim < 32) return false;      // TMA: 16-byte row pitch; tiny d is not worth a tile
  if (k > 128 || nq < 16) return false;                    // wide k / tiny batches: exact scan
  if (v.arena_rows <= 0 || v.arena_rows >= (1LL << 31)) return false;
  if (!v.norms || !v.vecs) return false;
  const int64_t npairs = nq * nprobe;
  if (npairs >= (1LL << 30)) return false;
  const int64_t bound = tc_item_bound(v, npairs);
  if (bound * TC_NQT * TC_SAMPLE * 4 > (2LL << 30)) return false;  // sample buffer too large
  return true;
}

void run_scan_mapped(IndexBase* ix, const ScanJob& job, int64_t nq_max, const int* qmap, const int* qcount, const float* queries,
                     int k, float* out_dist, long long* out_ids, cudaStream_t s);

void tc_search(IndexBase* ix, const TcView& v, bool l2, int64_t nq, const float* q, int k, const long long* probes,
               int nprobe, const SearchCtx& sc, float* out_dist, long long* out_ids, cudaStream_t s) {
  const int d = ix->dim;
  const int64_t npairs = nq * nprobe;
  const int64_t bound = tc_item_bound(v, npairs);
  const int cap = tc_cand_cap(k);
  Scratch& S = ix->scratch;
  float* q32 = S.alloc<float>((size_t)nq * d);
  float* qnorm = S.alloc<float>(nq);
  int* cnt = S.alloc<int>(v.nlist);
  int* pos = S.alloc<int>(npairs);
  int* pair_off = S.alloc<int>(v.nlist);
  int* item_off = S.alloc<int>(v.nlist);
  int* totals = S.alloc<int>(2);
  int* pair_query = S.alloc<int>(npairs);
  int* pair_of = S.alloc<int>(npairs);
  float* bws = S.alloc<float>((size_t)(npairs + TC_NQT) * d);
  TcItem* items = S.alloc<TcItem>(bound);
  float* sample = S.alloc<float>((size_t)bound * TC_NQT * TC_SAMPLE);
  float* tau = S.alloc<float>(nq);
  unsigned long long* cand = S.alloc<unsigned long long>((size_t)nq * cap);
  int* cand_cnt = S.alloc<int>(nq);
  int* flags = S.alloc<int>(nq);
  int* qmap = S.alloc<int>(nq);
  int* qcount = S.alloc<int>(1);

  B200VS_CUDA(cudaMemsetAsync(cnt, 0, (size_t)v.nlist * 4, s));
  B200VS_CUDA(cudaMemsetAsync(cand_cnt, 0, (size_t)nq * 4, s));
  tc_prep_queries_kernel<<<(unsigned)cdiv(nq * 32, 256), 256, 0, s>>>(q, nq, d, q32, qnorm);
  tc_count_pairs_kernel<<<(unsigned)cdiv(npairs, 256), 256, 0, s>>>(probes, npairs, cnt, pos);
  tc_plan_kernel<<<1, 1024, 0, s>>>(cnt, v.list_len, v.nlist, pair_off, item_off, totals);
  tc_fill_pairs_kernel<<<(unsigned)cdiv(npairs * 32, 256), 256, 0, s>>>(probes, pos, pair_off, npairs, nprobe, d, q32, pair_query, pair_of, bws);
  tc_items_kernel<<<(unsigned)cdiv(v.nlist, 128), 128, 0, s>>>(cnt, v.list_len, pair_off, item_off, v.nlist, items);
  B200VS_CUDA(cudaGetLastError());

  const CUtensorMap tmA = make_tmap(v.vecs, v.arena_rows, d, TC_BM);
  const CUtensorMap tmB = make_tmap(bws, npairs + TC_NQT, d, TC_NQT);
  TcParams p;
  p.ids = v.ids; p.norms = v.norms; p.list_off = v.list_off; p.d = d; p.items = items; p.n_items = totals;
  p.pair_query = pair_query; p.sample = sample; p.tau = tau; p.cand = cand; p.cand_cnt = cand_cnt; p.cap = cap;
  p.l2 = l2 ? 1 : 0;
  p.filt.has_range = sc.has_range; p.filt.negate = sc.negate; p.filt.rmin = sc.rmin; p.filt.rmax = sc.rmax;
  p.filt.sorted_ids = sc.sorted_ids_dev; p.filt.n_ids = sc.n_ids;

  static int num_sms = 0;
  if (!num_sms) {
    B200VS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, ix->device));
    B200VS_CUDA(cudaFuncSetAttribute(tc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  }
  const int grid = (int)std::min<int64_t>(num_sms, std::max<int64_t>(1, bound));
  const int pool = select_pool_cap(k, SCAN_THREADS);
  const size_t sel_smem = BlockSelect::smem_bytes(pool);

  // 1) sample pass -> per-query capture thresholds
  p.mode = 0;
  tc_scan_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(tmA, tmB, p);
  B200VS_CUDA(cudaFuncSetAttribute(tc_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  tc_tau_kernel<<<(unsigned)nq, SCAN_THREADS, sel_smem, s>>>(probes, pos, pair_of, item_off, v.list_len, pair_off, nprobe, sample, k, pool, tau);
  // 2) full pass: stream every probed list chunk once, capture rows under the threshold
  p.mode = 1;
  {
    ScopedKernelTimer timer(ix, s, ix->profiling);
    tc_scan_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(tmA, tmB, p);
    timer.stop();
  }
  // 3) window select + exact rerank + certification
  const size_t fin_smem = ((size_t)d * 4 + 15) / 16 * 16 + sel_smem;
  if (l2) {
    B200VS_CUDA(cudaFuncSetAttribute(tc_final_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    tc_final_kernel<true><<<(unsigned)nq, SCAN_THREADS, fin_smem, s>>>(cand, cand_cnt, cap, tau, qnorm, v.max_norm, q, v.vecs, v.ids, d, k, pool, out_dist, out_ids, flags);
  } else {
    B200VS_CUDA(cudaFuncSetAttribute(tc_final_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    tc_final_kernel<false><<<(unsigned)nq, SCAN_THREADS, fin_smem, s>>>(cand, cand_cnt, cap, tau, qnorm, v.max_norm, q, v.vecs, v.ids, d, k, pool, out_dist, out_ids, flags);
  }
  tc_compact_flags_kernel<<<1, 1024, 0, s>>>(flags, (int)nq, qmap, qcount);
  B200VS_CUDA(cudaGetLastError());
  ix->launch_count(10);
  ix->stats[1] = nq;
  // 4) uncertified queries re-run on the exact scan (device-side count: blocks beyond it exit immediately)
  ScanJob job;
  job.l2 = l2; job.vecs = v.vecs; job.ids = v.ids; job.d = d; job.sc = &sc;
  if (!v.flat) { job.mode = 1; job.probes = probes; job.nprobe = nprobe; job.list_off = v.list_off; job.list_len = v.list_len; }
  else { job.mode = 0; job.n = v.arena_rows; }
  run_scan_mapped(ix, job, nq, qmap, qcount, q, k, out_dist, out_ids, s);
  if (ix->profiling) {
    int h = 0;
    B200VS_CUDA(cudaMemcpyAsync(&h, qcount, 4, cudaMemcpyDeviceToHost, s));
    B200VS_CUDA(cudaStreamSynchronize(s));
    ix->stats[2] = h;
  }
}

}  // namespace b200vs
