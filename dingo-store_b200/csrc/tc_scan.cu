// tc_scan.cu — tensor-core candidate pass (see tc_scan.cuh for the contract, DESIGN.md for the proof sketch).
//
// Kernels:
//   tc_prep_queries_kernel   round queries to TF32 (RN), ||q||^2
//   tc_count_pairs_kernel    invert the probe table: how many queries probe each list
//   tc_plan_kernel           exclusive scans -> pair offsets, work-item offsets
//   tc_fill_pairs_kernel     gather the (TF32-rounded) queries of every list into one contiguous block
//   tc_items_kernel          emit work items (list chunk x query group) + the list of sampled items
//   tc_scan_kernel           *** the hot kernel: TMA -> smem ring -> tcgen05.mma (TF32) -> TMEM -> epilogue ***
//   tc_tau_kernel            per-query capture threshold from the sampled rows (radix select)
//   tc_final_fast_kernel     per-query window select + exact FP32 (AVX-512 order) rerank + certification
//                            (tc_final_kernel: the streaming variant for candidate buffers that do not fit shared memory)
//   tc_coarse_select_kernel  coarse-quantiser finish: register-resident radix select + exact rerank of the window
//                            (tc_coarse_final_fast_kernel / tc_coarse_final_kernel: one-sort / streaming fallbacks)
//   tc_compact_flags_kernel  list of uncertified queries for the exact re-run
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <mutex>

#include "scan_kernels.cuh"
#include "sm100_ptx.cuh"
#include "tc_scan.cuh"
#include "warp_select.cuh"

namespace b200vs {

using namespace ptx;

#define TC_INF __int_as_float(0x7f800000)

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
static __global__ void tc_prep_queries_kernel(const float* __restrict__ q, long long nq, int d, float* __restrict__ q32,
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

// error-compensated operands: hi = TF32 (low 13 mantissa bits cleared / rounded), lo = remainder
static __global__ void tc_prep_queries_split_kernel(const float* __restrict__ q, long long nq, int d, float* __restrict__ qhi,
                                                    float* __restrict__ qlo, float* __restrict__ qnorm) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nq) return;
  float acc = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = q[(size_t)w * d + i];
    acc = fmaf(v, v, acc);
    uint32_t r, r2;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    const float hi = __uint_as_float(r);
    const float lo = __fsub_rn(v, hi);  // exact
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r2) : "f"(lo));
    qhi[(size_t)w * d + i] = hi;
    qlo[(size_t)w * d + i] = __uint_as_float(r2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) qnorm[w] = acc;
}
static __global__ void tc_split_rows_kernel(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = __fsub_rn(v, h);  // exact; the tensor core truncates it to TF32 (error <= 2^-20 |x|)
}
void launch_split_rows(const float* x, int64_t n, int d, float* hi, float* lo, cudaStream_t s) {
  const long long tot = (long long)n * d;
  if (tot <= 0) return;
  tc_split_rows_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, s>>>(x, tot, hi, lo);
  B200VS_CUDA(cudaGetLastError());
}
// coarse pass: items = (row chunk) x (query group); B rows are the query rows themselves
static __global__ void tc_coarse_items_kernel(int nrows, int nq, int chunk, TcItem* items, int* totals) {
  const int nc = (nrows + chunk - 1) / chunk, ng = (nq + TC_NQT - 1) / TC_NQT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { totals[0] = nc * ng; totals[1] = nq; totals[2] = 0; }
  if (i >= nc * ng) return;
  const int ch = i / ng, g = i % ng;
  TcItem it;
  it.list = 0; it.row_begin = ch * chunk; it.row_end = min(nrows, (ch + 1) * chunk);
  it.pair_begin = g * TC_NQT; it.nq = min(TC_NQT, nq - g * TC_NQT); it.sample_slot = -1; it.pad[0] = it.pad[1] = 0;
  items[i] = it;
}

static __global__ void tc_count_pairs_kernel(const long long* __restrict__ probes, long long n, int* cnt, int* pos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long l = probes[i];
  pos[i] = l >= 0 ? atomicAdd(cnt + l, 1) : -1;
}

__device__ __forceinline__ int tc_items_of(int c, int len) {
  return (c > 0 && len > 0) ? ((c + TC_NQT - 1) / TC_NQT) * ((len + TC_CHUNK - 1) / TC_CHUNK) : 0;
}

// single block: exclusive scans over the lists
static __global__ void tc_plan_kernel(const int* __restrict__ cnt, const int* __restrict__ list_len, int nlist, int* pair_off,
                                      int* item_off, int* totals /*[0]=n_items [1]=n_pairs [2]=n_sample (zeroed here)*/) {
  __shared__ int s_pairs[1024], s_items[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (nlist + T - 1) / T;
  const int b = t * per, e = min(nlist, b + per);
  int sp = 0, si = 0;
  for (int l = b; l < e; ++l) { sp += cnt[l]; si += tc_items_of(cnt[l], list_len[l]); }
  // block-wide exclusive scan of (sp, si): warp shuffles + one scan of the 32 warp totals
  const int lane = t & 31, wid = t >> 5;
  int ip = sp, ii = si;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int vp = __shfl_up_sync(0xffffffffu, ip, o), vi = __shfl_up_sync(0xffffffffu, ii, o);
    if (lane >= o) { ip += vp; ii += vi; }
  }
  if (lane == 31) { s_pairs[wid] = ip; s_items[wid] = ii; }
  __syncthreads();
  if (wid == 0) {
    int wp = lane < (T >> 5) ? s_pairs[lane] : 0, wi = lane < (T >> 5) ? s_items[lane] : 0;
    const int tp = wp, ti = wi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vp = __shfl_up_sync(0xffffffffu, wp, o), vi = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) { wp += vp; wi += vi; }
    }
    if (lane == 31) { totals[0] = wi; totals[1] = wp; totals[2] = 0; totals[3] = 0; }
    s_pairs[lane] = wp - tp; s_items[lane] = wi - ti;  // exclusive warp offsets
  }
  __syncthreads();
  sp = s_pairs[wid] + ip - sp; si = s_items[wid] + ii - si;  // exclusive prefix of this thread
  for (int l = b; l < e; ++l) {
    pair_off[l] = sp; item_off[l] = si;
    sp += cnt[l]; si += tc_items_of(cnt[l], list_len[l]);
  }
}

// one warp per (query, probe): place the pair and copy the rounded query row
static __global__ void tc_fill_pairs_kernel(const long long* __restrict__ probes, const int* __restrict__ pos,
                                            const int* __restrict__ pair_off, const int* __restrict__ list_len, long long n,
                                            int nprobe, int d, const float* __restrict__ q32, int* pair_query, float* bws) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long l = probes[w];
  if (l < 0 || list_len[l] <= 0) return;  // empty list (e.g. owned by another shard): no work item will read this row
  const int p = pair_off[l] + pos[w];
  const int q = (int)(w / nprobe);
  if (lane == 0) pair_query[p] = q;
  if (!bws) return;  // gather4 mode: the scan kernel reads the query rows directly
  const float4* src = reinterpret_cast<const float4*>(q32 + (size_t)q * d);
  float4* dst = reinterpret_cast<float4*>(bws + (size_t)p * d);
  for (int i = lane; i < (d >> 2); i += 32) dst[i] = src[i];
}

// items of list l live at item_off[l] + chunk * ngroups + group  (the groups of one chunk are adjacent: the
// second group re-reads the chunk from L2 while it is hot)
static __global__ void tc_items_kernel(const int* __restrict__ cnt, const int* __restrict__ list_len,
                                       const int* __restrict__ pair_off, const int* __restrict__ item_off, int nlist, TcItem* items,
                                       int* totals, int* sample_list) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlist) return;
  const int c = cnt[l], len = list_len[l];
  if (c <= 0 || len <= 0) return;
  const int ng = (c + TC_NQT - 1) / TC_NQT, nc = (len + TC_CHUNK - 1) / TC_CHUNK;
  const int base = item_off[l];
  for (int ch = 0; ch < nc; ++ch)
    for (int g = 0; g < ng; ++g) {
      TcItem it;
      it.list = l;
      it.row_begin = ch * TC_CHUNK;
      it.row_end = min(len, (ch + 1) * TC_CHUNK);
      it.pair_begin = pair_off[l] + g * TC_NQT;
      it.nq = min(TC_NQT, c - g * TC_NQT);
      it.sample_slot = -1;
      if ((it.row_begin % TC_SPAN) == 0) {
        it.sample_slot = atomicAdd(totals + 2, 1);
        sample_list[it.sample_slot] = base + ch * ng + g;
      }
      it.pad[0] = it.pad[1] = 0;
      items[base + ch * ng + g] = it;
      // profiling aid: tensor-core work = 128-row tiles x padded query columns
      atomicAdd(totals + 3, ((it.row_end - it.row_begin + TC_BM - 1) / TC_BM) * (it.nq <= 16 ? 16 : (it.nq <= 32 ? 32 : (it.nq <= 64 ? 64 : 128))));
    }
}

// ---------------------------------------------------------------------------------------------
// THE HOT KERNEL.  Persistent (one CTA per SM), warp-specialised, dynamically scheduled:
//   warp 0 (one lane): scheduler + TMA producer — claims the next work item with one atomicAdd, publishes it through
//                      a 4-deep smem queue, then per K block loads the A tile = 128 list rows x 32 floats (16 KB,
//                      EVICT_FIRST: streamed once) and the B tile = 16/32/64/128 gathered queries x 32 floats (EVICT_LAST)
//   warp 1 (one lane): tcgen05.mma issuer — 4 x (M=128, N=16..128, K=8) TF32 MMAs per K block into TMEM
//   warp 2           : TMEM allocation (2 accumulator buffers x 128 columns)
//   warps 4-7        : epilogue — tcgen05.ld the 128 x N scores, add ||x||^2, then sample / capture (staged in shared
//                      memory, flushed per item) / dense store
// Algorithmic HBM bytes per item: rows x (d*4 + 4 + 8)  (vector, norm, id), read exactly once.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA32,
               const __grid_constant__ CUtensorMap tmB16, const __grid_constant__ CUtensorMap tmB32,
               const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmB128,
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo16,
               const __grid_constant__ CUtensorMap tmBlo32, const __grid_constant__ CUtensorMap tmBlo64,
               const __grid_constant__ CUtensorMap tmBlo128, const __grid_constant__ CUtensorMap tmQ, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)TC_STAGES * TC_A_BYTES;
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ __align__(8) uint64_t sched_full[TC_SQ], sched_empty[TC_SQ];
  __shared__ int s_sched[TC_SQ];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_tau[2][TC_NQT];
  __shared__ int s_q[2][TC_NQT];
  __shared__ int s_brow[TC_NQT];  // producer-private: query rows of the current item (gather4 mode)
  // capture staging, one buffer per epilogue warp: hits are appended with a shared-memory atomic and written out
  // (one global atomicAdd per hit, a whole warp of them in flight) when the item ends, so the score loop never
  // waits on a global round trip
  __shared__ unsigned long long s_hit_key[4][TC_HITS];
  __shared__ int s_hit_q[4][TC_HITS];
  __shared__ int s_hit_n[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 4) s_hit_n[threadIdx.x] = 0;
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmA32); prefetch_tmap(&tmB16); prefetch_tmap(&tmB32); prefetch_tmap(&tmB64); prefetch_tmap(&tmB128); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    for (int i = 0; i < TC_SQ; ++i) { mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], 5); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&s_tmem_base, 2 * TC_NQT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int nwork = p.mode == 0 ? p.totals[2] : p.totals[0];
  const int kblocks = (p.d + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- scheduler + TMA producer ----------------
      int stage = 0, slot = 0;
      uint32_t phase = 0, sphase = 0;
      for (;;) {
        mbar_wait(&sched_empty[slot], sphase ^ 1);
        const int w = atomicAdd(p.work_counter, 1);
        const int it = w < nwork ? (p.mode == 0 ? p.sample_list[w] : w) : -1;
        s_sched[slot] = it;
        mbar_arrive(&sched_full[slot]);
        if (++slot == TC_SQ) { slot = 0; sphase ^= 1; }
        if (it < 0) break;
        const TcItem I = p.items[it];
        const long long base_row = p.list_off[I.list] + I.row_begin;
        int rows = I.row_end - I.row_begin;
        if (p.mode == 0) rows = min(rows, p.sample_rows);
        const int ntiles = (rows + TC_BM - 1) / TC_BM;
        const int b_rows = I.nq <= 16 ? 16 : (I.nq <= 32 ? 32 : (I.nq <= 64 ? 64 : 128));
        const CUtensorMap* tb = I.nq <= 16 ? &tmB16 : (I.nq <= 32 ? &tmB32 : (I.nq <= 64 ? &tmB64 : &tmB128));
        const bool small_a = p.mode == 0 && p.sample_rows == TC_SAMPLE;  // 32-row sample box, else the full 128-row tile
        const CUtensorMap* ta = small_a ? &tmA32 : &tmA;
        const uint32_t bytes = (small_a ? (uint32_t)(TC_SAMPLE * 128) : TC_A_BYTES) + (uint32_t)b_rows * 128u;
        const CUtensorMap* tbl = I.nq <= 16 ? &tmBlo16 : (I.nq <= 32 ? &tmBlo32 : (I.nq <= 64 ? &tmBlo64 : &tmBlo128));
        const int npad_b = max(16, (I.nq + 15) & ~15);
        uint32_t bytes_g = bytes;
        if (p.b_gather) {
          for (int j = 0; j < npad_b; ++j) s_brow[j] = j < I.nq ? p.pair_query[I.pair_begin + j] : 0;
          bytes_g = (small_a ? (uint32_t)(TC_SAMPLE * 128) : TC_A_BYTES) + (uint32_t)npad_b * 128u;
        }
        for (int t = 0; t < ntiles; ++t)
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            // last tile of a list: fetch only the 32-row boxes that hold list rows (the rest of the stage keeps stale
            // data, whose scores the epilogue masks) instead of streaming up to 127 rows of the neighbouring list
            const int trows = rows - t * TC_BM;
            const int nbox = (p.mode == 1 && !p.split && trows <= TC_BM - TC_SAMPLE) ? (trows + TC_SAMPLE - 1) / TC_SAMPLE : 0;
            const uint32_t saved = nbox ? TC_A_BYTES - (uint32_t)nbox * (TC_SAMPLE * 128) : 0u;
            mbar_arrive_expect_tx(&full_bar[stage], (p.b_gather ? bytes_g : bytes) - saved);
            if (nbox) {
              for (int j = 0; j < nbox; ++j)
                tma_load_2d(sA + (size_t)stage * TC_A_BYTES + (size_t)j * (TC_SAMPLE * 128), &tmA32, &full_bar[stage], kb * TC_BK,
                            (int)(base_row + (long long)t * TC_BM + j * TC_SAMPLE), kEvictFirst);
            } else
            tma_load_2d(sA + (size_t)stage * TC_A_BYTES, ta, &full_bar[stage], kb * TC_BK, (int)(base_row + (long long)t * TC_BM), kEvictFirst);
            if (p.b_gather) {
              uint8_t* bdst = sB + (size_t)stage * TC_B_BYTES;
              for (int g = 0; g < npad_b; g += 4)
                tma_gather4_2d(bdst + (size_t)g * 128, &tmQ, &full_bar[stage], kb * TC_BK, s_brow[g], s_brow[g + 1], s_brow[g + 2], s_brow[g + 3], kEvictLast);
            } else
            tma_load_2d(sB + (size_t)stage * TC_B_BYTES, tb, &full_bar[stage], kb * TC_BK, I.pair_begin, kEvictLast);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            if (p.split) {  // error-compensated pass: the "lo" operands ride in the next ring stage
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], bytes);
              tma_load_2d(sA + (size_t)stage * TC_A_BYTES, &tmAlo, &full_bar[stage], kb * TC_BK, (int)(base_row + (long long)t * TC_BM), kEvictFirst);
              tma_load_2d(sB + (size_t)stage * TC_B_BYTES, tbl, &full_bar[stage], kb * TC_BK, I.pair_begin, kEvictLast);
              if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer ----------------
      int stage = 0, slot = 0;
      uint32_t phase = 0, sphase = 0, tcount = 0;
      for (;;) {
        mbar_wait(&sched_full[slot], sphase);
        const int it = s_sched[slot];
        mbar_arrive(&sched_empty[slot]);
        if (++slot == TC_SQ) { slot = 0; sphase ^= 1; }
        if (it < 0) break;
        const TcItem I = p.items[it];
        int rows = I.row_end - I.row_begin;
        if (p.mode == 0) rows = min(rows, p.sample_rows);
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
            const int stage_hi = stage;
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            if (p.split) {  // + lo*hi + hi*lo into the same accumulator
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint64_t adesc_lo = make_desc_k128(smem_u32(sA + (size_t)stage * TC_A_BYTES));
              const uint64_t bdesc_lo = make_desc_k128(smem_u32(sB + (size_t)stage * TC_B_BYTES));
#pragma unroll
              for (int k = 0; k < TC_BK / 8; ++k) {
                umma_tf32(d_tmem, adesc_lo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                umma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc_lo + (uint64_t)(2 * k), idesc, 1u);
              }
              umma_commit(&empty_bar[stage_hi]);
              umma_commit(&empty_bar[stage]);
              if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            } else {
              umma_commit(&empty_bar[stage_hi]);  // frees the smem slot when these MMAs retire
            }
          }
          umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: 4 warps, warp ew owns TMEM lanes [32*ew, 32*ew+32) ----------------
    const int ew = warp - 4;
    const int et = threadIdx.x - 128;
    uint32_t tcount = 0, sphase = 0;
    int icount = 0, slot = 0;
    for (;; ++icount) {
      mbar_wait(&sched_full[slot], sphase);
      const int it = s_sched[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sched_empty[slot]);
      if (++slot == TC_SQ) { slot = 0; sphase ^= 1; }
      if (it < 0) break;
      const TcItem I = p.items[it];
      const int buf = icount & 1;
      if (et < I.nq) {
        const int q = p.pair_query ? p.pair_query[I.pair_begin + et] : I.pair_begin + et;
        s_q[buf][et] = q;
        s_tau[buf][et] = p.mode == 1 ? p.tau[q] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long list_base = p.list_off[I.list];
      int rows = I.row_end - I.row_begin;
      if (p.mode == 0) rows = min(rows, p.sample_rows);
      const int ntiles = (rows + TC_BM - 1) / TC_BM;
      const int npad = max(16, (I.nq + 15) & ~15);
      for (int t = 0; t < ntiles; ++t, ++tcount) {
        const uint32_t acc = tcount & 1, aphase = (tcount >> 1) & 1;
        const int r = t * TC_BM + ew * 32 + lane;  // row within the item
        const bool inrange = r < rows;
        bool valid = inrange;
        const long long arow = list_base + I.row_begin + r;
        float nrm = 0.f;
        if (valid) {
          const long long id = p.ids[arow];
          valid = id >= 0 && filter_pass(p.filt, id);
          if (valid && p.l2 && p.add_norm) nrm = p.norms[arow];
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
              if (p.mode == 1) {
                if (valid && score <= s_tau[buf][n]) {
                  const int q = s_q[buf][n];
                  const unsigned long long key = ((unsigned long long)f2ord(score) << 32) | (unsigned long long)(uint32_t)arow;
                  const int h = atomicAdd(&s_hit_n[ew], 1);
                  if (h < TC_HITS) { s_hit_key[ew][h] = key; s_hit_q[ew][h] = q; }
                  else {  // staging full: write through
                    const int sl = atomicAdd(p.cand_cnt + q, 1);
                    if (sl < p.cap) p.cand[(size_t)q * p.cap + sl] = key;
                  }
                }
              } else if (p.mode == 0) {
                if (t == 0 && r < p.sample_rows) p.sample[((size_t)I.sample_slot * TC_NQT + n) * p.sample_rows + r] = valid ? score : TC_INF;
              } else if (inrange) {
                float* dst = p.dense + (size_t)s_q[buf][n] * p.dense_ld + (I.row_begin + r);
                if (p.dense_accum) *dst = *dst + score;
                else *dst = valid ? score : TC_INF;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if (p.mode == 1) {  // flush this warp's staged hits
        __syncwarp();
        const int nh = min(s_hit_n[ew], TC_HITS);
        for (int i = lane; i < nh; i += 32) {
          const int q = s_hit_q[ew][i];
          const int sl = atomicAdd(p.cand_cnt + q, 1);
          if (sl < p.cap) p.cand[(size_t)q * p.cap + sl] = s_hit_key[ew][i];
        }
        __syncwarp();
        if (lane == 0) s_hit_n[ew] = 0;
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * TC_NQT); }
}

// ---------------------------------------------------------------------------------------------
// per-query capture threshold T.  tau = the k-th smallest sampled score bounds the k-th smallest score overall (A_k)
// from above.  Thin samples (32 rows per span, ~1 % of the probed rows) leave tau far above A_k + 2 eps and T = tau:
// the finish kernel certifies A_k + 2 eps <= T and sends the rare failure to the exact scan.  Dense samples (the whole
// first tile, used when a query probes few local lists) put tau close to A_k, so T = tau + 2 eps: then
// A_k + 2 eps <= T by construction and certification can only fail on a buffer overflow.
// ---------------------------------------------------------------------------------------------
__device__ uint32_t block_kth_key(const unsigned long long* keys, int n, int k);
__device__ __forceinline__ float tc_eps(bool l2, float qnorm_sq, float max_norm, int d, bool split);

// Capture threshold from the sampled k-th score tau.  The finish kernel certifies A_k + 2 eps <= T (A_k = k-th captured
// score); a query that fails is re-run on the exact scan, and a single re-scanned query of a large shard stalls every rank
// of a sharded batch for milliseconds.  Two cases:
//   * at least TAU_SAFE sampled scores lie below tau - 2 eps: about TAU_SAFE * TC_SPAN / TC_SAMPLE rows do too, far more
//     than k, so A_k < tau - 2 eps almost surely (fewer than k such rows would have to produce TAU_SAFE sample hits at a
//     1 / 64 sampling rate: probability < 3e-4 even then) -> T = tau, the tightest capture set;
//   * otherwise the k-th row may sit within 2 eps of tau: capture under T = tau + 2 eps, which makes the certification true
//     by construction (A_k <= tau because the sample is a subset of the rows) at the price of a wider capture set — taken
//     only while the predicted number of captures leaves headroom in the capture buffer.
constexpr int TAU_SAFE = 3;
template <class CountLE>
__device__ __forceinline__ float tc_tau_with_margin(float tau, float two_eps, int cap, CountLE count_le) {
  if (count_le(__fsub_rd(tau, two_eps)) >= TAU_SAFE) return tau;
  const float lim = __fadd_ru(tau, two_eps);
  const float cs = (float)count_le(lim);  // sampled scores under T: the capture count is ~ (cs +- sqrt(cs)) * sampling ratio
  const float predicted = (cs + 4.f * sqrtf(cs)) * (float)(TC_SPAN / TC_SAMPLE);  // + 4 sigma: thousands of queries per batch
  return predicted <= (float)cap ? lim : tau;
}

constexpr int TAU_PL = 4096;  // (sample slot, column) pairs staged per query
constexpr int TAU_SORT = 4096;  // sampled scores sorted in one shot when they fit

static __global__ void __launch_bounds__(SCAN_THREADS)
tc_tau_kernel(const long long* __restrict__ probes, const int* __restrict__ pos, const int* __restrict__ cnt,
              const int* __restrict__ item_off, const int* __restrict__ list_len, const TcItem* __restrict__ items, int nprobe,
              const float* __restrict__ sample, int srows, int k, int pool_cap, int l2, const float* __restrict__ qnorm,
              float max_norm, int d, int margin, int cand_cap, float* tau, const int* redo) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_np;
  const int q = blockIdx.x;
  if (redo && !redo[q]) return;  // only the queries the warp kernel could not finish
  int2* s_pairs = reinterpret_cast<int2*>(smem);  // [TAU_PL]
  if (threadIdx.x == 0) s_np = 0;
  BlockSelect sel;
  sel.init(smem + (size_t)TAU_PL * 8, pool_cap, k);  // barrier inside
  for (int j = threadIdx.x; j < nprobe; j += blockDim.x) {
    const long long l = probes[(size_t)q * nprobe + j];
    if (l < 0) continue;
    const int len = list_len[l];
    if (len <= 0) continue;
    const int nc = (len + TC_CHUNK - 1) / TC_CHUNK, ng = (cnt[l] + TC_NQT - 1) / TC_NQT;
    const int ps = pos[(size_t)q * nprobe + j];
    const int g = ps / TC_NQT, n = ps % TC_NQT;
    const int nspan = (len + TC_SPAN - 1) / TC_SPAN;
    const int base = atomicAdd(&s_np, nspan);
    for (int sp = 0; sp < nspan; ++sp) {
      const int ch = sp * (TC_SPAN / TC_CHUNK);
      if (ch < nc && base + sp < TAU_PL) s_pairs[base + sp] = make_int2(items[item_off[l] + ch * ng + g].sample_slot, n);
    }
  }
  __syncthreads();
  const int np_all = s_np;
  const int np = min(np_all, TAU_PL);
  const int tot = np * srows;
  if (np_all <= TAU_PL && tot <= TAU_SORT) {  // common case: radix-select the k-th smallest sampled score in shared memory
    unsigned long long* vals = reinterpret_cast<unsigned long long*>(smem + (size_t)TAU_PL * 8);
    __shared__ int s_fin;
    if (threadIdx.x == 0) s_fin = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < tot; i += blockDim.x) {
      const int2 pr = s_pairs[i / srows];
      const float v = sample[((size_t)pr.x * TC_NQT + pr.y) * srows + (i % srows)];
      if (v < TC_INF) vals[atomicAdd(&s_fin, 1)] = (unsigned long long)f2ord(v) << 32;
    }
    __syncthreads();
    const int nfin = s_fin;
    const uint32_t kth = nfin >= k ? block_kth_key(vals, nfin, k) : 0u;
    if (nfin < k) { if (threadIdx.x == 0) tau[q] = TC_INF; return; }
    const float two_eps = 2.f * tc_eps(l2 != 0, qnorm[q], max_norm, d, false);
    if (margin) { if (threadIdx.x == 0) tau[q] = __fadd_rn(ord2f(kth), two_eps); return; }  // dense samples: always
    __shared__ int s_cle;
    auto count_le = [&](float lim) {  // block-wide count of sampled scores <= lim (all threads call)
      __syncthreads();
      if (threadIdx.x == 0) s_cle = 0;
      __syncthreads();
      int c = 0;
      for (int i = threadIdx.x; i < nfin; i += blockDim.x) c += ord2f((uint32_t)(vals[i] >> 32)) <= lim ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cle, c);
      __syncthreads();
      return s_cle;
    };
    const float t = tc_tau_with_margin(ord2f(kth), two_eps, cand_cap, count_le);
    if (threadIdx.x == 0) tau[q] = t;
    return;
  }
  for (int base = 0; base < tot; base += blockDim.x) {
    sel.maybe_prune(blockDim.x);
    const int i = base + threadIdx.x;
    if (i < tot) {
      const int2 pr = s_pairs[i / srows];
      const int s = i % srows;
      const float v = sample[((size_t)pr.x * TC_NQT + pr.y) * srows + s];
      if (v < TC_INF) {
        const uint32_t key = f2ord(v);
        if (sel.passes(key, i)) sel.push(key, i);
      }
    }
  }
  sel.prune();
  // more sampled spans than the staging area holds: capture everything (overflow then falls back to the exact scan)
  if (threadIdx.x == 0)
    tau[q] = (np_all <= TAU_PL && *sel.count >= k) ? __fadd_rn(ord2f(sel.kd[k - 1]), margin ? 2.f * tc_eps(l2 != 0, qnorm[q], max_norm, d, false) : 0.f) : TC_INF;
}


// Warp-per-query form of the threshold kernel (the common case: <= WT_PAIRS sampled spans of TC_SAMPLE rows per query).
// Lanes walk the probe list, stage one (sample slot, column) per sampled span, then the warp reads one span per step
// (32 contiguous floats) and radix-selects the k-th smallest finite score.  Queries with more spans set redo[q] and are
// finished by tc_tau_kernel.
constexpr int WT_WARPS = 4;
constexpr int WT_PAIRS = 64;
static __global__ void __launch_bounds__(WT_WARPS * 32)
tc_tau_warp_kernel(const long long* __restrict__ probes, const int* __restrict__ pos, const int* __restrict__ cnt,
                   const int* __restrict__ item_off, const int* __restrict__ list_len, const TcItem* __restrict__ items, int nprobe,
                   const float* __restrict__ sample, int k, int nq, int l2, const float* __restrict__ qnorm, float max_norm, int d, int cap,
                   float* tau, int* redo) {
  __shared__ int s_hist[WT_WARPS][WS_BINS];
  __shared__ int s_pair[WT_WARPS][WT_PAIRS];
  __shared__ uint32_t s_val[WT_WARPS][WT_PAIRS * TC_SAMPLE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * WT_WARPS + warp;
  if (q >= nq) return;
  int np = 0;
  for (int base = 0; base < nprobe; base += 32) {
    const int j = base + lane;
    int nspan = 0, first = 0, ng = 1, g = 0, n = 0;
    if (j < nprobe) {
      const long long l = probes[(size_t)q * nprobe + j];
      const int len = l >= 0 ? list_len[l] : 0;
      if (len > 0) {
        ng = (cnt[l] + TC_NQT - 1) / TC_NQT;
        const int ps = pos[(size_t)q * nprobe + j];
        g = ps / TC_NQT; n = ps % TC_NQT;
        nspan = (len + TC_SPAN - 1) / TC_SPAN;
        first = item_off[l];
      }
    }
    int incl = nspan;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int off = np + incl - nspan;
    for (int sp = 0; sp < nspan; ++sp)
      if (off + sp < WT_PAIRS) s_pair[warp][off + sp] = items[first + sp * (TC_SPAN / TC_CHUNK) * ng + g].sample_slot * TC_NQT + n;
    np += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (np > WT_PAIRS) { if (lane == 0) redo[q] = 1; return; }
  if (lane == 0) redo[q] = 0;
  __syncwarp();
  // one lane per sampled span: its 32 scores are one 128-byte line, fetched as 8 independent 16-byte loads, so the whole
  // query's sample is in flight at once (a warp-per-span loop would serialise one L2 round trip per span)
  __shared__ int s_nfin[WT_WARPS];
  if (lane == 0) s_nfin[warp] = 0;
  __syncwarp();
  for (int p = lane; p < np; p += 32) {
    const float4* src = reinterpret_cast<const float4*>(sample + (size_t)s_pair[warp][p] * TC_SAMPLE);
    float4 v[TC_SAMPLE / 4];
#pragma unroll
    for (int j = 0; j < TC_SAMPLE / 4; ++j) v[j] = src[j];
    const float* f = reinterpret_cast<const float*>(v);
    int c = 0;
#pragma unroll
    for (int j = 0; j < TC_SAMPLE; ++j) c += f[j] < TC_INF ? 1 : 0;
    int o = atomicAdd(&s_nfin[warp], c);
#pragma unroll
    for (int j = 0; j < TC_SAMPLE; ++j) if (f[j] < TC_INF) s_val[warp][o++] = f2ord(f[j]);
  }
  __syncwarp();
  const int nfin = s_nfin[warp];
  float t = TC_INF;
  if (nfin >= k) {
    int c_le;
    const uint32_t* vals = s_val[warp];
    t = ord2f(warp_kth_key(k, s_hist[warp], [&](auto f) { for (int i = lane; i < nfin; i += 32) f(vals[i]); }, c_le));
    t = tc_tau_with_margin(t, 2.f * tc_eps(l2 != 0, qnorm[q], max_norm, d, false), cap,
                           [&](float lim) { int c = 0; for (int i = lane; i < nfin; i += 32) c += ord2f(vals[i]) <= lim ? 1 : 0; return __reduce_add_sync(0xffffffffu, c); });
  }
  if (lane == 0) tau[q] = t;
}

// ---------------------------------------------------------------------------------------------
// per query: window select on the approximate scores, exact rerank, certification
// ---------------------------------------------------------------------------------------------
constexpr int FIN_SEG = 2048;  // in-window rows staged per round

__device__ __forceinline__ float tc_eps(bool l2, float qnorm_sq, float max_norm, int d, bool split) {
  // rigorous bound on |approx score - exact score|.  Operand term: single-pass TF32 = rows truncated
  // (< 2^-10 relative) + query rounded to nearest (<= 2^-11) = 1.5 * 2^-10 (the 2^-21 cross term and the FP32 accumulation
  // are covered by the second summand); split pass (hi*hi + lo*hi + hi*lo) -> 2^-19.
  // Accumulation term: FP32 sums of d terms on both sides (tensor core, norms, and the exact kernel itself).
  const float qn = sqrtf(qnorm_sq);
  const float operand = (l2 ? 2.f : 1.f) * (split ? 1.9073486e-6f /*2^-19*/ : 0.0014648438f /*2^-10 + 2^-11*/) * qn * max_norm;
  const float accum = 2.f * (float)(d + 64) * 5.9604645e-8f /*2^-24*/ * (l2 ? (max_norm + qn) * (max_norm + qn) : qn * max_norm);
  return operand + accum;
}

// ---- block-wide bitonic sorts in shared memory (ascending), m = power of two ----
__device__ __forceinline__ void block_sort_u64(unsigned long long* keys, int m) {
  for (int size = 2; size <= m; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
        const int pos = 2 * i - (i & (stride - 1)), j = pos + stride;
        const unsigned long long a = keys[pos], b = keys[j];
        if (((pos & size) == 0) ? (b < a) : (a < b)) { keys[pos] = b; keys[j] = a; }
      }
      __syncthreads();
    }
}
__device__ __forceinline__ void block_sort_pair(uint32_t* kd, long long* kid, int m) {
  for (int size = 2; size <= m; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
        const int pos = 2 * i - (i & (stride - 1)), j = pos + stride;
        const uint32_t ad = kd[pos], bd = kd[j];
        const long long ai = kid[pos], bi = kid[j];
        const bool swap = ((pos & size) == 0) ? key_less(bd, bi, ad, ai) : key_less(ad, ai, bd, bi);
        if (swap) { kd[pos] = bd; kd[j] = ad; kid[pos] = bi; kid[j] = ai; }
      }
      __syncthreads();
    }
}

// Fast window select shared by the list-scan finish and the coarse finish.  keys[0..n) = (ord(approx score) << 32 | row),
// unsorted, padded storage for the next power of two.  Sorts once, takes the k-th approximate score, re-scores the
// in-window prefix exactly (reference order) and returns the exact top-k in (ex_kd, ex_id)[0..k).
// Returns the number of exact entries kept (<= k) and whether the prefix fit in `maxw`.
// k-th smallest (k >= 1, n >= k) of the 32-bit keys stored in the high halves of keys[0..n): 4-pass radix select with a
// 256-bin shared histogram (no sorting network, ~10 barriers).  All threads call; returns the key to all threads.
__device__ uint32_t block_kth_key(const unsigned long long* keys, int n, int k) {
  __shared__ int s_hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_k;
  if (threadIdx.x == 0) { s_prefix = 0; s_k = k; }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = (uint32_t)(keys[i] >> 32);
      if (pass == 0 || ((key ^ prefix) >> (shift + 8)) == 0) atomicAdd(&s_hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // warp 0: locate the bin that holds the k-th element
      const int lane = threadIdx.x;
      int loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = s_hist[lane * 8 + j]; sum += loc[j]; }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const int kk = s_k;
      const int before = incl - sum;
      if (before < kk && kk <= incl) {  // exactly one lane
        int acc = before;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (acc < kk && kk <= acc + loc[j]) { s_prefix = prefix | ((uint32_t)(lane * 8 + j) << shift); s_k = kk - acc; }
          acc += loc[j];
        }
      }
    }
    __syncthreads();
  }
  return s_prefix;
}

// Window select shared by the list-scan finish and the coarse finish.  keys[0..n) = (ord(approx score) << 32 | row),
// unsorted.  Finds the k-th approximate score (radix select), gathers every row whose score is within two_eps of it,
// re-scores those rows exactly (reference order) and sorts the exact (key, id) pairs: ex_kd/ex_id[0..returned).
// Returns the number of exact entries kept (<= k); *fits = the window fit in `maxw` rows.
template <bool L2, bool ROW_IS_ID>
__device__ int window_select(const unsigned long long* keys, int n, int k, float two_eps, const float* qs, const float* __restrict__ vecs,
                             const long long* __restrict__ ids, int d, uint32_t* ex_kd, long long* ex_id, int* rows, int maxw, bool* fits,
                             float* a_k_out, int* s_m) {
  const float a_k = n >= k ? ord2f(block_kth_key(keys, n, k)) : TC_INF;
  const float window = a_k + two_eps;
  if (threadIdx.x == 0) *s_m = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long e = keys[i];
    if (ord2f((uint32_t)(e >> 32)) <= window) {
      const int pos = atomicAdd(s_m, 1);
      if (pos < maxw) rows[pos] = (int)(uint32_t)(e & 0xffffffffull);
    }
  }
  __syncthreads();
  int m = *s_m;
  *fits = m <= maxw;
  m = min(m, maxw);
  *a_k_out = a_k;
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const bool vec = (d & 3) == 0;
  {  // pull every in-window row into L2 at once (random 3 KB gathers: latency-bound otherwise)
    const int lines = (d * 4 + 127) / 128;
    for (int i = threadIdx.x; i < m * lines; i += blockDim.x) {
      const char* ptr = reinterpret_cast<const char*>(vecs + (size_t)(uint32_t)rows[i / lines] * d) + (size_t)(i % lines) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
  }
  const int nquads = blockDim.x >> 2;
  for (int base = 0; base < m; base += nquads) {
    const int i = base + quad;
    const bool valid = i < m;
    const long long row = (long long)(uint32_t)rows[valid ? i : 0];
    const float v = quad_distance<L2>(vecs + (size_t)row * d, qs, d, t, vec);
    if (valid && t == 0) { ex_kd[i] = f2ord(L2 ? v : -v); ex_id[i] = ROW_IS_ID ? row : ids[row]; }
  }
  int e2 = 2;
  while (e2 < m) e2 <<= 1;
  __syncthreads();
  for (int i = m + threadIdx.x; i < e2; i += blockDim.x) { ex_kd[i] = KEY_SENTINEL_D; ex_id[i] = KEY_SENTINEL_ID; }
  __syncthreads();
  block_sort_pair(ex_kd, ex_id, e2);
  return min(m, k);
}

constexpr int FIN_THREADS = 128;   // finish kernels: many small CTAs hide their barrier / gather latency better than few big ones
constexpr int FIN_MAXW = 1024;     // in-window rows re-scored per query on the fast path (more -> exact re-run)
constexpr int COARSE_FAST = 8192;  // coarse rows handled by the register-select / one-sort path

template <bool L2>
static __global__ void __launch_bounds__(SCAN_THREADS)
tc_final_fast_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int cap,
                     const float* __restrict__ tau, const float* __restrict__ qnorm, float max_norm, const float* __restrict__ q,
                     const float* __restrict__ vecs, const long long* __restrict__ ids, int d, int k, float* out_dist,
                     long long* out_ids, int* flags, const int* redo) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_m;
  const int qi = blockIdx.x;
  if (redo && !redo[qi]) return;  // only the queries the warp kernel could not finish
  float* qs = reinterpret_cast<float*>(smem);
  const size_t qbytes = ((size_t)d * 4 + 15) / 16 * 16;
  long long* ex_id = reinterpret_cast<long long*>(smem + qbytes);                    // [FIN_MAXW]
  uint32_t* ex_kd = reinterpret_cast<uint32_t*>(smem + qbytes + (size_t)FIN_MAXW * 8);  // [FIN_MAXW]
  int* rows = reinterpret_cast<int*>(smem + qbytes + (size_t)FIN_MAXW * 12);            // [FIN_MAXW]
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + qbytes + (size_t)FIN_MAXW * 16);  // [cap]
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[(size_t)qi * d + i];
  const int total = cand_cnt[qi];
  const int n = min(total, cap);
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = cand[(size_t)qi * cap + i];
  __syncthreads();
  bool fits;
  float a_k;
  const float two_eps = 2.f * tc_eps(L2, qnorm[qi], max_norm, d, false);
  const int have = window_select<L2, false>(keys, n, k, two_eps, qs, vecs, ids, d, ex_kd, ex_id, rows, FIN_MAXW, &fits, &a_k, &s_m);
  const float tq = tau[qi];
  const bool certified = fits && (total <= cap) && (tq == TC_INF || a_k + two_eps <= tq);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float api = 0.f;
    long long id = -1;
    if (i < have) {
      const float v = ord2f(ex_kd[i]);
      const float raw = L2 ? v : -v;
      api = L2 ? raw : __fsub_rn(1.0f, raw);
      id = ex_id[i];
    }
    out_dist[(size_t)qi * k + i] = api;
    out_ids[(size_t)qi * k + i] = id;
  }
  if (threadIdx.x == 0) flags[qi] = certified ? 0 : 1;
}


// ---------------------------------------------------------------------------------------------
// Finish kernel for the common case (at most HF_MAXW rows inside the 2*eps window): one 128-thread CTA per query.  Same
// rule as tc_final_fast_kernel — k-th approximate score A_k, every captured row with score <= A_k + 2 eps re-scored
// exactly in the reference order, exact top-k by (distance, id), certification — but the selection is done by ONE warp
// with shuffles / ballots and a warp-private histogram (no block-wide radix passes), all 32 quads of the CTA then
// re-score the window rows together (the ~40 random 3 KB row gathers per query are what this kernel waits on, so they are
// all issued at once) and a rank sort orders the exact pairs: three block barriers in total, 16 CTAs per SM.
// Queries whose window is larger set redo[q] = 1 and are finished by the block kernel.
// ---------------------------------------------------------------------------------------------
constexpr int HF_THREADS = 128;
constexpr int HF_STAGE = 512;   // captured keys staged in shared memory (the rest is re-read from global / L2 in every pass)
constexpr int HF_MAXW = 256;
constexpr int HF_ROW_BYTES = 48 * 1024;  // dynamic shared memory for the staged window rows (+ the query row)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// The ~40 window rows of a query are random 3 KB gathers from HBM: what the kernel waits on.  They are staged into
// shared memory with 16-byte cp.async copies (no registers held, so a CTA keeps a whole group of rows in flight), then
// quads re-score them from shared memory in the reference order.  rows_g rows of `pitch` bytes per group (d % 4 == 0).
template <bool L2>
static __global__ void __launch_bounds__(HF_THREADS)
tc_final_hybrid_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int cap,
                       const float* __restrict__ tau, const float* __restrict__ qnorm, float max_norm, const float* __restrict__ q,
                       const float* __restrict__ vecs, const long long* __restrict__ ids, int d, int k, int rows_g, int pitch,
                       float* out_dist, long long* out_ids, int* flags, int* redo) {
  extern __shared__ __align__(16) unsigned char s_dyn[];  // [pitch] query row, then rows_g staged rows
  __shared__ BlockSelShared S;
  __shared__ unsigned long long s_keys[HF_STAGE];
  __shared__ int s_rows[HF_MAXW];
  __shared__ uint32_t s_kd[HF_MAXW];
  __shared__ long long s_id[HF_MAXW];
  __shared__ int s_m;
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x;
  const int total = cand_cnt[qi];
  const int n = min(total, cap);
  const unsigned long long* c = cand + (size_t)qi * cap;
  const int nst = min(n, HF_STAGE);
  const int cpr = d >> 2;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < cpr; i += HF_THREADS) cp_async16(s_dyn + (size_t)i * 16, q + (size_t)qi * d + i * 4);
  for (int i = threadIdx.x; i < nst; i += HF_THREADS) s_keys[i] = c[i];
  const float two_eps = 2.f * tc_eps(L2, qnorm[qi], max_norm, d, false);
  if (threadIdx.x == 0) s_m = 0;
  __syncthreads();
  auto key_at = [&](int i) -> unsigned long long { return i < nst ? s_keys[i] : c[i]; };
  float a_k = TC_INF;
  if (n >= k) a_k = ord2f(block_kth_key_any(k, S, [&](auto f) { for (int i = threadIdx.x; i < n; i += HF_THREADS) f((uint32_t)(key_at(i) >> 32)); }));
  {
    const float window = a_k + two_eps;
    for (int base = 0; base < n; base += HF_THREADS) {
      const int i = base + threadIdx.x;
      unsigned long long e = 0;
      bool in = false;
      if (i < n) { e = key_at(i); in = ord2f((uint32_t)(e >> 32)) <= window; }
      const unsigned msk = __ballot_sync(0xffffffffu, in);
      int wbase = 0;
      if (lane == 0 && msk) wbase = atomicAdd(&s_m, __popc(msk));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      const int p = wbase + __popc(msk & ((1u << lane) - 1u));
      if (in && p < HF_MAXW) s_rows[p] = (int)(uint32_t)(e & 0xffffffffull);
    }
  }
  __syncthreads();
  const int m = s_m;
  if (m > HF_MAXW) { if (threadIdx.x == 0) redo[qi] = 1; return; }
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const float* qs = reinterpret_cast<const float*>(s_dyn);
  if (rows_g <= 0) {
    // direct form: every in-window row is prefetched towards L2 at once, then 32 quads score 32 rows per round straight from
    // global memory (no staging buffer: 16 CTAs per SM keep ~500 rows in flight per SM)
    const int lines = (d * 4 + 127) / 128;
    for (int i = threadIdx.x; i < m * lines; i += HF_THREADS) {
      const char* ptr = reinterpret_cast<const char*>(vecs + (size_t)(uint32_t)s_rows[i / lines] * d) + (size_t)(i % lines) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
    cp_async_wait_all();  // the query row
    __syncthreads();
    for (int base = 0; base < m; base += HF_THREADS / 4) {
      const int i = base + quad;
      const bool valid = i < m;
      const long long row = (long long)(uint32_t)s_rows[valid ? i : 0];
      const long long id = valid ? ids[row] : -1;
      const float v = quad_distance<L2>(vecs + (size_t)row * d, qs, d, t, true);
      if (valid && t == 0) { s_kd[i] = f2ord(L2 ? v : -v); s_id[i] = id; }
    }
    __syncthreads();
  } else {
  unsigned char* stage = s_dyn + pitch;
  for (int g0 = 0; g0 < m; g0 += rows_g) {
    const int ng = min(rows_g, m - g0);
    for (int i = threadIdx.x; i < ng * cpr; i += HF_THREADS) {
      const int r = i / cpr, ch = i - r * cpr;
      cp_async16(stage + (size_t)r * pitch + (size_t)ch * 16, vecs + (size_t)(uint32_t)s_rows[g0 + r] * d + ch * 4);
    }
    if (threadIdx.x < ng) s_id[g0 + threadIdx.x] = ids[(uint32_t)s_rows[g0 + threadIdx.x]];
    cp_async_wait_all();
    __syncthreads();
    for (int base = 0; base < ng; base += HF_THREADS / 4) {
      const int i = base + quad;
      const bool valid = i < ng;
      const float v = quad_distance<L2>(reinterpret_cast<const float*>(stage + (size_t)(valid ? i : 0) * pitch), qs, d, t, true);
      if (valid && t == 0) s_kd[g0 + i] = f2ord(L2 ? v : -v);
    }
    __syncthreads();
  }
  }
  const int have = min(m, k);
  for (int e = threadIdx.x; e < m; e += HF_THREADS) {
    const uint32_t d0 = s_kd[e];
    const long long i0 = s_id[e];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const uint32_t dj = s_kd[j];
      const long long ij = s_id[j];
      rank += (key_less(dj, ij, d0, i0) || (dj == d0 && ij == i0 && j < e)) ? 1 : 0;
    }
    if (rank < k) {
      const float v = ord2f(d0);
      const float raw = L2 ? v : -v;
      out_dist[(size_t)qi * k + rank] = L2 ? raw : __fsub_rn(1.0f, raw);
      out_ids[(size_t)qi * k + rank] = i0;
    }
  }
  for (int i = have + threadIdx.x; i < k; i += HF_THREADS) { out_dist[(size_t)qi * k + i] = 0.f; out_ids[(size_t)qi * k + i] = -1; }
  if (threadIdx.x == 0) {
    const float tq = tau[qi];
    flags[qi] = (total <= cap ? 0 : 2) | ((tq == TC_INF || a_k + two_eps <= tq) ? 0 : 1);  // 1: threshold too tight, 2: capture overflow
    redo[qi] = 0;
  }
}

template <bool L2>
static __global__ void __launch_bounds__(SCAN_THREADS)
tc_coarse_final_fast_kernel(const float* __restrict__ dense, long long ld, int nrows, const float* __restrict__ qnorm, float max_norm,
                            const float* __restrict__ q, const float* __restrict__ vecs, int d, int k, int maxw,
                            long long id_offset, int api_scores, long long* out_probes, float* out_raw, const int* redo) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_m;
  const int qi = blockIdx.x;
  if (redo && !redo[qi]) return;  // only the queries the register-select kernel could not finish
  float* qs = reinterpret_cast<float*>(smem);
  const size_t qbytes = ((size_t)d * 4 + 15) / 16 * 16;
  long long* ex_id = reinterpret_cast<long long*>(smem + qbytes);                       // [maxw]   maxw = pow2(nrows)
  uint32_t* ex_kd = reinterpret_cast<uint32_t*>(smem + qbytes + (size_t)maxw * 8);      // [maxw]
  int* rows = reinterpret_cast<int*>(smem + qbytes + (size_t)maxw * 12);                // [maxw]
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + qbytes + (size_t)maxw * 16);  // [maxw]
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[(size_t)qi * d + i];
  for (int i = threadIdx.x; i < nrows; i += blockDim.x) keys[i] = ((unsigned long long)f2ord(dense[(size_t)qi * ld + i]) << 32) | (unsigned)i;
  __syncthreads();
  bool fits;
  float a_k;
  const float two_eps = 2.f * tc_eps(L2, qnorm[qi], max_norm, d, true);
  const int have = window_select<L2, true>(keys, nrows, k, two_eps, qs, vecs, nullptr, d, ex_kd, ex_id, rows, maxw, &fits, &a_k, &s_m);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    out_probes[(size_t)qi * k + i] = i < have ? ex_id[i] + id_offset : -1;
    if (out_raw) {
      const float v = i < have ? ord2f(ex_kd[i]) : 0.f;
      const float raw = L2 ? v : -v;
      out_raw[(size_t)qi * k + i] = api_scores ? v : raw;  // api_scores: the ranking score itself (L2 distance | -ip), ascending
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident select (coarse finish): each of SEL_THREADS threads holds KPT keys, so a query's whole score row
// lives in registers and the CTA needs only ~13 KB of shared memory (8 CTAs per SM hide each other's barriers).
// ---------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int SEL_BINS = 1024;   // histogram bins per narrowing step (10 bits)
constexpr int SEL_WCAP = 512;    // window rows re-scored per query (more -> the one-sort fallback kernel)

struct SelShared {
  int hist[SEL_BINS];
  uint32_t red[2][SEL_THREADS / 32];
  uint32_t lo, hi;
  int k, c_le, m, ns;
};

// k-th smallest (k >= 1, at least k live keys) of the CTA's keys; 0xFFFFFFFF marks an empty slot.  Range-normalised
// radix select: bin = (key - lo) >> shift over the live range [lo, hi], narrow to the bin that holds the k-th key and
// repeat until bins are one key wide (two steps for the ~2^20-wide ranges of real score rows; never more than four).
// On return S.c_le = number of keys <= the result when known exactly, else k + 1.
template <int KPT>
__device__ uint32_t reg_kth_key(const uint32_t (&key)[KPT], int k, SelShared& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
  for (int j = 0; j < KPT; ++j)
    if (key[j] != 0xFFFFFFFFu) { mn = min(mn, key[j]); mx = max(mx, key[j]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if (lane == 0) { S.red[0][warp] = mn; S.red[1][warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0xFFFFFFFFu, b = 0u;
    for (int w = 0; w < SEL_THREADS / 32; ++w) { a = min(a, S.red[0][w]); b = max(b, S.red[1][w]); }
    S.lo = a; S.hi = b; S.k = k; S.c_le = k + 1;
  }
  __syncthreads();
  for (;;) {
    const uint32_t lo = S.lo, hi = S.hi;
    const int kk = S.k;
    const uint32_t range = hi - lo;
    if (range == 0) return lo;  // every remaining key is equal
    const int bits = 32 - __clz(range);
    const int shift = max(0, bits - 10);
    for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) S.hist[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < KPT; ++j)
      if (key[j] >= lo && key[j] <= hi) atomicAdd(&S.hist[(key[j] - lo) >> shift], 1);
    __syncthreads();
    if (warp == 0) {  // locate the bin that holds the kk-th key: 32 bins per lane
      int sum = 0;
      for (int j = 0; j < SEL_BINS / 32; ++j) sum += S.hist[lane * (SEL_BINS / 32) + j];
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const int before = incl - sum;
      if (before < kk && kk <= incl) {  // exactly one lane
        int acc = before;
        for (int j = 0; j < SEL_BINS / 32; ++j) {
          const int c = S.hist[lane * (SEL_BINS / 32) + j];
          if (acc < kk && kk <= acc + c) {
            const uint32_t nlo = lo + ((uint32_t)(lane * (SEL_BINS / 32) + j) << shift);
            S.lo = nlo;
            S.hi = min(hi, nlo + ((1u << shift) - 1u));
            S.k = kk - acc;
            if (shift == 0) S.c_le = (k - kk) + acc + c;
            break;
          }
          acc += c;
        }
      }
    }
    __syncthreads();
    if (shift == 0) return S.lo;
  }
}

// Coarse finish: top-nprobe rows of one query's dense score row.
//   full mode (out_raw != NULL or ties at the k-th score): re-score every row within 2*eps of the k-th approximate score
//     exactly and emit them in exact (score, row) order.
//   set mode  (probe tables, whose order nobody reads): rows more than 2*eps below the k-th score are certainly among
//     the exact top-k (only the k - 1 rows below it can beat them), rows more than 2*eps above are certainly not; only the
//     few rows in between are re-scored to fill the remaining slots.
// flags[qi] = 1 when the window overflowed SEL_WCAP: the one-sort kernel then redoes that query.
template <bool L2, int KPT>
static __global__ void __launch_bounds__(SEL_THREADS)
tc_coarse_select_kernel(const float* __restrict__ dense, long long ld, int nrows, const float* __restrict__ qnorm, float max_norm,
                        const float* __restrict__ q, const float* __restrict__ vecs, int d, int k, long long id_offset, int api_scores,
                        long long* out_probes, float* out_raw, int* flags, const int* redo_in) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SelShared S;
  __shared__ int s_rows[SEL_WCAP];
  __shared__ uint32_t s_kd[SEL_WCAP];
  __shared__ long long s_id[SEL_WCAP];
  const int qi = blockIdx.x;
  if (redo_in && !redo_in[qi]) { if (threadIdx.x == 0) flags[qi] = 0; return; }  // finished by the warp kernel
  float* qs = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < d; i += SEL_THREADS) qs[i] = q[(size_t)qi * d + i];
  uint32_t key[KPT];
#pragma unroll
  for (int j = 0; j < KPT; ++j) {
    const int r = threadIdx.x + j * SEL_THREADS;
    key[j] = r < nrows ? f2ord(dense[(size_t)qi * ld + r]) : 0xFFFFFFFFu;
  }
  const uint32_t kth = reg_kth_key<KPT>(key, k, S);  // barriers inside: qs is visible afterwards
  const float a_k = ord2f(kth);
  const float two_eps = 2.f * tc_eps(L2, qnorm[qi], max_norm, d, true);
  const uint32_t thr_hi = f2ord(__fadd_ru(a_k, two_eps));
  const bool set_mode = out_raw == nullptr && S.c_le == k;
  const uint32_t thr_lo = f2ord(__fsub_rd(a_k, two_eps));
  if (threadIdx.x == 0) { S.m = 0; S.ns = 0; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < KPT; ++j) {
    if (key[j] > thr_hi) continue;  // empty slots (0xFFFFFFFF) drop out here too
    const int r = threadIdx.x + j * SEL_THREADS;
    if (set_mode && key[j] <= thr_lo) out_probes[(size_t)qi * k + atomicAdd(&S.ns, 1)] = r + id_offset;
    else { const int pos = atomicAdd(&S.m, 1); if (pos < SEL_WCAP) s_rows[pos] = r; }
  }
  __syncthreads();
  const int m = S.m, ns = S.ns;
  if (m > SEL_WCAP) {  // degenerate score row (many near-equal scores): redo on the one-sort path
    if (threadIdx.x == 0) flags[qi] = 1;
    return;
  }
  if (threadIdx.x == 0) flags[qi] = 0;
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const bool vec = (d & 3) == 0;
  for (int base = 0; base < m; base += SEL_THREADS / 4) {
    const int i = base + quad;
    const bool valid = i < m;
    const long long row = (long long)s_rows[valid ? i : 0];
    const float v = quad_distance<L2>(vecs + (size_t)row * d, qs, d, t, vec);
    if (valid && t == 0) { s_kd[i] = f2ord(L2 ? v : -v); s_id[i] = row; }
  }
  int e2 = 2;
  while (e2 < m) e2 <<= 1;
  __syncthreads();
  for (int i = m + threadIdx.x; i < e2; i += SEL_THREADS) { s_kd[i] = KEY_SENTINEL_D; s_id[i] = KEY_SENTINEL_ID; }
  __syncthreads();
  block_sort_pair(s_kd, s_id, e2);
  const int need = k - ns;  // slots still open (set mode), all k otherwise
  for (int i = threadIdx.x; i < need; i += SEL_THREADS) {
    const bool have = i < m;
    out_probes[(size_t)qi * k + ns + i] = have ? s_id[i] + id_offset : -1;
    if (out_raw) {
      const float v = have ? ord2f(s_kd[i]) : 0.f;
      out_raw[(size_t)qi * k + i] = api_scores ? v : (L2 ? v : -v);
    }
  }
}


// Coarse finish for small centroid tables (nrows <= 32 * KPT <= 1024): one 128-thread CTA per query.  Warp 0 holds the
// score row in registers and does the selection (same full / set modes and +-2 eps rule as tc_coarse_select_kernel) with
// shuffles / ballots; all 32 quads then re-score the window rows together; a rank sort orders them.  Three block
// barriers.  Queries whose re-score window exceeds HC_MAXW rows set flags[qi] = 1 and go to the block kernels.
constexpr int HC_THREADS = 128;
constexpr int HC_MAXW = 256;
template <bool L2, int KPT>
static __global__ void __launch_bounds__(HC_THREADS)
tc_coarse_select_hybrid_kernel(const float* __restrict__ dense, long long ld, int nrows, const float* __restrict__ qnorm, float max_norm,
                               const float* __restrict__ q, const float* __restrict__ vecs, int d, int k, long long id_offset, int api_scores,
                               long long* out_probes, float* out_raw, int* flags) {
  __shared__ int s_hist[WS_BINS];
  __shared__ int s_rows[HC_MAXW];
  __shared__ uint32_t s_kd[HC_MAXW];
  __shared__ long long s_id[HC_MAXW];
  __shared__ int s_m, s_ns;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x;
  if (warp == 0) {
    uint32_t key[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      const int r = lane + j * 32;
      key[j] = r < nrows ? f2ord(dense[(size_t)qi * ld + r]) : 0xFFFFFFFFu;
    }
    int c_le;
    const uint32_t kth = warp_kth_key(k, s_hist, [&](auto f) {
#pragma unroll
      for (int j = 0; j < KPT; ++j) if (key[j] != 0xFFFFFFFFu) f(key[j]);
    }, c_le);
    const float a_k = ord2f(kth);
    const float two_eps = 2.f * tc_eps(L2, qnorm[qi], max_norm, d, true);
    const uint32_t thr_hi = f2ord(__fadd_ru(a_k, two_eps));
    const uint32_t thr_lo = f2ord(__fsub_rd(a_k, two_eps));
    const bool set_mode = out_raw == nullptr && c_le == k;
    int ns = 0, m = 0;
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      const int r = lane + j * 32;
      const bool in = key[j] <= thr_hi && key[j] != 0xFFFFFFFFu;
      const bool sure = in && set_mode && key[j] <= thr_lo;
      const unsigned ms = __ballot_sync(0xffffffffu, sure), mw = __ballot_sync(0xffffffffu, in && !sure);
      const unsigned lt = (1u << lane) - 1u;
      if (sure) out_probes[(size_t)qi * k + ns + __popc(ms & lt)] = r + id_offset;
      if (in && !sure) { const int p = m + __popc(mw & lt); if (p < HC_MAXW) s_rows[p] = r; }
      ns += __popc(ms);
      m += __popc(mw);
    }
    if (lane == 0) { s_m = m; s_ns = ns; }
  }
  __syncthreads();
  const int m = s_m, ns = s_ns;
  if (m > HC_MAXW) { if (threadIdx.x == 0) flags[qi] = 1; return; }  // (the block kernel rewrites the whole output row)
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const bool vec = (d & 3) == 0;
  const float* qrow = q + (size_t)qi * d;
  for (int base = 0; base < m; base += HC_THREADS / 4) {
    const int i = base + quad;
    const bool valid = i < m;
    const long long row = (long long)s_rows[valid ? i : 0];
    const float v = quad_distance<L2>(vecs + (size_t)row * d, qrow, d, t, vec);
    if (valid && t == 0) { s_kd[i] = f2ord(L2 ? v : -v); s_id[i] = row; }
  }
  __syncthreads();
  const int need = k - ns;  // slots still open (set mode), all k otherwise
  for (int e = threadIdx.x; e < m; e += HC_THREADS) {
    const uint32_t d0 = s_kd[e];
    const long long i0 = s_id[e];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const uint32_t dj = s_kd[j];
      const long long ij = s_id[j];
      rank += (key_less(dj, ij, d0, i0) || (dj == d0 && ij == i0 && j < e)) ? 1 : 0;
    }
    if (rank < need) {
      out_probes[(size_t)qi * k + ns + rank] = i0 + id_offset;
      if (out_raw) {
        const float v = ord2f(d0);
        out_raw[(size_t)qi * k + rank] = api_scores ? v : (L2 ? v : -v);
      }
    }
  }
  for (int i = m + threadIdx.x; i < need; i += HC_THREADS) {
    out_probes[(size_t)qi * k + ns + i] = -1;
    if (out_raw) out_raw[(size_t)qi * k + i] = 0.f;
  }
  if (threadIdx.x == 0) flags[qi] = 0;
}

template <bool L2>
static __global__ void __launch_bounds__(SCAN_THREADS)
tc_final_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int cap,
                const float* __restrict__ tau, const float* __restrict__ qnorm, float max_norm, const float* __restrict__ q,
                const float* __restrict__ vecs, const long long* __restrict__ ids, int d, int k, int pool_cap,
                float* out_dist, long long* out_ids, int* flags, const int* redo) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_m;
  const int qi = blockIdx.x;
  if (redo && !redo[qi]) return;
  float* qs = reinterpret_cast<float*>(smem);
  const size_t qbytes = ((size_t)d * 4 + 15) / 16 * 16;
  int* s_rows = reinterpret_cast<int*>(smem + qbytes);  // [FIN_SEG]
  unsigned char* pool = smem + qbytes + (size_t)FIN_SEG * 4;
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[(size_t)qi * d + i];
  const int total = cand_cnt[qi];
  const int n = min(total, cap);
  const unsigned long long* c = cand + (size_t)qi * cap;
  BlockSelect sel;
  sel.init(pool, pool_cap, k);
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
  const float a_k = have1 >= k ? ord2f(sel.kd[k - 1]) : TC_INF;
  __syncthreads();
  const float window = a_k + 2.f * tc_eps(L2, qnorm[qi], max_norm, d, false);  // +inf when fewer than k rows were captured
  const float tq = tau[qi];
  const bool certified = (total <= cap) && (tq == TC_INF || window <= tq);
  // pass 2: exact distances (reference AVX-512 order) of every captured row inside the window, FIN_SEG at a time
  sel.init(pool, pool_cap, k);
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const bool vec = (d & 3) == 0;
  for (int seg = 0; seg < n; seg += FIN_SEG) {
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    const int e1 = min(n, seg + FIN_SEG);
    for (int i = seg + threadIdx.x; i < e1; i += blockDim.x) {
      const unsigned long long e = c[i];
      if (ord2f((uint32_t)(e >> 32)) <= window) s_rows[atomicAdd(&s_m, 1)] = (int)(uint32_t)(e & 0xffffffffull);
    }
    __syncthreads();
    const int m = s_m;
    for (int base = 0; base < m; base += SCAN_QUADS) {
      sel.maybe_prune(SCAN_QUADS);
      const int i = base + quad;
      const bool valid = i < m;
      const long long row = (long long)(uint32_t)s_rows[valid ? i : 0];
      const float v = quad_distance<L2>(vecs + (size_t)row * d, qs, d, t, vec);
      if (valid && t == 0) {
        const uint32_t key = f2ord(L2 ? v : -v);
        const long long id = ids[row];
        if (sel.passes(key, id)) sel.push(key, id);
      }
    }
    __syncthreads();
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

// coarse quantiser: every approximate score is available (dense row), so the window select is always certified
template <bool L2>
static __global__ void __launch_bounds__(SCAN_THREADS)
tc_coarse_final_kernel(const float* __restrict__ dense, long long ld, int nrows, const float* __restrict__ qnorm, float max_norm,
                       const float* __restrict__ q, const float* __restrict__ vecs, int d, int k, int pool_cap,
                       long long id_offset, int api_scores, long long* out_probes, float* out_raw) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_m;
  const int qi = blockIdx.x;
  float* qs = reinterpret_cast<float*>(smem);
  const size_t qbytes = ((size_t)d * 4 + 15) / 16 * 16;
  int* s_rows = reinterpret_cast<int*>(smem + qbytes);
  unsigned char* pool = smem + qbytes + (size_t)FIN_SEG * 4;
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[(size_t)qi * d + i];
  const float* row_scores = dense + (size_t)qi * ld;
  BlockSelect sel;
  sel.init(pool, pool_cap, k);
  for (int base = 0; base < nrows; base += blockDim.x) {
    sel.maybe_prune(blockDim.x);
    const int i = base + threadIdx.x;
    if (i < nrows) {
      const float v = row_scores[i];
      if (v < TC_INF) { const uint32_t key = f2ord(v); if (sel.passes(key, i)) sel.push(key, i); }
    }
  }
  sel.prune();
  const float a_k = *sel.count >= k ? ord2f(sel.kd[k - 1]) : TC_INF;
  __syncthreads();
  const float window = a_k + 2.f * tc_eps(L2, qnorm[qi], max_norm, d, true);
  sel.init(pool, pool_cap, k);
  const int quad = threadIdx.x >> 2, t = threadIdx.x & 3;
  const bool vec = (d & 3) == 0;
  for (int seg = 0; seg < nrows; seg += FIN_SEG) {
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    const int e1 = min(nrows, seg + FIN_SEG);
    for (int i = seg + threadIdx.x; i < e1; i += blockDim.x)
      if (row_scores[i] <= window) s_rows[atomicAdd(&s_m, 1)] = i;
    __syncthreads();
    const int m = s_m;
    for (int base = 0; base < m; base += SCAN_QUADS) {
      sel.maybe_prune(SCAN_QUADS);
      const int i = base + quad;
      const bool valid = i < m;
      const int row = s_rows[valid ? i : 0];
      const float v = quad_distance<L2>(vecs + (size_t)row * d, qs, d, t, vec);
      if (valid && t == 0) {
        const uint32_t key = f2ord(L2 ? v : -v);
        if (sel.passes(key, row)) sel.push(key, row);
      }
    }
    __syncthreads();
  }
  sel.prune();
  const int have = *sel.count;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    out_probes[(size_t)qi * k + i] = i < have ? sel.kid[i] + id_offset : -1;
    if (out_raw) {
      const float v = i < have ? ord2f(sel.kd[i]) : 0.f;
      const float raw = L2 ? v : -v;
      out_raw[(size_t)qi * k + i] = api_scores ? v : raw;  // api_scores: the ranking score itself (L2 distance | -ip), ascending
    }
  }
}

static __global__ void tc_compact_flags_kernel(const int* __restrict__ flags, int nq, int* qmap, int* count) {
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

static __global__ void max_norm_kernel(const float* __restrict__ n2, long long n, float* out) {
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

static __global__ void row_norms_kernel(const float* __restrict__ x, long long n, int d, float* out) {
  const long long quad = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const int t = threadIdx.x & 3;
  const bool valid = quad < n;
  const float* row = x + (size_t)(valid ? quad : 0) * d;
  const float n2 = quad_distance<false>(row, row, d, t, (d & 3) == 0);
  if (valid && t == 0) out[quad] = n2;
}
void launch_row_norms(const float* x, int64_t n, int d, float* out, cudaStream_t s) {
  if (n <= 0) return;
  row_norms_kernel<<<(unsigned)cdiv(n * 4, 256), 256, 0, s>>>(x, n, d, out);
  B200VS_CUDA(cudaGetLastError());
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
static int64_t tc_sample_bound(const TcView& v, int64_t npairs) {  // spans are 4 chunks: at most (chunks + 3) / 4 per list
  const int64_t max_spans = (v.max_chunks_per_list + 3) / 4 + 1;
  return (npairs / TC_NQT + 1) * max_spans + (v.total_chunks + 3) / 4 + v.nlist;
}
static int tc_cand_cap(int k) { return std::min(16384, std::max(8192, next_pow2(256 * k))); }

// Per-device one-time setup: cudaFuncSetAttribute applies to the CURRENT device only, and one process may hold indexes
// on several GPUs (b200vs_params.device), so the opt-in shared-memory limits are raised once per device ordinal.
constexpr int TC_MAX_DEVICES = 64;
constexpr int TC_FAST_SMEM = 200 * 1024;  // opt-in limit of the one-sort finish kernels
static std::mutex g_tc_init_mu;
static int g_dev_sms[TC_MAX_DEVICES] = {0};
static int tc_init(int device) {
  if (device < 0 || device >= TC_MAX_DEVICES) fail(B200VS_EILLEGAL_PARAMETERS, "bad CUDA device ordinal");
  std::lock_guard<std::mutex> g(g_tc_init_mu);
  if (g_dev_sms[device]) return g_dev_sms[device];
  int sms = 0;
  B200VS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  B200VS_CUDA(cudaFuncSetAttribute(tc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  B200VS_CUDA(cudaFuncSetAttribute(tc_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));  // + static smem
  B200VS_CUDA(cudaFuncSetAttribute(tc_final_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));  // + static smem
  B200VS_CUDA(cudaFuncSetAttribute(tc_final_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));  // + static smem
  B200VS_CUDA(cudaFuncSetAttribute(tc_coarse_final_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));  // + static smem
  B200VS_CUDA(cudaFuncSetAttribute(tc_coarse_final_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));  // + static smem
  B200VS_CUDA(cudaFuncSetAttribute(tc_final_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_FAST_SMEM));
  B200VS_CUDA(cudaFuncSetAttribute(tc_final_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_FAST_SMEM));
  B200VS_CUDA(cudaFuncSetAttribute(tc_coarse_final_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_FAST_SMEM));
  B200VS_CUDA(cudaFuncSetAttribute(tc_coarse_final_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_FAST_SMEM));
  B200VS_CUDA(cudaFuncSetAttribute(tc_final_hybrid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  B200VS_CUDA(cudaFuncSetAttribute(tc_final_hybrid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  g_dev_sms[device] = sms;  // only after every attribute call succeeded
  return sms;
}

bool tc_eligible(const IndexBase* ix, const TcView& v, int64_t nq, int k, int nprobe, const SearchCtx& sc) {
  if (sc.exact_only) return false;
  if (ix->dim % 4 != 0 || ix->dim < 32) return false;      // TMA: 16-byte row pitch; tiny d is not worth a tile
  if (k > 128 || nq < 16) return false;                    // wide k / tiny batches: exact scan
  if (v.arena_rows <= 0 || v.arena_rows >= (1LL << 31)) return false;
  if (!v.norms || !v.vecs) return false;
  const int64_t npairs = nq * nprobe;
  if (npairs >= (1LL << 30)) return false;
  if (tc_item_bound(v, npairs) >= (1LL << 28)) return false;
  if (tc_sample_bound(v, npairs) * TC_NQT * TC_SAMPLE * 4 > (4LL << 30)) return false;  // sample buffer too large (TC_SAMPLE rows per sampled item)
  return true;
}

// everything the passes share: rounded queries, pair inversion, gathered B block, work items, tensor maps
struct TcPlan {
  float* q32; float* qnorm;
  int *cnt, *pos, *pair_off, *item_off, *totals, *pair_query, *sample_list, *work;
  float* bws;
  TcItem* items;
  int64_t bound, sbound, npairs;
  CUtensorMap tmA, tmA32, tmB16, tmB32, tmB64, tmB128, tmQ;
  bool gather;
  int num_sms;
};

static TcPlan tc_prepare(IndexBase* ix, const TcView& v, int64_t nq, const float* q, const long long* probes, int nprobe, cudaStream_t s) {
  const int num_sms = tc_init(ix->device);
  const int d = ix->dim;
  auto& S = ix->scratch;
  TcPlan P;
  P.num_sms = num_sms;
  P.npairs = nq * nprobe;
  P.bound = tc_item_bound(v, P.npairs);
  P.sbound = tc_sample_bound(v, P.npairs);
  P.q32 = S.alloc<float>((size_t)nq * d);
  P.qnorm = S.alloc<float>(nq);
  P.cnt = S.alloc<int>(v.nlist);
  P.pos = S.alloc<int>(P.npairs);
  P.pair_off = S.alloc<int>(v.nlist);
  P.item_off = S.alloc<int>(v.nlist);
  P.totals = S.alloc<int>(4);
  P.work = S.alloc<int>(4);
  P.pair_query = S.alloc<int>(P.npairs);
  P.sample_list = S.alloc<int>(P.sbound);
  P.bws = S.alloc<float>((size_t)(P.npairs + TC_NQT) * d);
  P.items = S.alloc<TcItem>(P.bound);
  B200VS_CUDA(cudaMemsetAsync(P.cnt, 0, (size_t)v.nlist * 4, s));
  B200VS_CUDA(cudaMemsetAsync(P.work, 0, 16, s));
  tc_prep_queries_kernel<<<(unsigned)cdiv(nq * 32, 256), 256, 0, s>>>(q, nq, d, P.q32, P.qnorm);
  tc_count_pairs_kernel<<<(unsigned)cdiv(P.npairs, 256), 256, 0, s>>>(probes, P.npairs, P.cnt, P.pos);
  tc_plan_kernel<<<1, 1024, 0, s>>>(P.cnt, v.list_len, v.nlist, P.pair_off, P.item_off, P.totals);
  static const bool use_gather4 = getenv("B200VS_GATHER4") && atoi(getenv("B200VS_GATHER4")) != 0;
  P.gather = use_gather4;
  tc_fill_pairs_kernel<<<(unsigned)cdiv(P.npairs * 32, 256), 256, 0, s>>>(probes, P.pos, P.pair_off, v.list_len, P.npairs, nprobe, d, P.q32, P.pair_query, P.gather ? nullptr : P.bws);
  tc_items_kernel<<<(unsigned)cdiv(v.nlist, 128), 128, 0, s>>>(P.cnt, v.list_len, P.pair_off, P.item_off, v.nlist, P.items, P.totals, P.sample_list);
  B200VS_CUDA(cudaGetLastError());
  P.tmA = make_tmap(v.vecs, v.arena_rows, d, TC_BM);
  P.tmA32 = make_tmap(v.vecs, v.arena_rows, d, TC_SAMPLE);
  P.tmB16 = make_tmap(P.bws, P.npairs + TC_NQT, d, 16);
  P.tmB32 = make_tmap(P.bws, P.npairs + TC_NQT, d, 32);
  P.tmB64 = make_tmap(P.bws, P.npairs + TC_NQT, d, 64);
  P.tmB128 = make_tmap(P.bws, P.npairs + TC_NQT, d, 128);
  P.tmQ = make_tmap(P.q32, nq, d, 1);  // gather4 source: one row per box
  ix->launch_count(5);
  return P;
}

static TcParams tc_params(const TcView& v, const TcPlan& P, int d, bool l2) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.ids = v.ids; p.norms = v.norms; p.list_off = v.list_off; p.d = d; p.items = P.items; p.totals = P.totals;
  p.sample_list = P.sample_list; p.pair_query = P.pair_query; p.l2 = l2 ? 1 : 0; p.add_norm = 1; p.b_gather = P.gather ? 1 : 0;
  return p;
}

static void tc_launch(const TcPlan& P, const TcParams& p, int64_t work_bound, cudaStream_t s) {
  const int grid = (int)std::min<int64_t>(P.num_sms, std::max<int64_t>(1, work_bound));
  tc_scan_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(P.tmA, P.tmA32, P.tmB16, P.tmB32, P.tmB64, P.tmB128, P.tmA, P.tmB16, P.tmB32, P.tmB64, P.tmB128, P.tmQ, p);
  B200VS_CUDA(cudaGetLastError());
}

void run_scan_mapped(IndexBase* ix, const ScanJob& job, int64_t nq_max, const int* qmap, const int* qcount, const float* queries,
                     int k, float* out_dist, long long* out_ids, cudaStream_t s);

void tc_search(IndexBase* ix, const TcView& v, bool l2, int64_t nq, const float* q, int k, const long long* probes,
               int nprobe, const SearchCtx& sc, float* out_dist, long long* out_ids, cudaStream_t s) {
  const int d = ix->dim;
  const int cap = tc_cand_cap(k);
  auto& S = ix->scratch;
  ix->phase(IndexBase::PH_PLAN, s);
  TcPlan P = tc_prepare(ix, v, nq, q, probes, nprobe, s);
  // rows sampled per span for the thresholds: 32 (one TMA box, capture threshold = tau) by default.
  // B200VS_SAMPLE_ROWS=128 samples the whole first tile of each span and captures under tau + 2 eps: certified by
  // construction (no fallback except on overflow) at the price of a longer sample / threshold pass — the setting for data
  // whose thin samples fail certification often.  Measured on the benchmark (list-sharded over 4 GPUs as well) the thin
  // sample wins: the finish kernel is bound by per-query latency, not by the number of captured rows.
  const char* env_srows = getenv("B200VS_SAMPLE_ROWS");
  const int srows = env_srows && atoi(env_srows) == TC_BM && P.sbound * TC_NQT * TC_BM * 4 <= (4LL << 30) ? TC_BM : TC_SAMPLE;
  float* sample = S.alloc<float>((size_t)P.sbound * TC_NQT * srows);
  float* tau = S.alloc<float>(nq);
  unsigned long long* cand = S.alloc<unsigned long long>((size_t)nq * cap);
  int* cand_cnt = S.alloc<int>(nq);
  int* flags = S.alloc<int>(nq);
  int* qmap = S.alloc<int>(nq);
  int* qcount = S.alloc<int>(1);
  B200VS_CUDA(cudaMemsetAsync(cand_cnt, 0, (size_t)nq * 4, s));

  TcParams p = tc_params(v, P, d, l2);
  p.sample = sample; p.sample_rows = srows; p.tau = tau; p.cand = cand; p.cand_cnt = cand_cnt; p.cap = cap;
  p.filt.has_range = sc.has_range; p.filt.negate = sc.negate; p.filt.rmin = sc.rmin; p.filt.rmax = sc.rmax;
  p.filt.sorted_ids = sc.sorted_ids_dev; p.filt.n_ids = sc.n_ids;
  const int pool = select_pool_cap(k, SCAN_THREADS);
  const size_t sel_smem = BlockSelect::smem_bytes(pool);

  // 1) sample pass -> per-query capture thresholds
  p.mode = 0; p.work_counter = P.work;
  ix->phase(IndexBase::PH_SAMPLE, s);
  tc_launch(P, p, P.sbound, s);
  ix->phase(IndexBase::PH_TAU, s);
  int* redo = S.alloc<int>(nq);
  const size_t tau_smem = (size_t)TAU_PL * 8 + std::max(sel_smem, (size_t)TAU_SORT * 8);
  if (srows == TC_SAMPLE) {  // warp per query; the block kernel only redoes queries with more sampled spans than a warp stages
    tc_tau_warp_kernel<<<(unsigned)cdiv(nq, WT_WARPS), WT_WARPS * 32, 0, s>>>(probes, P.pos, P.cnt, P.item_off, v.list_len, P.items, nprobe, sample, k, (int)nq, l2 ? 1 : 0, P.qnorm, v.max_norm, d, cap, tau, redo);
    tc_tau_kernel<<<(unsigned)nq, SCAN_THREADS, tau_smem, s>>>(probes, P.pos, P.cnt, P.item_off, v.list_len, P.items, nprobe, sample, srows, k, pool, l2 ? 1 : 0, P.qnorm, v.max_norm, d, 0, cap, tau, redo);
  } else {
    tc_tau_kernel<<<(unsigned)nq, SCAN_THREADS, tau_smem, s>>>(probes, P.pos, P.cnt, P.item_off, v.list_len, P.items, nprobe, sample, srows, k, pool, l2 ? 1 : 0, P.qnorm, v.max_norm, d, 1, cap, tau, nullptr);
  }
  // 2) capture pass: stream every probed list chunk once, keep rows under the threshold
  p.mode = 1; p.work_counter = P.work + 1;
  ix->phase(IndexBase::PH_CAPTURE, s);
  {
    ScopedKernelTimer timer(ix, s, ix->profiling);
    tc_launch(P, p, P.bound, s);
    timer.stop();
  }
  // 3) window select + exact rerank + certification
  ix->phase(IndexBase::PH_FINAL, s);
  const size_t fast_smem = ((size_t)d * 4 + 15) / 16 * 16 + (size_t)FIN_MAXW * 16 + (size_t)cap * 8;
  const int* fin_redo = nullptr;
  const int hf_pitch = d * 4 + 16;  // staged row pitch: 16-byte aligned, rows land on different banks
  static const bool hf_direct = !(getenv("B200VS_FINAL_STAGED") && atoi(getenv("B200VS_FINAL_STAGED")) != 0);
  const int hf_rows = hf_direct ? 0 : std::min(32, std::max(HF_ROW_BYTES, 2 * hf_pitch) / hf_pitch - 1);
  const size_t hsm = (size_t)(hf_rows + 1) * hf_pitch;
  if (2 * k <= HF_MAXW && hsm <= 96 * 1024) {  // warp select + block re-score; the block kernels below only redo queries whose window exceeds HF_MAXW rows
    if (l2) tc_final_hybrid_kernel<true><<<(unsigned)nq, HF_THREADS, hsm, s>>>(cand, cand_cnt, cap, tau, P.qnorm, v.max_norm, q, v.vecs, v.ids, d, k, hf_rows, hf_pitch, out_dist, out_ids, flags, redo);
    else tc_final_hybrid_kernel<false><<<(unsigned)nq, HF_THREADS, hsm, s>>>(cand, cand_cnt, cap, tau, P.qnorm, v.max_norm, q, v.vecs, v.ids, d, k, hf_rows, hf_pitch, out_dist, out_ids, flags, redo);
    fin_redo = redo;
    ix->launch_count(1);
  }
  if (fast_smem <= (size_t)TC_FAST_SMEM) {  // one 64-bit sort of the captured rows + a small exact sort
    if (l2) tc_final_fast_kernel<true><<<(unsigned)nq, SCAN_THREADS, fast_smem, s>>>(cand, cand_cnt, cap, tau, P.qnorm, v.max_norm, q, v.vecs, v.ids, d, k, out_dist, out_ids, flags, fin_redo);
    else tc_final_fast_kernel<false><<<(unsigned)nq, SCAN_THREADS, fast_smem, s>>>(cand, cand_cnt, cap, tau, P.qnorm, v.max_norm, q, v.vecs, v.ids, d, k, out_dist, out_ids, flags, fin_redo);
  } else {
    const size_t fin_smem = ((size_t)d * 4 + 15) / 16 * 16 + (size_t)FIN_SEG * 4 + sel_smem;
    if (l2) tc_final_kernel<true><<<(unsigned)nq, SCAN_THREADS, fin_smem, s>>>(cand, cand_cnt, cap, tau, P.qnorm, v.max_norm, q, v.vecs, v.ids, d, k, pool, out_dist, out_ids, flags, fin_redo);
    else tc_final_kernel<false><<<(unsigned)nq, SCAN_THREADS, fin_smem, s>>>(cand, cand_cnt, cap, tau, P.qnorm, v.max_norm, q, v.vecs, v.ids, d, k, pool, out_dist, out_ids, flags, fin_redo);
  }
  tc_compact_flags_kernel<<<1, 1024, 0, s>>>(flags, (int)nq, qmap, qcount);
  B200VS_CUDA(cudaGetLastError());
  ix->launch_count(5);
  ix->stats[1] = nq;
  // 4) uncertified queries re-run on the exact scan (device-side count: blocks beyond it exit immediately)
  ScanJob job;
  job.l2 = l2; job.vecs = v.vecs; job.ids = v.ids; job.d = d; job.sc = &sc;
  if (!v.flat) {
    job.mode = 1; job.probes = probes; job.nprobe = nprobe; job.list_off = v.list_off; job.list_len = v.list_len;
    job.avg_candidates = v.nlist > 0 ? (double)v.arena_rows * nprobe / v.nlist : 0;
  }
  else { job.mode = 0; job.n = v.arena_rows; }
  ix->phase(IndexBase::PH_FALLBACK, s);
  run_scan_mapped(ix, job, nq, qmap, qcount, q, k, out_dist, out_ids, s);
  ix->phase(IndexBase::PH_OTHER, s);
  if (ix->profiling && getenv("B200VS_DEBUG_FLAGS")) {
    std::vector<int> hf(nq), hc(nq);
    std::vector<float> ht(nq);
    B200VS_CUDA(cudaMemcpyAsync(hf.data(), flags, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    B200VS_CUDA(cudaMemcpyAsync(hc.data(), cand_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    B200VS_CUDA(cudaMemcpyAsync(ht.data(), tau, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    B200VS_CUDA(cudaStreamSynchronize(s));
    long long tot = 0; int mx = 0, r1 = 0, r2 = 0, ninf = 0;
    for (int64_t i = 0; i < nq; ++i) { tot += hc[i]; mx = std::max(mx, hc[i]); r1 += (hf[i] & 1); r2 += (hf[i] & 2) ? 1 : 0; ninf += std::isinf(ht[i]) ? 1 : 0; }
    fprintf(stderr, "[b200vs] tc_search nq=%lld cap=%d captured avg=%.1f max=%d flagged: threshold=%d overflow=%d tau_inf=%d\n", (long long)nq, cap, (double)tot / nq, mx, r1, r2, ninf);
    for (int64_t i = 0, shown = 0; i < nq && shown < 4; ++i) if (hf[i]) { fprintf(stderr, "[b200vs]   q=%lld flag=%d captured=%d tau=%g\n", (long long)i, hf[i], hc[i], ht[i]); ++shown; }
  }
  if (ix->profiling) {
    int h = 0;
    B200VS_CUDA(cudaMemcpyAsync(&h, qcount, 4, cudaMemcpyDeviceToHost, s));
    B200VS_CUDA(cudaStreamSynchronize(s));
    ix->stats[2] = h;
    int t[4] = {0, 0, 0, 0};
    B200VS_CUDA(cudaMemcpyAsync(t, P.totals, 16, cudaMemcpyDeviceToHost, s));
    B200VS_CUDA(cudaStreamSynchronize(s));
    ix->stats[6] = t[0];  // work items of the capture pass
    ix->stats[7] = t[3];  // sum over items of (128-row tiles x padded query columns)
  }
}

bool tc_coarse_eligible(const IndexBase* ix, int64_t nq, int nrows, int nprobe) {
  if (ix->dim % 4 != 0 || ix->dim < 32) return false;
  if (nq < 16 || nrows < 64 || nprobe > 1024) return false;
  if (ix->dim > 8192) return false;  // the finish kernels keep the query row in (un-opted) dynamic shared memory
  if ((int64_t)nq * nrows * 4 > (1LL << 30)) return false;  // dense score matrix
  return true;
}

// Coarse quantiser: ONE dense launch of the tile kernel in split mode — error-compensated TF32, the three products
// hi*hi + lo*hi + hi*lo accumulate into the same TMEM accumulator (the lo operands ride in the next ring stage).  The
// operand error drops to 2^-19 relative, so the exact re-score window stays a few dozen rows even when all centroid
// distances are nearly equal (uniform high-dimensional data); what remains of eps is the FP32 accumulation term.
void tc_coarse(IndexBase* ix, const TcView& v, bool l2, int64_t nq, const float* q, int nprobe, long long* out_probes,
               float* out_raw, cudaStream_t s) {
  const int num_sms = tc_init(ix->device);
  const int d = ix->dim;
  const int nrows = (int)v.arena_rows;
  auto& S = ix->scratch;
  float* qhi = S.alloc<float>((size_t)nq * d);
  float* qlo = S.alloc<float>((size_t)nq * d);
  float* qnorm = S.alloc<float>(nq);
  const int chunk = nrows <= 16384 ? TC_BM : TC_CHUNK;  // small tables: one tile per item so the whole chip has work
  const int nitems = (int)(cdiv(nrows, chunk) * cdiv(nq, TC_NQT));
  TcItem* items = S.alloc<TcItem>(nitems);
  int* totals = S.alloc<int>(4);
  int* work = S.alloc<int>(4);
  float* dense = S.alloc<float>((size_t)nq * nrows);
  ix->phase(IndexBase::PH_COARSE_PREP, s);
  B200VS_CUDA(cudaMemsetAsync(work, 0, 16, s));
  tc_prep_queries_split_kernel<<<(unsigned)cdiv(nq * 32, 256), 256, 0, s>>>(q, nq, d, qhi, qlo, qnorm);
  tc_coarse_items_kernel<<<(unsigned)cdiv(nitems, 128), 128, 0, s>>>(nrows, (int)nq, chunk, items, totals);
  TcPlan P;
  memset(&P, 0, sizeof(P));
  P.items = items; P.totals = totals;
  TcParams p = tc_params(v, P, d, l2);
  p.mode = 2; p.dense = dense; p.dense_ld = nrows; p.pair_query = nullptr;
  const CUtensorMap a_hi = make_tmap(v.vecs_hi, nrows, d, TC_BM), a_lo = make_tmap(v.vecs_lo, nrows, d, TC_BM);
  const CUtensorMap b16 = make_tmap(qhi, nq, d, 16), b32 = make_tmap(qhi, nq, d, 32), b64 = make_tmap(qhi, nq, d, 64), b128 = make_tmap(qhi, nq, d, 128);
  const CUtensorMap l16 = make_tmap(qlo, nq, d, 16), l32 = make_tmap(qlo, nq, d, 32), l64 = make_tmap(qlo, nq, d, 64), l128 = make_tmap(qlo, nq, d, 128);
  const int grid = (int)std::min<int64_t>(num_sms, std::max(1, nitems));
  p.work_counter = work; p.split = 1; p.dense_accum = 0; p.add_norm = 1;
  ix->phase(IndexBase::PH_COARSE_SCAN, s);
  tc_scan_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(a_hi, a_hi, b16, b32, b64, b128, a_lo, l16, l32, l64, l128, b16, p);
  B200VS_CUDA(cudaGetLastError());
  ix->phase(IndexBase::PH_COARSE_FINAL, s);
  const int maxw = std::max(2, next_pow2(nrows));
  const size_t fast_smem = ((size_t)d * 4 + 15) / 16 * 16 + (size_t)maxw * 24;
  if (nrows <= COARSE_FAST && nprobe <= nrows && fast_smem <= (size_t)TC_FAST_SMEM) {
    // warp-select / block-re-score kernel for small tables, then the register-resident block select (8 CTAs per SM) for larger tables
    // and for the queries the warp kernel flagged, then the one-sort kernel for what is still flagged (normally none)
    int* redo0 = S.alloc<int>(nq);
    int* redo = S.alloc<int>(nq);
    const size_t qsm = ((size_t)d * 4 + 15) / 16 * 16;
    const int kpt = (nrows + SEL_THREADS - 1) / SEL_THREADS;
    const int* redo_in = nullptr;
    if (nrows <= 1024 && 2 * nprobe <= HC_MAXW) {
#define B200VS_HSEL(L2_, KPT_) tc_coarse_select_hybrid_kernel<L2_, KPT_><<<(unsigned)nq, HC_THREADS, 0, s>>>(dense, nrows, nrows, qnorm, v.max_norm, q, v.vecs, d, nprobe, v.id_offset, v.api_scores ? 1 : 0, out_probes, out_raw, redo0)
      if (nrows <= 256) { if (l2) B200VS_HSEL(true, 8); else B200VS_HSEL(false, 8); }
      else if (nrows <= 512) { if (l2) B200VS_HSEL(true, 16); else B200VS_HSEL(false, 16); }
      else { if (l2) B200VS_HSEL(true, 32); else B200VS_HSEL(false, 32); }
#undef B200VS_HSEL
      redo_in = redo0;
      ix->launch_count(1);
    }
#define B200VS_SEL(L2_, KPT_) tc_coarse_select_kernel<L2_, KPT_><<<(unsigned)nq, SEL_THREADS, qsm, s>>>(dense, nrows, nrows, qnorm, v.max_norm, q, v.vecs, d, nprobe, v.id_offset, v.api_scores ? 1 : 0, out_probes, out_raw, redo, redo_in)
    if (kpt <= 4) { if (l2) B200VS_SEL(true, 4); else B200VS_SEL(false, 4); }
    else if (kpt <= 8) { if (l2) B200VS_SEL(true, 8); else B200VS_SEL(false, 8); }
    else if (kpt <= 16) { if (l2) B200VS_SEL(true, 16); else B200VS_SEL(false, 16); }
    else { if (l2) B200VS_SEL(true, 32); else B200VS_SEL(false, 32); }
#undef B200VS_SEL
    if (l2) tc_coarse_final_fast_kernel<true><<<(unsigned)nq, FIN_THREADS, fast_smem, s>>>(dense, nrows, nrows, qnorm, v.max_norm, q, v.vecs, d, nprobe, maxw, v.id_offset, v.api_scores ? 1 : 0, out_probes, out_raw, redo);
    else tc_coarse_final_fast_kernel<false><<<(unsigned)nq, FIN_THREADS, fast_smem, s>>>(dense, nrows, nrows, qnorm, v.max_norm, q, v.vecs, d, nprobe, maxw, v.id_offset, v.api_scores ? 1 : 0, out_probes, out_raw, redo);
    ix->launch_count(1);
  } else {
    const int pool = select_pool_cap(nprobe, SCAN_THREADS);
    const size_t smem = ((size_t)d * 4 + 15) / 16 * 16 + (size_t)FIN_SEG * 4 + BlockSelect::smem_bytes(pool);
    if (smem > 160 * 1024) fail(B200VS_EILLEGAL_PARAMETERS, "coarse quantiser: dimension / nprobe too large for one SM's shared memory");
    if (l2) tc_coarse_final_kernel<true><<<(unsigned)nq, SCAN_THREADS, smem, s>>>(dense, nrows, nrows, qnorm, v.max_norm, q, v.vecs, d, nprobe, pool, v.id_offset, v.api_scores ? 1 : 0, out_probes, out_raw);
    else tc_coarse_final_kernel<false><<<(unsigned)nq, SCAN_THREADS, smem, s>>>(dense, nrows, nrows, qnorm, v.max_norm, q, v.vecs, d, nprobe, pool, v.id_offset, v.api_scores ? 1 : 0, out_probes, out_raw);
  }
  B200VS_CUDA(cudaGetLastError());
  ix->phase(IndexBase::PH_OTHER, s);
  ix->launch_count(4);
}

}  // namespace b200vs
