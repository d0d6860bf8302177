// ivf_pq.cu — IVF-PQ index: coarse quantiser + by-residual product quantiser, LUT scan on the GPU.
//
// Replaces faiss::IndexIVFPQ behind VectorIndexRawIvfPq (src/vector/vector_index_raw_ivf_pq.cc:551-571 Init,
// :457-500 Train, :157-210 Search) and the outer VectorIndexIvfPq switch between an inner Flat index and the
// real IVF-PQ (src/vector/vector_index_ivf_pq.cc:327-395: trained with fewer than
// max(256*nlist, 256*2^nbits) vectors -> Flat).
//
// Arithmetic (matches oracle/oracle_pq.cc so results are comparable bit for bit on a shared trained state):
//   encode : residual r = x - c[list]; per sub-space argmin_j ||r_m - cw[m][j]||^2, sequential un-fused sums,
//            first minimum wins.
//   IP     : dis = <q, c[list]> (hooked AVX-512 order) + sum_m sim[m][code_m],      sim[m][j] = <q_m, cw[m][j]>
//   L2     : dis = ||q - c[list]||^2 + sum_m (T[list][m][code_m] - 2 sim[m][code_m]), T = ||cw||^2 + 2 <c_m, cw>
//   LUT sums run sequentially over m in FP32.
// Kernels: pq_sim_kernel (per-query LUT), pq_precompute_kernel (T), pq_encode_kernel, pq_scan_select_kernel.
#include <algorithm>
#include <memory>

#include "index.h"
#include "ivf_common.h"
#include "scan_kernels.cuh"

namespace b200vs {

IndexBase* make_flat(b200vs_metric m, int d, const b200vs_params& p);

namespace {

constexpr int KSUB = 256;
constexpr int PQ_THREADS = 512;  // code-scan CTA: one 96 KB LUT serves 16 warps (2 CTAs per SM -> 32 warps hide the code-load latency)

__device__ __forceinline__ float ip_seq(const float* a, const float* b, int n) {
  float r = 0.f;
  for (int i = 0; i < n; ++i) r = __fadd_rn(r, __fmul_rn(a[i], b[i]));
  return r;
}

// sim[q][m][j] = <q_m, cw[m][j]>
__global__ void pq_sim_kernel(const float* __restrict__ q, const float* __restrict__ cb, long long nq, int d, int M, float* sim) {
  const int dsub = d / M;
  const long long qi = blockIdx.x;
  const float* qv = q + (size_t)qi * d;
  for (int i = threadIdx.x; i < M * KSUB; i += blockDim.x) {
    const int m = i / KSUB;
    sim[(size_t)qi * M * KSUB + i] = ip_seq(qv + m * dsub, cb + (size_t)i * dsub, dsub);
  }
}

// T[l][m][j] = ||cw||^2 + 2 <c_l,m , cw>
__global__ void pq_precompute_kernel(const float* __restrict__ cent, const float* __restrict__ cb, int d, int M, float* pre) {
  const int dsub = d / M;
  const int l = blockIdx.x;
  for (int i = threadIdx.x; i < M * KSUB; i += blockDim.x) {
    const int m = i / KSUB;
    const float* cw = cb + (size_t)i * dsub;
    const float r2 = ip_seq(cw, cw, dsub);
    const float cr = ip_seq(cent + (size_t)l * d + m * dsub, cw, dsub);
    pre[(size_t)l * M * KSUB + i] = __fadd_rn(r2, __fmul_rn(2.f, cr));
  }
}

// one warp per row: code[m] = argmin_j ||(x - c)_m - cw[m][j]||^2
__global__ void pq_encode_kernel(const float* __restrict__ x, const long long* __restrict__ list, const float* __restrict__ cent,
                                 const float* __restrict__ cb, long long n, int d, int M, unsigned char* codes_out) {
  extern __shared__ float s_res[];  // [warps_per_block][d]
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  if (w >= n) return;
  const int dsub = d / M;
  float* res = s_res + (size_t)wib * d;
  const float* xr = x + (size_t)w * d;
  const float* c = cent + (size_t)list[w] * d;
  for (int i = lane; i < d; i += 32) res[i] = __fsub_rn(xr[i], c[i]);
  __syncwarp();
  for (int m = 0; m < M; ++m) {
    const float* rs = res + m * dsub;
    float best = __int_as_float(0x7f800000);
    int bj = 0x7fffffff;
    for (int j = lane; j < KSUB; j += 32) {
      const float* cw = cb + ((size_t)m * KSUB + j) * dsub;
      float v = 0.f;
      for (int i = 0; i < dsub; ++i) { const float t = __fsub_rn(rs[i], cw[i]); v = __fadd_rn(v, __fmul_rn(t, t)); }
      if (v < best) { best = v; bj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov < best || (ov == best && oj < bj)) { best = ov; bj = oj; }
    }
    if (lane == 0) codes_out[(size_t)w * M + m] = (unsigned char)bj;
  }
}

__global__ void scatter_codes_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ src_ids,
                                     const long long* __restrict__ slots, long long n, int M, unsigned char* codes, long long* ids) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * M) return;
  const long long r = i / M;
  const int m = (int)(i % M);
  const long long s = slots[r];
  codes[(size_t)s * M + m] = src[i];
  if (m == 0) ids[s] = src_ids[r];
}

__global__ void move_codes_kernel(const unsigned char* __restrict__ scodes, const long long* __restrict__ sids,
                                  const long long* __restrict__ src_rows, const long long* __restrict__ dst_rows, long long n, int M,
                                  unsigned char* dcodes, long long* dids) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * M) return;
  const long long r = i / M;
  const int m = (int)(i % M);
  dcodes[(size_t)dst_rows[r] * M + m] = scodes[(size_t)src_rows[r] * M + m];
  if (m == 0) dids[dst_rows[r]] = sids[src_rows[r]];
}

struct PqScanArgs {
  const unsigned char* codes;
  const long long* ids;
  const long long* probes;  // [nq, nprobe]
  const float* coarse;      // [nq, nprobe] raw metric value (ip or L2)
  const long long* list_off;
  const int* list_len;
  const float* sim;  // [nq, M*256]
  const float* pre;  // [nlist, M*256] (L2) or null
  int M, nprobe, k, nsplit, pool_cap;
  int has_thr;       // range search: fixed initial threshold (strict), k = max results kept
  uint32_t thr_key;
  uint32_t* ws_kd;
  long long* ws_kid;
  FilterDev filt;
};

// block (split, query): probes split round-robin; LUT in shared memory; one thread per code row
template <bool L2>
__global__ void __launch_bounds__(PQ_THREADS) pq_scan_select_kernel(const PqScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* tab = reinterpret_cast<float*>(smem);
  const int M = a.M, T = M * KSUB;
  const int qi = blockIdx.y, split = blockIdx.x;
  BlockSelect sel;
  sel.init(smem + (size_t)T * 4, a.pool_cap, a.k);
  if (a.has_thr && threadIdx.x == 0) { *sel.thr_d = a.thr_key; *sel.thr_id = (long long)0x8000000000000000LL; }
  __syncthreads();
  const float* simq = a.sim + (size_t)qi * T;
  if (!L2) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) tab[i] = simq[i];
    __syncthreads();
  }
  for (int p = split; p < a.nprobe; p += a.nsplit) {
    const long long l = a.probes[(size_t)qi * a.nprobe + p];
    if (l < 0) continue;
    const int len = a.list_len[l];
    if (len <= 0) continue;
    const float dis0 = a.coarse[(size_t)qi * a.nprobe + p];
    if (L2) {  // fvec_madd(n, precomputed, -2, sim, tab)
      __syncthreads();
      const float* pl = a.pre + (size_t)l * T;
      for (int i = threadIdx.x; i < T; i += blockDim.x) tab[i] = __fadd_rn(pl[i], __fmul_rn(-2.0f, simq[i]));
      __syncthreads();
    }
    const long long base = a.list_off[l];
    for (int r0 = 0; r0 < len; r0 += blockDim.x) {
      sel.maybe_prune(blockDim.x);
      const int r = r0 + threadIdx.x;
      if (r < len) {
        const long long row = base + r;
        const long long id = a.ids[row];
        if (id >= 0 && filter_pass(a.filt, id)) {
          const unsigned char* code = a.codes + (size_t)row * M;
          float dis = dis0;
          if ((M & 15) == 0 && M <= 128) {
            // all code words of the row first (<= 8 independent 16-byte loads in flight), then the strictly ordered LUT sums
            uint4 cw[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) if (g * 16 < M) cw[g] = *reinterpret_cast<const uint4*>(code + g * 16);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g * 16 < M) {
                const uint32_t w[4] = {cw[g].x, cw[g].y, cw[g].z, cw[g].w};
#pragma unroll
                for (int j = 0; j < 16; ++j) dis = __fadd_rn(dis, tab[(g * 16 + j) * KSUB + ((w[j >> 2] >> (8 * (j & 3))) & 0xff)]);
              }
            }
          } else if ((M & 15) == 0) {
            for (int m0 = 0; m0 < M; m0 += 16) {
              const uint4 c = *reinterpret_cast<const uint4*>(code + m0);
              const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
              for (int j = 0; j < 16; ++j) dis = __fadd_rn(dis, tab[(m0 + j) * KSUB + ((w[j >> 2] >> (8 * (j & 3))) & 0xff)]);
            }
          } else {
            for (int m = 0; m < M; ++m) dis = __fadd_rn(dis, tab[m * KSUB + code[m]]);
          }
          const uint32_t key = f2ord(L2 ? dis : -dis);
          if (sel.passes(key, id)) sel.push(key, id);
        }
      }
    }
  }
  sel.prune();
  const int have = *sel.count;
  uint32_t* okd = a.ws_kd + ((size_t)qi * a.nsplit + split) * a.k;
  long long* oki = a.ws_kid + ((size_t)qi * a.nsplit + split) * a.k;
  for (int i = threadIdx.x; i < a.k; i += blockDim.x) {
    okd[i] = i < have ? sel.kd[i] : KEY_SENTINEL_D;
    oki[i] = i < have ? sel.kid[i] : KEY_SENTINEL_ID;
  }
}

}  // namespace

struct IvfPqIndex : IndexBase {
  enum Mode { kNone, kFlat, kIvfPq } mode = kNone;
  int nlist, M, nbits;
  std::unique_ptr<IndexBase> flat;  // inner Flat index when there is too little data to train (ivf_pq.cc:339-353)
  DevBuf<float> centroids, codebooks, pre;
  DevBuf<long long> cent_ids;
  DevBuf<unsigned char> codes;
  DevBuf<long long> ids;
  IvfLists L;

  IvfPqIndex(b200vs_metric m, int d, const b200vs_params& p) : IndexBase(B200VS_IVF_PQ, m, d, p) {
    nlist = p.nlist > 0 ? p.nlist : 2048;      // Constant::kCreateIvfPqParamNcentroids
    M = p.pq_m > 0 ? p.pq_m : 64;              // kCreateIvfPqParamNsubvector
    nbits = p.pq_nbits > 0 ? p.pq_nbits : 8;   // kCreateIvfPqParamNbitsPerIdx
    if (nbits != 8) fail(B200VS_EILLEGAL_PARAMETERS, "only nbits_per_idx = 8 is implemented");
    if (d % M != 0) fail(B200VS_EILLEGAL_PARAMETERS, "dimension must be divisible by nsubvector");
  }
  bool is_trained() const override { return mode != kNone; }
  int sub_type() const override { return mode == kFlat ? (int)B200VS_FLAT : mode == kIvfPq ? (int)B200VS_IVF_PQ : -1; }  // ivf_pq.cc:474

  void install(const float* h_cent, const float* h_cb) {
    quiesce();
    centroids.free(); codebooks.free(); pre.free(); cent_ids.free();
    centroids.reserve((size_t)nlist * dim, 0, stream);
    codebooks.reserve((size_t)M * KSUB * (dim / M), 0, stream);
    cent_ids.reserve(nlist, 0, stream);
    B200VS_CUDA(cudaMemcpyAsync(centroids.p, h_cent, (size_t)nlist * dim * 4, cudaMemcpyHostToDevice, stream));
    B200VS_CUDA(cudaMemcpyAsync(codebooks.p, h_cb, (size_t)M * KSUB * (dim / M) * 4, cudaMemcpyHostToDevice, stream));
    launch_iota(cent_ids.p, nlist, stream);
    if (metric == B200VS_L2) {
      pre.reserve((size_t)nlist * M * KSUB, 0, stream);
      pq_precompute_kernel<<<nlist, 256, 0, stream>>>(centroids.p, codebooks.p, dim, M, pre.p);
      B200VS_CUDA(cudaGetLastError());
    }
    B200VS_CUDA(cudaStreamSynchronize(stream));
    L.init(nlist, stream);
    codes.free(); ids.free();
    mode = kIvfPq;
  }

  // blob: int64 hdr[6] = {magic 'IVPQ', nlist, dim, metric, M, nbits}; centroids[nlist*dim]; codebooks[M*256*dsub]
  void set_state(const void* blob, size_t len) override {
    std::unique_lock<std::shared_mutex> wl(rw);
    std::lock_guard<std::mutex> gl(gpu_mu);
    set_device();
    if (len < 48) fail(B200VS_EILLEGAL_PARAMETERS, "state blob too short");
    const int64_t* hdr = (const int64_t*)blob;
    if (hdr[0] != 0x51505649 || hdr[2] != dim || hdr[4] != M || hdr[5] != 8) fail(B200VS_EILLEGAL_PARAMETERS, "bad IVF-PQ state blob");
    nlist = (int)hdr[1];
    const size_t nc = (size_t)nlist * dim, ncb = (size_t)M * KSUB * (dim / M);
    if (len < 48 + (nc + ncb) * 4) fail(B200VS_EILLEGAL_PARAMETERS, "state blob truncated");
    const float* f = (const float*)((const char*)blob + 48);
    install(f, f + nc);
  }
  int64_t get_state(void* blob, size_t cap) override {
    RwSharedGuard rl(this);
    if (mode != kIvfPq) return 0;
    const size_t nc = (size_t)nlist * dim, ncb = (size_t)M * KSUB * (dim / M);
    const size_t need = 48 + (nc + ncb) * 4;
    if (!blob || cap < need) return (int64_t)need;
    set_device();
    int64_t hdr[6] = {0x51505649, nlist, dim, (int64_t)metric, M, 8};
    memcpy(blob, hdr, 48);
    B200VS_CUDA(cudaMemcpy((char*)blob + 48, centroids.p, nc * 4, cudaMemcpyDeviceToHost));
    B200VS_CUDA(cudaMemcpy((char*)blob + 48 + nc * 4, codebooks.p, ncb * 4, cudaMemcpyDeviceToHost));
    return (int64_t)need;
  }

  void assign_dev(const float* cd, int kk, const long long* cid, const float* x_dev, int dd, bool l2, int64_t n, long long* out) {
    ScanJob j;
    j.l2 = l2; j.vecs = cd; j.ids = cid; j.d = dd; j.mode = 0; j.n = kk;
    const int64_t chunk = 32768;
    for (int64_t a = 0; a < n; a += chunk) {
      const int64_t m = std::min(chunk, n - a);
      const auto mark = scratch.mark();
      run_scan(this, j, m, x_dev + (size_t)a * dd, 1, nullptr, nullptr, out + a, nullptr, stream);
      scratch.release(mark);
    }
  }

  // VectorIndexIvfPq::Train (ivf_pq.cc:327-395) + VectorIndexRawIvfPq::Train (raw_ivf_pq.cc:457-500)
  void train(int64_t n, const float* x) override {
    if (n <= 0) fail(B200VS_EILLEGAL_PARAMETERS, "data size invalid");
    {
      std::unique_lock<std::shared_mutex> wl(rw);
      if (mode != kNone) return;
      const int64_t need = std::max<int64_t>(256LL * nlist, 256LL * KSUB);
      if (n < need) {  // inner index = Flat
        flat.reset(make_flat(metric, dim, params));
        mode = kFlat;
        return;
      }
    }
    std::unique_lock<std::shared_mutex> wl(rw);
    if (mode != kNone) return;  // another caller trained while the lock was released
    std::lock_guard<std::mutex> gl(gpu_mu);
    set_device();
    quiesce();
    scratch.reset(stream);
    const bool l2 = metric == B200VS_L2;
    DevBuf<long long> cid;
    // level-1 quantiser (niter 10)
    std::vector<float> cent;
    kmeans_gpu(this, metric, dim, n, x, nlist, 10, 256, 1234, cent,
               [&](const float* xd, int64_t m, const float* cd, int kk, long long* out) { assign_dev(cd, kk, cid.p, xd, dim, l2, m, out); },
               [&](int kk) { cid.free(); cid.reserve(kk, 0, stream); launch_iota(cid.p, kk, stream); });
    // PQ training set: <= 256 * ksub vectors (fvecs_maybe_subsample, seed 1234), residuals to their centroid
    int64_t m = std::min<int64_t>(n, 256LL * KSUB);
    std::vector<int64_t> perm(n);
    for (int64_t i = 0; i < n; ++i) perm[i] = i;
    if (m < n) {
      std::mt19937 mt(1234);
      for (int64_t i = 0; i + 1 < n; ++i) { int64_t i2 = i + (int64_t)(mt() % (unsigned long)(n - i)); std::swap(perm[i], perm[i2]); }
    }
    std::vector<float> sub((size_t)m * dim);
    for (int64_t i = 0; i < m; ++i) memcpy(&sub[(size_t)i * dim], x + (size_t)perm[i] * dim, (size_t)dim * 4);
    DevBuf<float> xd, cd;
    DevBuf<long long> asg;
    xd.reserve((size_t)m * dim, 0, stream); cd.reserve((size_t)nlist * dim, 0, stream); asg.reserve(m, 0, stream);
    B200VS_CUDA(cudaMemcpyAsync(xd.p, sub.data(), (size_t)m * dim * 4, cudaMemcpyHostToDevice, stream));
    B200VS_CUDA(cudaMemcpyAsync(cd.p, cent.data(), (size_t)nlist * dim * 4, cudaMemcpyHostToDevice, stream));
    if (metric == B200VS_COSINE) launch_normalize_faiss(xd.p, m, dim, stream);
    cid.free(); cid.reserve(nlist, 0, stream); launch_iota(cid.p, nlist, stream);
    assign_dev(cd.p, nlist, cid.p, xd.p, dim, l2, m, asg.p);
    std::vector<long long> h_asg(m);
    B200VS_CUDA(cudaMemcpyAsync(h_asg.data(), asg.p, (size_t)m * 8, cudaMemcpyDeviceToHost, stream));
    B200VS_CUDA(cudaMemcpyAsync(sub.data(), xd.p, (size_t)m * dim * 4, cudaMemcpyDeviceToHost, stream));
    B200VS_CUDA(cudaStreamSynchronize(stream));
    for (int64_t i = 0; i < m; ++i) {
      const float* c = &cent[(size_t)h_asg[i] * dim];
      float* r = &sub[(size_t)i * dim];
      for (int j = 0; j < dim; ++j) r[j] -= c[j];
    }
    // one k-means per sub-space (ksub centroids, 25 iterations, L2)
    const int dsub = dim / M;
    std::vector<float> cb((size_t)M * KSUB * dsub), slice((size_t)m * dsub), cbm;
    for (int mm = 0; mm < M; ++mm) {
      for (int64_t i = 0; i < m; ++i) memcpy(&slice[(size_t)i * dsub], &sub[(size_t)i * dim + mm * dsub], (size_t)dsub * 4);
      kmeans_gpu(this, B200VS_L2, dsub, m, slice.data(), KSUB, 25, 256, 1234, cbm,
                 [&](const float* xs, int64_t mr, const float* cs, int kk, long long* out) { assign_dev(cs, kk, cid.p, xs, dsub, true, mr, out); },
                 [&](int kk) { cid.free(); cid.reserve(kk, 0, stream); launch_iota(cid.p, kk, stream); });
      memcpy(&cb[(size_t)mm * KSUB * dsub], cbm.data(), (size_t)KSUB * dsub * 4);
    }
    install(cent.data(), cb.data());
  }

  void add(int64_t n, const float* x, const int64_t* in_ids, bool upsert) override {
    if (mode == kFlat) { flat->add(n, x, in_ids, upsert); return; }
    std::unique_lock<std::shared_mutex> wl(rw);
    if (mode == kNone) fail(B200VS_EVECTOR_NOT_TRAIN, "not train");
    std::lock_guard<std::mutex> gl(gpu_mu);
    set_device();
    quiesce();
    scratch.reset(stream);
    if (upsert) remove_locked(n, in_ids);
    float* st = scratch.alloc<float>((size_t)n * dim);
    long long* st_ids = scratch.alloc<long long>(n);
    long long* st_list = scratch.alloc<long long>(n);
    unsigned char* st_codes = scratch.alloc<unsigned char>((size_t)n * M);
    B200VS_CUDA(cudaMemcpyAsync(st, x, (size_t)n * dim * 4, cudaMemcpyHostToDevice, stream));
    B200VS_CUDA(cudaMemcpyAsync(st_ids, in_ids, (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    if (metric == B200VS_COSINE) launch_normalize_faiss(st, n, dim, stream);
    assign_dev(centroids.p, nlist, cent_ids.p, st, dim, metric == B200VS_L2, n, st_list);
    const int wpb = 4;
    pq_encode_kernel<<<(unsigned)cdiv(n, wpb), wpb * 32, (size_t)wpb * dim * 4, stream>>>(st, st_list, centroids.p, codebooks.p, n, dim, M, st_codes);
    B200VS_CUDA(cudaGetLastError());
    std::vector<long long> h_list(n);
    B200VS_CUDA(cudaMemcpyAsync(h_list.data(), st_list, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
    B200VS_CUDA(cudaStreamSynchronize(stream));
    append_encoded(n, st_codes, st_ids, in_ids, h_list.data());
  }

  // rows already encoded (device codes [n, M], device ids) go to the tails of their lists; callers hold the write lock
  void append_encoded(int64_t n, const unsigned char* d_codes, const long long* d_ids, const int64_t* h_ids_in, const long long* h_list) {
    std::vector<long long> slots(n);
    long long* st_slots = scratch.alloc<long long>(n);
    std::vector<int> need(nlist, 0);
    for (int64_t i = 0; i < n; ++i) need[h_list[i]]++;
    L.reserve_for(need, [&](int64_t rows) {
      codes.reserve((size_t)rows * M, (size_t)L.arena_used_before * M, stream);
      ids.reserve((size_t)rows, (size_t)L.arena_used_before, stream);
    }, [&](int64_t src, int64_t dst, int64_t len) {
      B200VS_CUDA(cudaMemcpyAsync(codes.p + (size_t)dst * M, codes.p + (size_t)src * M, (size_t)len * M, cudaMemcpyDeviceToDevice, stream));
      B200VS_CUDA(cudaMemcpyAsync(ids.p + dst, ids.p + src, (size_t)len * 8, cudaMemcpyDeviceToDevice, stream));
    });
    for (int64_t i = 0; i < n; ++i) slots[i] = L.append((int)h_list[i], h_ids_in[i]);
    B200VS_CUDA(cudaMemcpyAsync(st_slots, slots.data(), (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    scatter_codes_kernel<<<(unsigned)cdiv(n * M, 256), 256, 0, stream>>>(d_codes, d_ids, st_slots, n, M, codes.p, ids.p);
    B200VS_CUDA(cudaGetLastError());
    L.upload(stream);
    B200VS_CUDA(cudaStreamSynchronize(stream));
  }

  int64_t remove_locked(int64_t n, const int64_t* del) {
    std::vector<int64_t> rws;
    L.remove_ids(n, del, rws);
    if (!rws.empty()) {
      long long* d_rows = scratch.alloc<long long>(rws.size());
      B200VS_CUDA(cudaMemcpyAsync(d_rows, rws.data(), rws.size() * 8, cudaMemcpyHostToDevice, stream));
      launch_set_ids(ids.p, d_rows, (int64_t)rws.size(), -1, stream);
      B200VS_CUDA(cudaStreamSynchronize(stream));
    }
    return (int64_t)rws.size();
  }
  void maybe_compact() {
    if (!L.needs_compaction()) return;
    std::vector<long long> src, dst;
    const int64_t new_rows = std::max<int64_t>(1, L.plan_compaction(src, dst));
    DevBuf<unsigned char> nc; DevBuf<long long> ni;
    nc.reserve((size_t)new_rows * M, 0, stream); ni.reserve(new_rows, 0, stream);
    const int64_t m = (int64_t)src.size();
    if (m) {
      long long* d_src = scratch.alloc<long long>(m);
      long long* d_dst = scratch.alloc<long long>(m);
      B200VS_CUDA(cudaMemcpyAsync(d_src, src.data(), m * 8, cudaMemcpyHostToDevice, stream));
      B200VS_CUDA(cudaMemcpyAsync(d_dst, dst.data(), m * 8, cudaMemcpyHostToDevice, stream));
      move_codes_kernel<<<(unsigned)cdiv(m * M, 256), 256, 0, stream>>>(codes.p, ids.p, d_src, d_dst, m, M, nc.p, ni.p);
      B200VS_CUDA(cudaGetLastError());
    }
    B200VS_CUDA(cudaStreamSynchronize(stream));
    std::swap(codes.p, nc.p); std::swap(codes.cap, nc.cap);
    std::swap(ids.p, ni.p); std::swap(ids.cap, ni.cap);
    L.commit_compaction();
    L.upload(stream);
  }
  int64_t remove(int64_t n, const int64_t* del) override {
    if (mode == kFlat) { const int64_t r = flat->remove(n, del); return r == 0 ? -1 : r; }  // Flat semantics: unknown ids are OK
    std::unique_lock<std::shared_mutex> wl(rw);
    if (mode == kNone) return -1;
    std::lock_guard<std::mutex> gl(gpu_mu);
    set_device();
    quiesce();
    scratch.reset(stream);
    const int64_t r = remove_locked(n, del);
    maybe_compact();
    return r;
  }

  void search_dev(int64_t nq, const float* xq, int k, const SearchCtx& sc, float* od, long long* oi, cudaStream_t s) override {
    if (mode == kNone) { fill_empty_results(nq, k, od, oi, s); return; }  // ivf_pq.cc:159-163
    if (mode == kFlat) {  // delegate to the inner Flat index, sharing this call's stream and scratch discipline
      std::shared_lock<std::shared_mutex> rl(flat->rw);
      LaneGuard lg(flat.get(), s);
      flat->search_dev(nq, xq, k, sc, od, oi, s);  // (an id-list filter lives in the outer lane's scratch: same stream, still locked)
      for (int i = 0; i < 8; ++i) stats[i] = flat->stats[i].load();
      return;
    }
    scan_codes(nq, xq, k, sc, false, 0.f, od, oi, nullptr, s);
  }

  // coarse quantiser -> per-query LUT -> code scan + select -> merge.  Range mode (has_thr): keep hits strictly inside the
  // raw threshold (L2: dis < thr; IP: dis > thr), at most k of them, closest first; out_counts = hits per query.
  void scan_codes(int64_t nq, const float* xq, int k, const SearchCtx& sc, bool has_thr, float thr_raw, float* od, long long* oi, int* oc,
                  cudaStream_t s) {
    const float* q = prepare_queries(nq, xq, s);
    const bool l2 = metric == B200VS_L2;
    int nprobe = sc.nprobe > 0 ? sc.nprobe : 80;  // Constant::kSearchIvfPqParamNprobe, raw_ivf_pq.cc:170
    nprobe = std::min(nprobe, nlist);              // raw_ivf_pq.cc:190
    // coarse quantiser: top-nprobe lists + raw coarse values
    ScanJob cj;
    cj.l2 = l2; cj.vecs = centroids.p; cj.ids = cent_ids.p; cj.d = dim; cj.mode = 0; cj.n = nlist;
    long long* probes = scratch.alloc<long long>((size_t)nq * nprobe);
    float* coarse = scratch.alloc<float>((size_t)nq * nprobe);
    run_scan(this, cj, nq, q, nprobe, nullptr, coarse, probes, nullptr, s);
    if (profiling) profile_probed(this, probes, nq * nprobe, nlist, L.d_len.p, s);
    const int T = M * KSUB;
    float* sim = scratch.alloc<float>((size_t)nq * T);
    pq_sim_kernel<<<(unsigned)nq, 256, 0, s>>>(q, codebooks.p, nq, dim, M, sim);
    PqScanArgs a;
    a.codes = codes.p; a.ids = ids.p; a.probes = probes; a.coarse = coarse; a.list_off = L.d_off.p; a.list_len = L.d_len.p;
    a.sim = sim; a.pre = l2 ? pre.p : nullptr; a.M = M; a.nprobe = nprobe; a.k = k;
    a.nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(nprobe, (148 * 2 + nq - 1) / nq));
    a.pool_cap = select_pool_cap(k, PQ_THREADS);
    a.has_thr = has_thr ? 1 : 0;
    a.thr_key = has_thr ? f2ord(l2 ? thr_raw : -thr_raw) : 0;
    a.ws_kd = scratch.alloc<uint32_t>((size_t)nq * a.nsplit * k);
    a.ws_kid = scratch.alloc<long long>((size_t)nq * a.nsplit * k);
    a.filt.has_range = sc.has_range; a.filt.negate = sc.negate; a.filt.rmin = sc.rmin; a.filt.rmax = sc.rmax;
    a.filt.sorted_ids = sc.sorted_ids_dev; a.filt.n_ids = sc.n_ids;
    const size_t smem = (size_t)T * 4 + BlockSelect::smem_bytes(a.pool_cap);
    if (smem > 227 * 1024) fail(B200VS_EILLEGAL_PARAMETERS, "nsubvector / topk too large for the shared-memory LUT");
    const size_t smem2 = BlockSelect::smem_bytes(a.pool_cap);
    dim3 grid(a.nsplit, (unsigned)nq);
    ScopedKernelTimer timer(this, s, profiling);
    if (l2) {
      B200VS_CUDA(cudaFuncSetAttribute(pq_scan_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      pq_scan_select_kernel<true><<<grid, PQ_THREADS, smem, s>>>(a);
      timer.stop();
      B200VS_CUDA(cudaFuncSetAttribute(merge_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      merge_select_kernel<true><<<(unsigned)nq, SCAN_THREADS, smem2, s>>>(a.ws_kd, a.ws_kid, a.nsplit, k, a.pool_cap, od, nullptr, oi, oc, nullptr, nullptr);
    } else {
      B200VS_CUDA(cudaFuncSetAttribute(pq_scan_select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      pq_scan_select_kernel<false><<<grid, PQ_THREADS, smem, s>>>(a);
      timer.stop();
      B200VS_CUDA(cudaFuncSetAttribute(merge_select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      merge_select_kernel<false><<<(unsigned)nq, SCAN_THREADS, smem2, s>>>(a.ws_kd, a.ws_kid, a.nsplit, k, a.pool_cap, od, nullptr, oi, oc, nullptr, nullptr);
    }
    B200VS_CUDA(cudaGetLastError());
    launch_count(3);
  }

  // VectorIndexRawIvfPq::RangeSearch, vector_index_raw_ivf_pq.cc:212-278: faiss IndexIVFPQ::range_search on the codes —
  // radius -> 1 - radius for IP / cosine (:229-232), L2 keeps dis < radius, IP keeps dis > radius; the same LUT scan with a
  // threshold epilogue instead of the k-th-best threshold.  Untrained -> empty results (:219-222).
  void range_search_dev(int64_t nq, const float* xq, float radius, int max_results, const SearchCtx& sc, float* od, long long* oi,
                        int* oc, cudaStream_t s) override {
    if (mode == kNone) {
      fill_empty_results(nq, max_results, od, oi, s);
      if (oc) B200VS_CUDA(cudaMemsetAsync(oc, 0, (size_t)nq * 4, s));
      return;
    }
    if (mode == kFlat) {
      std::shared_lock<std::shared_mutex> rl(flat->rw);
      LaneGuard lg(flat.get(), s);
      flat->range_search_dev(nq, xq, radius, max_results, sc, od, oi, oc, s);
      return;
    }
    scan_codes(nq, xq, max_results, sc, true, ip_like() ? 1.0F - radius : radius, od, oi, oc, s);
  }

  int64_t count() const override { return mode == kFlat ? flat->count() : L.live; }
  int64_t deleted_count() const override { return mode == kFlat ? flat->deleted_count() : L.dead; }
  int64_t memory_size() const override {
    if (mode == kFlat) return flat->memory_size();
    return (int64_t)(codes.cap + ids.cap * 8 + centroids.cap * 4 + codebooks.cap * 4 + pre.cap * 4);  // raw_ivf_pq.cc:440-450
  }
  // Save / Load (reference: faiss::write_index / read_index of the IndexIVFPQ, vector_index_raw_ivf_pq.cc:298 and :314): own
  // container "B2VSPQ01" = mode, then either the inner Flat index's rows or {trained-state blob, list offsets, ids,
  // PQ codes in list-major order}.  Codes are restored as stored (never re-encoded: the vectors are gone).
  void save(const std::string& path) override {
    RwSharedGuard hold(this);  // one reader hold across count / trained state / export
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) fail(B200VS_EINTERNAL, "cannot open " + path);
    auto wr = [&](const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) fail(B200VS_EINTERNAL, "short write"); };
    try {
      int64_t hdr[8] = {0, (int64_t)mode, (int64_t)metric, dim, nlist, M, 0, 0};
      memcpy(hdr, "B2VSPQ01", 8);
      if (mode == kFlat) {
        const int64_t n = flat->count();
        std::vector<int64_t> off(2, 0), fids((size_t)n);
        std::vector<float> vec((size_t)n * dim);
        if (n) flat->export_lists(off.data(), vec.data(), nullptr, fids.data());
        hdr[6] = n;
        wr(hdr, sizeof(hdr)); wr(fids.data(), fids.size() * 8); wr(vec.data(), vec.size() * 4);
      } else if (mode == kIvfPq) {
        const int64_t st_len = get_state(nullptr, 0);
        std::vector<unsigned char> st((size_t)st_len);
        get_state(st.data(), st.size());
        const int64_t n = count();
        std::vector<int64_t> off((size_t)nlist + 1, 0), fids((size_t)n);
        std::vector<uint8_t> cds((size_t)n * M);
        export_lists(off.data(), nullptr, cds.data(), fids.data());
        hdr[6] = n; hdr[7] = st_len;
        wr(hdr, sizeof(hdr)); wr(st.data(), st.size()); wr(off.data(), off.size() * 8); wr(fids.data(), fids.size() * 8); wr(cds.data(), cds.size());
      } else {
        wr(hdr, sizeof(hdr));  // untrained: nothing but the header
      }
    } catch (...) { fclose(f); throw; }
    fclose(f);
  }
  void load(const std::string& path) override {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) fail(B200VS_EINTERNAL, "cannot open " + path);
    auto rd = [&](void* p, size_t n) { if (n && fread(p, 1, n, f) != n) fail(B200VS_EINTERNAL, "short read / truncated index file"); };
    try {
      int64_t hdr[8];
      rd(hdr, sizeof(hdr));
      if (memcmp(hdr, "B2VSPQ01", 8) != 0 || hdr[2] != (int64_t)metric || hdr[3] != dim || hdr[5] != M)
        fail(B200VS_EINTERNAL, "index file does not match this index (type / metric / dimension / nsubvector)");
      if (mode != kNone) fail(B200VS_EINTERNAL, "load into a trained index");
      const int64_t n = hdr[6];
      if (hdr[1] == kFlat) {
        std::vector<int64_t> fids((size_t)n);
        std::vector<float> vec((size_t)n * dim);
        rd(fids.data(), fids.size() * 8); rd(vec.data(), vec.size() * 4);
        { std::unique_lock<std::shared_mutex> wl(rw); flat.reset(make_flat(metric, dim, params)); mode = kFlat; }
        flat->loading = true;  // rows come back exactly as stored (already normalised for cosine)
        try {
          for (int64_t a = 0; a < n; a += 32768) flat->add(std::min<int64_t>(32768, n - a), vec.data() + (size_t)a * dim, fids.data() + a, false);
        } catch (...) { flat->loading = false; throw; }
        flat->loading = false;
      } else if (hdr[1] == kIvfPq) {
        std::vector<unsigned char> st((size_t)hdr[7]);
        rd(st.data(), st.size());
        set_state(st.data(), st.size());  // centroids, codebooks, precomputed table; nlist from the blob
        std::vector<int64_t> off((size_t)nlist + 1), fids((size_t)n);
        std::vector<uint8_t> cds((size_t)n * M);
        rd(off.data(), off.size() * 8); rd(fids.data(), fids.size() * 8); rd(cds.data(), cds.size());
        if (off[nlist] != n) fail(B200VS_EINTERNAL, "corrupt index file (list offsets)");
        std::unique_lock<std::shared_mutex> wl(rw);
        std::lock_guard<std::mutex> gl(gpu_mu);
        set_device();
        quiesce();
        std::vector<long long> h_list((size_t)n);
        for (int l = 0; l < nlist; ++l) for (int64_t i = off[l]; i < off[l + 1]; ++i) h_list[(size_t)i] = l;
        for (int64_t a = 0; a < n; a += 262144) {
          const int64_t m = std::min<int64_t>(262144, n - a);
          scratch.reset(stream);
          unsigned char* d_codes = scratch.alloc<unsigned char>((size_t)m * M);
          long long* d_ids = scratch.alloc<long long>(m);
          B200VS_CUDA(cudaMemcpyAsync(d_codes, cds.data() + (size_t)a * M, (size_t)m * M, cudaMemcpyHostToDevice, stream));
          B200VS_CUDA(cudaMemcpyAsync(d_ids, fids.data() + a, (size_t)m * 8, cudaMemcpyHostToDevice, stream));
          append_encoded(m, d_codes, d_ids, fids.data() + a, h_list.data() + a);
        }
      }
    } catch (...) { fclose(f); throw; }
    fclose(f);
  }

  void export_lists(int64_t* list_off, float* vectors, uint8_t* out_codes, int64_t* out_ids) override {
    if (mode == kFlat) { flat->export_lists(list_off, vectors, out_codes, out_ids); return; }
    RwSharedGuard rl(this);
    std::lock_guard<std::mutex> gl(gpu_mu);
    set_device();
    std::vector<unsigned char> buf;
    int64_t o = 0;
    for (int l = 0; l < nlist; ++l) {
      if (list_off) list_off[l] = o;
      const auto& m = L.lists[l];
      if (m.len == 0) continue;
      if (out_codes) {
        buf.resize((size_t)m.len * M);
        B200VS_CUDA(cudaMemcpy(buf.data(), codes.p + (size_t)m.off * M, (size_t)m.len * M, cudaMemcpyDeviceToHost));
      }
      for (int p = 0; p < m.len; ++p) {
        const int64_t id = L.h_ids[m.off + p];
        if (id < 0) continue;
        if (out_ids) out_ids[o] = id;
        if (out_codes) memcpy(out_codes + (size_t)o * M, buf.data() + (size_t)p * M, (size_t)M);
        ++o;
      }
    }
    if (list_off) list_off[nlist] = o;
  }
};

IndexBase* make_ivf_pq(b200vs_metric m, int d, const b200vs_params& p) { return new IvfPqIndex(m, d, p); }

}  // namespace b200vs
