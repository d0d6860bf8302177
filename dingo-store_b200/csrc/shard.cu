// shard.cu — list-sharded multi-GPU deployment behind the C ABI (b200vs_shard_*, include/b200vs.h).
//
// One process per GPU.  Every rank holds the replicated centroid table of ONE logical IVF-Flat index and the rows of the
// inverted lists it owns (contiguous blocks of lists per rank — equivalently Raft regions mapped to GPUs,
// src/vector/vector_index.h:54-55).  A batched search is:
//   1. coarse quantiser on this rank's SLICE of the batch (all ranks hold all queries), one all-gather of the probe table
//      (the faiss quantizer->search step of IndexIVF::search, src/vector/vector_index_ivf_flat.cc:247-251);
//   2. tile scan of the probed lists this rank owns (csrc/tc_scan.cu) -> exact local top-k;
//   3. ONE ncclAllGather of the packed per-shard top-k ((distance, id) 16-byte records) over NVLink, then the k-way merge
//      kernel — the engine's analogue of VectorIndexWrapper::MergeSearchResults (src/vector/vector_index.cc:1056-1108).
// Rows are routed to their list owner at add time (b200vs_shard_add*: assignment on the sender, grouped ncclSend/ncclRecv).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a single-GPU dingo-store node needs no NCCL at all, and a process
// that already carries an NCCL (e.g. the one bundled with PyTorch in the test harness) shares it.  Each in-flight lane owns
// its communicator (ncclCommSplit), so batches in flight overlap across streams; callers order batches with a sequence
// number, which maps every batch to the same communicator on every rank.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <memory>
#include <vector>

#include "index.h"

namespace b200vs {

// ---------------------------------------------------------------------------------------------
// NCCL, resolved lazily
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommSplit) CommSplit = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (api.h) return api;
  const char* env = getenv("B200VS_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) fail(B200VS_EVECTOR_NOT_SUPPORT, "NCCL not found (libnccl.so.2): list-sharded search needs it; set B200VS_NCCL_LIB");
#define B200VS_SYM(field, name)                                                                 \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                            \
  if (!api.field) fail(B200VS_EVECTOR_NOT_SUPPORT, std::string("NCCL symbol missing: ") + name)
  B200VS_SYM(GetUniqueId, "ncclGetUniqueId");
  B200VS_SYM(CommInitRank, "ncclCommInitRank");
  B200VS_SYM(CommSplit, "ncclCommSplit");
  B200VS_SYM(CommDestroy, "ncclCommDestroy");
  B200VS_SYM(AllGather, "ncclAllGather");
  B200VS_SYM(AllReduce, "ncclAllReduce");
  B200VS_SYM(Broadcast, "ncclBroadcast");
  B200VS_SYM(Send, "ncclSend");
  B200VS_SYM(Recv, "ncclRecv");
  B200VS_SYM(GroupStart, "ncclGroupStart");
  B200VS_SYM(GroupEnd, "ncclGroupEnd");
  B200VS_SYM(GetErrorString, "ncclGetErrorString");
  B200VS_SYM(GetVersion, "ncclGetVersion");
#undef B200VS_SYM
  api.h = h;
  return api;
}

#define B200VS_NCCL(expr)                                                                                          \
  do {                                                                                                             \
    ncclResult_t _r = (expr);                                                                                      \
    if (_r != ncclSuccess && _r != ncclInProgress)                                                                 \
      fail(B200VS_EINTERNAL, std::string(#expr) + " failed: " + nccl().GetErrorString(_r));                        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// kernels: pack / merge of 16-byte (distance, id) records, owner routing
// ---------------------------------------------------------------------------------------------
struct __align__(16) TopkRec {
  float dist;
  int pad;
  long long id;
};
static_assert(sizeof(TopkRec) == 16, "one 16-byte record per hit");

static __global__ void pack_topk_kernel(const float* __restrict__ d, const long long* __restrict__ id, long long n, TopkRec* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  TopkRec r;
  r.dist = d[i]; r.pad = 0; r.id = id[i];
  out[i] = r;
}

// parts [world, nq, k] records (ascending per part, id < 0 = empty) -> merged [nq, k]: (distance, id) ascending.
// One warp per query: world * k <= 1024 records, each lane ranks its records against all others... for the usual
// world * k <= 256 a rank sort in shared memory is cheapest.
constexpr int MERGE_WARPS = 4;
constexpr int MERGE_MAX = 256;
static __global__ void __launch_bounds__(MERGE_WARPS * 32)
merge_packed_warp_kernel(const TopkRec* __restrict__ parts, int world, long long nq, int k, float* out_dist, long long* out_ids) {
  __shared__ uint32_t s_kd[MERGE_WARPS][MERGE_MAX];
  __shared__ long long s_id[MERGE_WARPS][MERGE_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * MERGE_WARPS + warp;
  if (q >= nq) return;
  const int tot = world * k;
  int m = 0;
  for (int base = 0; base < tot; base += 32) {
    const int i = base + lane;
    TopkRec r;
    r.id = -1; r.dist = 0.f;
    if (i < tot) r = parts[((size_t)(i / k) * nq + q) * k + (i % k)];
    const bool ok = r.id >= 0;
    const unsigned msk = __ballot_sync(0xffffffffu, ok);
    if (ok) { const int p = m + __popc(msk & ((1u << lane) - 1u)); s_kd[warp][p] = f2ord(r.dist); s_id[warp][p] = r.id; }
    m += __popc(msk);
  }
  __syncwarp();
  for (int e = lane; e < m; e += 32) {
    const uint32_t d0 = s_kd[warp][e];
    const long long i0 = s_id[warp][e];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const uint32_t dj = s_kd[warp][j];
      const long long ij = s_id[warp][j];
      rank += (key_less(dj, ij, d0, i0) || (dj == d0 && ij == i0 && j < e)) ? 1 : 0;
    }
    if (rank < k) { out_dist[(size_t)q * k + rank] = ord2f(d0); out_ids[(size_t)q * k + rank] = i0; }
  }
  for (int i = min(m, k) + lane; i < k; i += 32) { out_dist[(size_t)q * k + i] = 0.f; out_ids[(size_t)q * k + i] = -1; }
}

static __global__ void unpack_parts_kernel(const TopkRec* __restrict__ parts, long long n, float* d, long long* id) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TopkRec r = parts[i];
  d[i] = r.dist; id[i] = r.id;
}

static __global__ void fill_ll_kernel(long long* p, long long n, long long v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// gather rows by a permutation: dst[i] = src[perm[i]]  (rows, ids, lists)
static __global__ void gather_rows_kernel(const float* __restrict__ x, const long long* __restrict__ ids, const long long* __restrict__ lists,
                                          const long long* __restrict__ perm, long long n, int d, float* ox, long long* oids, long long* olists) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long src = perm[w];
  const float4* s4 = reinterpret_cast<const float4*>(x + (size_t)src * d);
  float4* d4 = reinterpret_cast<float4*>(ox + (size_t)w * d);
  if ((d & 3) == 0) for (int i = lane; i < (d >> 2); i += 32) d4[i] = s4[i];
  else for (int i = lane; i < d; i += 32) ox[(size_t)w * d + i] = x[(size_t)src * d + i];
  if (lane == 0) { oids[w] = ids[src]; olists[w] = lists[src]; }
}

static __global__ void count_lists_kernel(const long long* __restrict__ lists, long long n, int nlist, unsigned long long* counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const long long l = lists[i]; if (l >= 0 && l < nlist) atomicAdd(counts + l, 1ULL); }
}

}  // namespace b200vs

using namespace b200vs;

// ---------------------------------------------------------------------------------------------
// the shard handle
// ---------------------------------------------------------------------------------------------
struct b200vs_shard {
  IndexBase* ix = nullptr;
  int rank = 0, world = 1, nlanes = 1;
  struct SLane {
    ncclComm_t comm = nullptr;
    std::mutex mu;
    DevBuf<long long> probes;   // [per * world, nprobe]
    DevBuf<TopkRec> parts;      // [world, nq, k]
    DevBuf<float> q_all;        // host-pointer variant: [per * world, d]
    DevBuf<float> od;           // local / merged results
    DevBuf<long long> oi;
    cudaStream_t own = nullptr;
    cudaStream_t last = nullptr;  // stream of the lane's previous batch
    cudaEvent_t done = nullptr;   // recorded behind it: a batch on another stream waits for it before touching the buffers
  };
  std::vector<std::unique_ptr<SLane>> lanes;
  std::mutex seq_mu;
  std::condition_variable seq_cv;
  int64_t next_auto = 0;             // sequence numbers handed out when the caller passes seq < 0
  int64_t next_seq = 0;              // the batch whose turn it is to be enqueued
  // write path (collective, one at a time)
  std::mutex add_mu;
  cudaStream_t ws = nullptr;
  DevBuf<unsigned long long> plan_counts;  // [nlist] rows per list seen by b200vs_shard_plan_add*
  bool planning = false;

  ~b200vs_shard() {
    if (ix) cudaSetDevice(ix->device);
    for (auto& l : lanes) {
      if (l->own) { cudaStreamSynchronize(l->own); cudaStreamDestroy(l->own); }
      if (l->done) { cudaEventSynchronize(l->done); cudaEventDestroy(l->done); }
      if (l->comm) nccl().CommDestroy(l->comm);
    }
    if (ws) { cudaStreamSynchronize(ws); cudaStreamDestroy(ws); }
  }
  int lists_per_rank() const { return (ix->nlist_now() + world - 1) / world; }
};

namespace {

template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const StatusError& e) {
    g_last_error = e.msg;
    return e.code;
  } catch (const CudaError& e) {
    g_last_error = e.what();
    return B200VS_EINTERNAL;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return B200VS_EINTERNAL;
  } catch (...) {
    g_last_error = "unknown error";
    return B200VS_EINTERNAL;
  }
}

b200vs_shard* get(b200vs_shard* s) {
  if (!s || !s->ix) fail(B200VS_EILLEGAL_PARAMETERS, "null shard handle");
  return s;
}

// Take the turn of batch `seq` (seq < 0: the next one in call order).  Batches are ENQUEUED strictly in sequence order on
// every rank (the enqueue is asynchronous, so batches still overlap on the GPU): a device-wide synchronisation hidden in
// one batch's enqueue (cudaFree / cudaMalloc of a growing buffer) then only ever waits for kernels of earlier batches,
// which every rank has already enqueued — enqueueing two communicators' collectives in different orders on different ranks
// is the classic multi-communicator deadlock.  Batch seq uses lane / communicator seq % nlanes.
struct SeqLane {
  b200vs_shard* sh;
  b200vs_shard::SLane* lane;
  int64_t seq;
  SeqLane(b200vs_shard* s, int64_t seq_) : sh(s), seq(seq_) {
    std::unique_lock<std::mutex> lk(sh->seq_mu);
    if (seq < 0) seq = sh->next_auto;
    sh->next_auto = std::max(sh->next_auto, seq + 1);
    if (seq < sh->next_seq) fail(B200VS_EILLEGAL_PARAMETERS, "shard search: sequence number already used");
    sh->seq_cv.wait(lk, [&] { return sh->next_seq == seq; });
    lane = sh->lanes[(size_t)(seq % sh->nlanes)].get();
    lane->mu.lock();
  }
  ~SeqLane() {
    lane->mu.unlock();
    {
      std::lock_guard<std::mutex> lk(sh->seq_mu);
      sh->next_seq = seq + 1;
    }
    sh->seq_cv.notify_all();
  }
};

SearchCtx make_ctx_shard(IndexBase* ix, const b200vs_search_params* sp, cudaStream_t s) {
  SearchCtx sc;
  if (!sp) return sc;
  sc.nprobe = sp->nprobe; sc.efsearch = sp->efsearch; sc.exact_only = sp->exact_only;
  sc.has_range = sp->has_range; sc.negate = sp->negate; sc.rmin = sp->range_min; sc.rmax = sp->range_max;
  if (sp->sorted_ids) {
    long long* d = ix->scratch.alloc<long long>((size_t)std::max<int64_t>(sp->n_ids, 1));
    if (sp->n_ids > 0) B200VS_CUDA(cudaMemcpyAsync(d, sp->sorted_ids, (size_t)sp->n_ids * 8, cudaMemcpyHostToDevice, s));
    sc.sorted_ids_dev = d; sc.n_ids = sp->n_ids;
  }
  return sc;
}

// the search proper: queries (all nq of them, raw) are on the device in `xq`; results land in out_* (device) on stream s
void shard_search_stream(b200vs_shard* sh, b200vs_shard::SLane& L, int64_t nq, const float* xq, int k, const b200vs_search_params* sp,
                         float* out_dist, long long* out_ids, cudaStream_t s) {
  IndexBase* ix = sh->ix;
  const int W = sh->world, r = sh->rank, d = ix->dim;
  if (L.last && L.last != s && L.done) B200VS_CUDA(cudaStreamWaitEvent(s, L.done, 0));
  SearchCtx sc = make_ctx_shard(ix, sp, s);
  const int nprobe = ix->resolve_nprobe_api(sc);
  const int64_t per = (nq + W - 1) / W;
  // cosine: normalise once (NormalizeVectorForFaiss), the coarse pass and the list scan both take prepared queries
  const float* q = ix->prepare_queries(nq, xq, s);
  // 1) coarse quantiser on this rank's slice, all-gather of the probe table (in place)
  L.probes.reserve((size_t)per * W * nprobe, 0, s);
  const int64_t q0 = std::min<int64_t>(nq, per * r), q1 = std::min<int64_t>(nq, per * (r + 1));
  long long* mine = L.probes.p + (size_t)per * r * nprobe;
  if (q1 - q0 < per) fill_ll_kernel<<<(unsigned)cdiv(per * nprobe, 256), 256, 0, s>>>(mine, per * nprobe, -1);
  if (q1 > q0) ix->coarse_probes_dev(q1 - q0, q + (size_t)q0 * d, nprobe, mine, s);
  ix->phase(IndexBase::PH_COMM, s);
  if (W > 1) B200VS_NCCL(nccl().AllGather(mine, L.probes.p, (size_t)per * nprobe, ncclInt64, L.comm, s));
  ix->phase(IndexBase::PH_OTHER, s);
  // 2) tile scan of the probed lists this rank owns -> exact local top-k
  L.od.reserve((size_t)nq * k, 0, s);
  L.oi.reserve((size_t)nq * k, 0, s);
  ix->search_probes_prepared_dev(nq, q, k, L.probes.p, nprobe, sc, L.od.p, L.oi.p, s);
  // 3) ONE all-gather of the packed per-shard top-k + merge
  const long long nk = (long long)nq * k;
  L.parts.reserve((size_t)W * nk, 0, s);
  ix->phase(IndexBase::PH_MERGE, s);
  pack_topk_kernel<<<(unsigned)cdiv(nk, 256), 256, 0, s>>>(L.od.p, L.oi.p, nk, L.parts.p + (size_t)r * nk);
  ix->phase(IndexBase::PH_COMM, s);
  if (W > 1) B200VS_NCCL(nccl().AllGather(L.parts.p + (size_t)r * nk, L.parts.p, (size_t)nk * sizeof(TopkRec), ncclUint8, L.comm, s));
  ix->phase(IndexBase::PH_MERGE, s);
  if (W * k <= MERGE_MAX) {
    merge_packed_warp_kernel<<<(unsigned)cdiv(nq, MERGE_WARPS), MERGE_WARPS * 32, 0, s>>>(L.parts.p, W, nq, k, out_dist, out_ids);
  } else {  // wide merges: unpack and use the block merge kernel
    float* pd = ix->scratch.alloc<float>((size_t)W * nk);
    long long* pi = ix->scratch.alloc<long long>((size_t)W * nk);
    unpack_parts_kernel<<<(unsigned)cdiv(W * nk, 256), 256, 0, s>>>(L.parts.p, W * nk, pd, pi);
    launch_merge_api(W, nq, k, pd, pi, out_dist, out_ids, s);
  }
  B200VS_CUDA(cudaGetLastError());
  ix->phase(IndexBase::PH_OTHER, s);
  ix->launch_count(3);
  if (!L.done) B200VS_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
  B200VS_CUDA(cudaEventRecord(L.done, s));
  L.last = s;
}

}  // namespace

extern "C" {

int b200vs_shard_unique_id(uint8_t id[B200VS_SHARD_ID_BYTES]) {
  return guarded([&]() -> int {
    if (!id) fail(B200VS_EILLEGAL_PARAMETERS, "null id");
    static_assert(sizeof(ncclUniqueId) <= B200VS_SHARD_ID_BYTES, "ncclUniqueId must fit the ABI's id blob");
    ncclUniqueId u;
    B200VS_NCCL(nccl().GetUniqueId(&u));
    memset(id, 0, B200VS_SHARD_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return B200VS_OK;
  });
}

int b200vs_shard_create(b200vs_index* idx, int32_t rank, int32_t world, const uint8_t id[B200VS_SHARD_ID_BYTES], int32_t lanes, b200vs_shard** out) {
  return guarded([&]() -> int {
    if (!out) fail(B200VS_EILLEGAL_PARAMETERS, "out is null");
    *out = nullptr;
    IndexBase* ix = index_impl(idx);
    if (ix->type != B200VS_IVF_FLAT) fail(B200VS_EVECTOR_NOT_SUPPORT, "list sharding is implemented for IVF_FLAT (HNSW: replicas only)");
    if (world < 1 || rank < 0 || rank >= world) fail(B200VS_EILLEGAL_PARAMETERS, "bad rank / world");
    if (world > 1 && !id) fail(B200VS_EILLEGAL_PARAMETERS, "null communicator id");
    lanes = std::max(1, std::min(lanes <= 0 ? 2 : lanes, kLanes));
    std::unique_ptr<b200vs_shard> sh(new b200vs_shard());
    sh->ix = ix; sh->rank = rank; sh->world = world; sh->nlanes = lanes;
    ix->set_device();
    B200VS_CUDA(cudaStreamCreateWithFlags(&sh->ws, cudaStreamNonBlocking));
    for (int c = 0; c < lanes; ++c) {
      sh->lanes.emplace_back(new b200vs_shard::SLane());
      B200VS_CUDA(cudaStreamCreateWithFlags(&sh->lanes.back()->own, cudaStreamNonBlocking));
    }
    if (world > 1) {
      ncclUniqueId u;
      memcpy(&u, id, sizeof(u));
      B200VS_NCCL(nccl().CommInitRank(&sh->lanes[0]->comm, world, u, rank));
      for (int c = 1; c < lanes; ++c) B200VS_NCCL(nccl().CommSplit(sh->lanes[0]->comm, 0, rank, &sh->lanes[c]->comm, nullptr));
    }
    *out = sh.release();
    return B200VS_OK;
  });
}

void b200vs_shard_destroy(b200vs_shard* sh) {
  try { delete sh; } catch (...) {}
}

int b200vs_shard_list_range(b200vs_shard* h, int32_t rank, int32_t* begin, int32_t* end) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    if (rank < 0 || rank >= sh->world) fail(B200VS_EILLEGAL_PARAMETERS, "bad rank");
    const int nl = sh->ix->nlist_now(), per = sh->lists_per_rank();
    if (begin) *begin = std::min(nl, per * rank);
    if (end) *end = std::min(nl, per * (rank + 1));
    return B200VS_OK;
  });
}

// Distributed training: every rank clusters ITS training rows into nlist / world centroids (faiss::Clustering shape, as
// b200vs_train), the blocks are all-gathered into the replicated global table.  Build time stays flat as ranks are added.
int b200vs_shard_train(b200vs_shard* h, int64_t n, const float* x) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    IndexBase* ix = sh->ix;
    if (n <= 0 || !x) fail(B200VS_EILLEGAL_PARAMETERS, "data size invalid");
    std::lock_guard<std::mutex> g(sh->add_mu);
    const int W = sh->world, d = ix->dim, nlist = ix->params.nlist > 0 ? ix->params.nlist : 2048;
    if (nlist % W != 0) fail(B200VS_EILLEGAL_PARAMETERS, "shard_train: nlist must be a multiple of the world size");
    const int nl = nlist / W;
    b200vs_params p = ix->params;
    p.nlist = nl;
    std::unique_ptr<IndexBase> loc(make_ivf_flat(ix->metric, d, p));
    loc->train(n, x);
    std::vector<unsigned char> blob((size_t)loc->get_state(nullptr, 0));
    loc->get_state(blob.data(), blob.size());
    const int64_t* hdr = reinterpret_cast<const int64_t*>(blob.data());
    if ((int)hdr[1] != nl) fail(B200VS_EILLEGAL_PARAMETERS, "shard_train: too few training rows on this rank for nlist / world centroids");
    ix->set_device();
    DevBuf<float> all;
    all.reserve((size_t)nlist * d, 0, sh->ws);
    B200VS_CUDA(cudaMemcpyAsync(all.p + (size_t)sh->rank * nl * d, blob.data() + 32, (size_t)nl * d * 4, cudaMemcpyHostToDevice, sh->ws));
    if (W > 1) B200VS_NCCL(nccl().AllGather(all.p + (size_t)sh->rank * nl * d, all.p, (size_t)nl * d, ncclFloat, sh->lanes[0]->comm, sh->ws));
    std::vector<unsigned char> st(32 + (size_t)nlist * d * 4);
    int64_t h4[4] = {0x43465649, nlist, d, (int64_t)ix->metric};
    memcpy(st.data(), h4, 32);
    B200VS_CUDA(cudaMemcpyAsync(st.data() + 32, all.p, (size_t)nlist * d * 4, cudaMemcpyDeviceToHost, sh->ws));
    B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    ix->set_state(st.data(), st.size());
    return B200VS_OK;
  });
}

// Replicate rank `root`'s trained state (centroids) on every rank.
int b200vs_shard_broadcast_state(b200vs_shard* h, int32_t root) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    IndexBase* ix = sh->ix;
    if (root < 0 || root >= sh->world) fail(B200VS_EILLEGAL_PARAMETERS, "bad root");
    std::lock_guard<std::mutex> g(sh->add_mu);
    ix->set_device();
    DevBuf<long long> len;
    len.reserve(1, 0, sh->ws);
    long long need = sh->rank == root ? (long long)ix->get_state(nullptr, 0) : 0;
    if (sh->rank == root && need <= 0) fail(B200VS_EVECTOR_NOT_TRAIN, "root has no trained state");
    B200VS_CUDA(cudaMemcpyAsync(len.p, &need, 8, cudaMemcpyHostToDevice, sh->ws));
    if (sh->world > 1) B200VS_NCCL(nccl().Broadcast(len.p, len.p, 1, ncclInt64, root, sh->lanes[0]->comm, sh->ws));
    B200VS_CUDA(cudaMemcpyAsync(&need, len.p, 8, cudaMemcpyDeviceToHost, sh->ws));
    B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    std::vector<unsigned char> blob((size_t)need);
    if (sh->rank == root) ix->get_state(blob.data(), blob.size());
    DevBuf<unsigned char> db;
    db.reserve((size_t)need, 0, sh->ws);
    if (sh->rank == root) B200VS_CUDA(cudaMemcpyAsync(db.p, blob.data(), (size_t)need, cudaMemcpyHostToDevice, sh->ws));
    if (sh->world > 1) B200VS_NCCL(nccl().Broadcast(db.p, db.p, (size_t)need, ncclUint8, root, sh->lanes[0]->comm, sh->ws));
    if (sh->rank != root) {
      B200VS_CUDA(cudaMemcpyAsync(blob.data(), db.p, (size_t)need, cudaMemcpyDeviceToHost, sh->ws));
      B200VS_CUDA(cudaStreamSynchronize(sh->ws));
      ix->set_state(blob.data(), blob.size());
    } else {
      B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    }
    return B200VS_OK;
  });
}

}  // extern "C"

namespace {

// assignment of n device rows (raw) -> prepared rows (cosine: normalised copy) + lists, on the write stream.
// Returns the prepared rows (either x itself or `prep`).
const float* assign_rows(b200vs_shard* sh, int64_t n, const float* x_dev, DevBuf<float>& prep, DevBuf<long long>& lists) {
  IndexBase* ix = sh->ix;
  cudaStream_t s = sh->ws;
  const float* xp = x_dev;
  if (ix->metric == B200VS_COSINE) {
    prep.reserve((size_t)n * ix->dim, 0, s);
    B200VS_CUDA(cudaMemcpyAsync(prep.p, x_dev, (size_t)n * ix->dim * 4, cudaMemcpyDeviceToDevice, s));
    launch_normalize_faiss(prep.p, n, ix->dim, s);
    xp = prep.p;
  }
  lists.reserve((size_t)n, 0, s);
  {
    std::shared_lock<std::shared_mutex> rl(ix->rw);
    LaneGuard lane(ix, s);
    ix->assign_lists_dev(n, xp, lists.p, s);
  }
  return xp;
}

void shard_add_device_impl(b200vs_shard* sh, int64_t n, const float* x_dev, const long long* ids_dev) {
  IndexBase* ix = sh->ix;
  const int W = sh->world, d = ix->dim;
  cudaStream_t s = sh->ws;
  ix->set_device();
  DevBuf<float> prep, sx, rx;
  DevBuf<long long> lists, perm, sids, slists, rids, rlists, cnt_all;
  const float* xp = n > 0 ? assign_rows(sh, n, x_dev, prep, lists) : x_dev;
  // owner of every row, stable order by owner (host counting sort over 8 bytes per row)
  std::vector<long long> h_lists((size_t)n), h_perm((size_t)n);
  if (n) B200VS_CUDA(cudaMemcpyAsync(h_lists.data(), lists.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  B200VS_CUDA(cudaStreamSynchronize(s));
  const int per = sh->lists_per_rank();
  std::vector<long long> send_cnt(W, 0), send_off(W + 1, 0);
  for (int64_t i = 0; i < n; ++i) send_cnt[std::min<long long>(W - 1, h_lists[i] / per)]++;
  for (int w = 0; w < W; ++w) send_off[w + 1] = send_off[w] + send_cnt[w];
  {
    std::vector<long long> cur(send_off.begin(), send_off.end() - 1);
    for (int64_t i = 0; i < n; ++i) h_perm[cur[std::min<long long>(W - 1, h_lists[i] / per)]++] = i;
  }
  // counts matrix: every rank learns how many rows each peer sends it
  cnt_all.reserve((size_t)W * W, 0, s);
  B200VS_CUDA(cudaMemcpyAsync(cnt_all.p + (size_t)sh->rank * W, send_cnt.data(), (size_t)W * 8, cudaMemcpyHostToDevice, s));
  if (W > 1) B200VS_NCCL(nccl().AllGather(cnt_all.p + (size_t)sh->rank * W, cnt_all.p, (size_t)W, ncclInt64, sh->lanes[0]->comm, s));
  std::vector<long long> h_cnt((size_t)W * W);
  B200VS_CUDA(cudaMemcpyAsync(h_cnt.data(), cnt_all.p, (size_t)W * W * 8, cudaMemcpyDeviceToHost, s));
  B200VS_CUDA(cudaStreamSynchronize(s));
  std::vector<long long> recv_cnt(W), recv_off(W + 1, 0);
  for (int w = 0; w < W; ++w) { recv_cnt[w] = h_cnt[(size_t)w * W + sh->rank]; recv_off[w + 1] = recv_off[w] + recv_cnt[w]; }
  const long long nrecv = recv_off[W];
  // rows sorted by owner
  if (n) {
    perm.reserve((size_t)n, 0, s); sx.reserve((size_t)n * d, 0, s); sids.reserve((size_t)n, 0, s); slists.reserve((size_t)n, 0, s);
    B200VS_CUDA(cudaMemcpyAsync(perm.p, h_perm.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    gather_rows_kernel<<<(unsigned)cdiv(n * 32, 256), 256, 0, s>>>(xp, ids_dev, lists.p, perm.p, n, d, sx.p, sids.p, slists.p);
    B200VS_CUDA(cudaGetLastError());
  }
  rx.reserve((size_t)std::max<long long>(nrecv, 1) * d, 0, s); rids.reserve((size_t)std::max<long long>(nrecv, 1), 0, s);
  rlists.reserve((size_t)std::max<long long>(nrecv, 1), 0, s);
  if (W > 1) {
    B200VS_NCCL(nccl().GroupStart());
    for (int w = 0; w < W; ++w) {
      if (send_cnt[w]) {
        B200VS_NCCL(nccl().Send(sx.p + (size_t)send_off[w] * d, (size_t)send_cnt[w] * d, ncclFloat, w, sh->lanes[0]->comm, s));
        B200VS_NCCL(nccl().Send(sids.p + send_off[w], (size_t)send_cnt[w], ncclInt64, w, sh->lanes[0]->comm, s));
        B200VS_NCCL(nccl().Send(slists.p + send_off[w], (size_t)send_cnt[w], ncclInt64, w, sh->lanes[0]->comm, s));
      }
      if (recv_cnt[w]) {
        B200VS_NCCL(nccl().Recv(rx.p + (size_t)recv_off[w] * d, (size_t)recv_cnt[w] * d, ncclFloat, w, sh->lanes[0]->comm, s));
        B200VS_NCCL(nccl().Recv(rids.p + recv_off[w], (size_t)recv_cnt[w], ncclInt64, w, sh->lanes[0]->comm, s));
        B200VS_NCCL(nccl().Recv(rlists.p + recv_off[w], (size_t)recv_cnt[w], ncclInt64, w, sh->lanes[0]->comm, s));
      }
    }
    B200VS_NCCL(nccl().GroupEnd());
  } else if (n) {
    B200VS_CUDA(cudaMemcpyAsync(rx.p, sx.p, (size_t)n * d * 4, cudaMemcpyDeviceToDevice, s));
    B200VS_CUDA(cudaMemcpyAsync(rids.p, sids.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    B200VS_CUDA(cudaMemcpyAsync(rlists.p, slists.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
  }
  B200VS_CUDA(cudaStreamSynchronize(s));
  if (nrecv) ix->add_dev(nrecv, rx.p, rids.p, rlists.p, false, true);
}

}  // namespace

extern "C" {

// Collective: every rank passes the rows IT holds (any rows; n may differ per rank, 0 allowed); each row is assigned to its
// nearest centroid and travels to the rank that owns that list (grouped ncclSend / ncclRecv over NVLink).
int b200vs_shard_add_device(b200vs_shard* h, int64_t n, const float* x_dev, const int64_t* ids_dev) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    if (n < 0 || (n > 0 && (!x_dev || !ids_dev))) fail(B200VS_EILLEGAL_PARAMETERS, "bad add arguments");
    std::lock_guard<std::mutex> g(sh->add_mu);
    shard_add_device_impl(sh, n, x_dev, (const long long*)ids_dev);
    return B200VS_OK;
  });
}

int b200vs_shard_add(b200vs_shard* h, int64_t n, const float* x, const int64_t* ids) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    if (n < 0 || (n > 0 && (!x || !ids))) fail(B200VS_EILLEGAL_PARAMETERS, "bad add arguments");
    std::lock_guard<std::mutex> g(sh->add_mu);
    sh->ix->set_device();
    DevBuf<float> dx;
    DevBuf<long long> di;
    if (n) {
      dx.reserve((size_t)n * sh->ix->dim, 0, sh->ws); di.reserve((size_t)n, 0, sh->ws);
      B200VS_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)n * sh->ix->dim * 4, cudaMemcpyHostToDevice, sh->ws));
      B200VS_CUDA(cudaMemcpyAsync(di.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice, sh->ws));
      B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    }
    shard_add_device_impl(sh, n, dx.p, di.p);
    return B200VS_OK;
  });
}

// Collective delete (VectorIndexIvfFlat::Delete, vector_index_ivf_flat.cc:162-189, over all shards): every rank passes the same
// id list and drops the rows it holds; the removed counts are summed over the ranks.  "remove not found vector id" ->
// EVECTOR_INVALID only when NO rank held any of the ids (:180-184).  Upsert of a sharded index = this call (a not-found
// status ignored), then b200vs_shard_add*: the new row may live on another rank than the old one.
int b200vs_shard_remove_ids(b200vs_shard* h, int64_t n, const int64_t* ids, int64_t* n_removed) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    if (n_removed) *n_removed = 0;
    if (n <= 0) return B200VS_OK;  // "delete_ids.empty() -> OK"
    if (!ids) fail(B200VS_EILLEGAL_PARAMETERS, "null ids");
    std::lock_guard<std::mutex> g(sh->add_mu);
    IndexBase* ix = sh->ix;
    const int64_t local = ix->remove(n, ids);  // -1: untrained (OK, nothing to do)
    long long tot[2] = {local > 0 ? (long long)local : 0, local < 0 ? 1 : 0};
    if (sh->world > 1) {
      ix->set_device();
      DevBuf<long long> d;
      d.reserve(2, 0, sh->ws);
      B200VS_CUDA(cudaMemcpyAsync(d.p, tot, 16, cudaMemcpyHostToDevice, sh->ws));
      B200VS_NCCL(nccl().AllReduce(d.p, d.p, 2, ncclInt64, ncclSum, sh->lanes[0]->comm, sh->ws));
      B200VS_CUDA(cudaMemcpyAsync(tot, d.p, 16, cudaMemcpyDeviceToHost, sh->ws));
      B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    }
    if (n_removed) *n_removed = tot[0];
    if (tot[0] == 0 && tot[1] == 0) fail(B200VS_EVECTOR_INVALID, "remove not found vector id");
    return B200VS_OK;
  });
}

// Bulk builds of large shards: first pass every chunk through plan_add (assignment only, rows are NOT stored), then
// plan_commit all-reduces the per-list counts and pre-sizes the owned lists in one arena allocation; the second pass
// (b200vs_shard_add*) then never relocates a list or re-allocates the arena.
int b200vs_shard_plan_add_device(b200vs_shard* h, int64_t n, const float* x_dev) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    if (n <= 0 || !x_dev) fail(B200VS_EILLEGAL_PARAMETERS, "bad plan arguments");
    std::lock_guard<std::mutex> g(sh->add_mu);
    IndexBase* ix = sh->ix;
    ix->set_device();
    const int nlist = ix->nlist_now();
    if (!sh->planning) {
      sh->plan_counts.reserve((size_t)nlist, 0, sh->ws);
      B200VS_CUDA(cudaMemsetAsync(sh->plan_counts.p, 0, (size_t)nlist * 8, sh->ws));
      sh->planning = true;
    }
    DevBuf<float> prep;
    DevBuf<long long> lists;
    assign_rows(sh, n, x_dev, prep, lists);
    count_lists_kernel<<<(unsigned)cdiv(n, 256), 256, 0, sh->ws>>>(lists.p, n, nlist, sh->plan_counts.p);
    B200VS_CUDA(cudaGetLastError());
    B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    return B200VS_OK;
  });
}

int b200vs_shard_plan_commit(b200vs_shard* h) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    std::lock_guard<std::mutex> g(sh->add_mu);
    IndexBase* ix = sh->ix;
    ix->set_device();
    const int nlist = ix->nlist_now();
    if (!sh->planning) {
      sh->plan_counts.reserve((size_t)nlist, 0, sh->ws);
      B200VS_CUDA(cudaMemsetAsync(sh->plan_counts.p, 0, (size_t)nlist * 8, sh->ws));
    }
    if (sh->world > 1) B200VS_NCCL(nccl().AllReduce(sh->plan_counts.p, sh->plan_counts.p, (size_t)nlist, ncclUint64, ncclSum, sh->lanes[0]->comm, sh->ws));
    std::vector<int64_t> cnt((size_t)nlist);
    B200VS_CUDA(cudaMemcpyAsync(cnt.data(), sh->plan_counts.p, (size_t)nlist * 8, cudaMemcpyDeviceToHost, sh->ws));
    B200VS_CUDA(cudaStreamSynchronize(sh->ws));
    sh->planning = false;
    const int per = sh->lists_per_rank();
    for (int l = 0; l < nlist; ++l)
      if (std::min(sh->world - 1, l / per) != sh->rank) cnt[l] = 0;  // lists owned elsewhere stay empty here
    ix->reserve_lists(cnt.data(), nlist);
    return B200VS_OK;
  });
}

int b200vs_shard_search_device(b200vs_shard* h, int64_t seq, int64_t nq, const float* xq_dev, int32_t k, const b200vs_search_params* sp,
                               float* out_dist_dev, int64_t* out_ids_dev, void* stream) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    IndexBase* ix = sh->ix;
    if (nq <= 0 || !xq_dev) fail(B200VS_EILLEGAL_PARAMETERS, "vector_with_ids is empty");
    if (k <= 0) return B200VS_OK;
    if (!out_dist_dev || !out_ids_dev) fail(B200VS_EILLEGAL_PARAMETERS, "null output");
    SeqLane sl(sh, seq);
    std::shared_lock<std::shared_mutex> rl(ix->rw);
    ix->set_device();
    cudaStream_t s = stream ? (cudaStream_t)stream : sl.lane->own;
    LaneGuard lane(ix, s);
    ix->reset_stats();
    shard_search_stream(sh, *sl.lane, nq, xq_dev, k, sp, out_dist_dev, (long long*)out_ids_dev, s);
    ix->phases_finish(s);
    if (!stream) B200VS_CUDA(cudaStreamSynchronize(s));
    return B200VS_OK;
  });
}

// Host-pointer variant (the call a dingo-store node makes): every rank passes the SAME batch; each rank uploads only its
// slice and the slices are all-gathered over NVLink, so the PCIe traffic per rank does not grow with the world size.
int b200vs_shard_search(b200vs_shard* h, int64_t seq, int64_t nq, const float* xq, int32_t k, const b200vs_search_params* sp, float* out_dist,
                        int64_t* out_ids) {
  return guarded([&]() -> int {
    b200vs_shard* sh = get(h);
    IndexBase* ix = sh->ix;
    if (nq <= 0 || !xq) fail(B200VS_EILLEGAL_PARAMETERS, "vector_with_ids is empty");
    if (k <= 0) return B200VS_OK;
    if (!out_dist || !out_ids) fail(B200VS_EILLEGAL_PARAMETERS, "null output");
    const int W = sh->world, r = sh->rank, d = ix->dim;
    cudaStream_t s;
    float* dd;
    long long* di;
    {
      SeqLane sl(sh, seq);
      std::shared_lock<std::shared_mutex> rl(ix->rw);
      ix->set_device();
      b200vs_shard::SLane& L = *sl.lane;
      s = L.own;
      LaneGuard lane(ix, s);
      ix->reset_stats();
      const int64_t per = (nq + W - 1) / W;
      L.q_all.reserve((size_t)per * W * d, 0, s);
      const int64_t q0 = std::min<int64_t>(nq, per * r), q1 = std::min<int64_t>(nq, per * (r + 1));
      float* mine = L.q_all.p + (size_t)per * r * d;
      if (q1 > q0) B200VS_CUDA(cudaMemcpyAsync(mine, xq + (size_t)q0 * d, (size_t)(q1 - q0) * d * 4, cudaMemcpyHostToDevice, s));
      if (q1 - q0 < per) B200VS_CUDA(cudaMemsetAsync(mine + (size_t)(q1 - q0) * d, 0, (size_t)(per - (q1 - q0)) * d * 4, s));
      if (W > 1) B200VS_NCCL(nccl().AllGather(mine, L.q_all.p, (size_t)per * d, ncclFloat, L.comm, s));
      dd = ix->scratch.alloc<float>((size_t)nq * k);
      di = ix->scratch.alloc<long long>((size_t)nq * k);
      shard_search_stream(sh, L, nq, L.q_all.p, k, sp, dd, di, s);
      ix->phases_finish(s);
      B200VS_CUDA(cudaMemcpyAsync(out_dist, dd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
      B200VS_CUDA(cudaMemcpyAsync(out_ids, di, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    }  // the lane is released before the host waits: the next batch of this lane may be enqueued behind this one
    B200VS_CUDA(cudaStreamSynchronize(s));
    return B200VS_OK;
  });
}

}  // extern "C"
