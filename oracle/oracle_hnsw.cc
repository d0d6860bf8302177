// oracle_hnsw.cc — CPU ORACLE (test infrastructure): HNSW build + search.
//
// Reference call sites (src/vector/vector_index_hnsw.cc): spaces :153-160 (L2Space; IP and cosine use
// InnerProductSpace, cosine normalises with NormalizeVectorForHnsw); constructor
// HierarchicalNSW(space, max_elements, M=nlinks, efConstruction, random_seed=100,
// allow_replace_deleted=false) :181-182; addPoint :231,:242; setEf :427 (sticky);
// searchKnn(q, k, filter) :436,:459; results popped back-to-front so the output is ascending :400-419;
// the distance is emitted as hnswlib returns it (L2: squared L2; IP space: 1 - ip) :374.
//
// The graph algorithm lives in dingodb/hnswlib@1964db3e (NOT vendored).  It is restated here from the
// published hnswlib 0.7 algorithm ("parity unpinned"): level = (int)(-ln(U)/ln(M)) with
// std::default_random_engine(seed) and uniform_real_distribution<double>(0,1); maxM0 = 2M;
// ef_construction = max(efc, M); heuristic neighbour selection; searchBaseLayerST with
// ef = max(ef_, k); filtered / deleted nodes are traversed but never returned.
// The heaps are std::priority_queue with a compare-by-distance-only comparator, as in hnswlib.
#include <queue>
#include <random>
#include <unordered_map>

#include "oracle_common.h"

using namespace oracle;

namespace {
typedef uint32_t tableint;
typedef std::pair<float, tableint> Pair;
struct CompareByFirst {
  constexpr bool operator()(Pair const& a, Pair const& b) const noexcept { return a.first < b.first; }
};
typedef std::priority_queue<Pair, std::vector<Pair>, CompareByFirst> Heap;
}  // namespace

struct oracle_hnsw {
  int metric;
  int32_t d;
  int64_t max_elements;
  size_t M, maxM, maxM0, ef_construction;
  size_t ef = 10;
  double mult, revSize;
  std::default_random_engine level_generator;
  int64_t cur = 0;
  int maxlevel = -1;
  tableint enterpoint = (tableint)-1;
  std::vector<float> data;                     // [n, d], hnsw-normalised for cosine
  std::vector<int64_t> labels;                 // [n]
  std::vector<int> levels;                     // [n]
  std::vector<tableint> link0;                 // [n, maxM0+1]  (count, neighbours...)
  std::vector<std::vector<tableint>> linkup;   // per element: levels * (maxM+1)
  std::unordered_map<int64_t, tableint> lookup;
  std::vector<uint32_t> visited;               // construction-time visited tags
  uint32_t tag = 0;

  inline const float* vec(tableint i) const { return &data[(size_t)i * d]; }
  inline float dist(const float* a, const float* b) const {
    if (metric == ORACLE_L2) return oracle_fvec_L2sqr(a, b, d);
    return 1.0f - oracle_fvec_inner_product(a, b, d);  // hnswlib InnerProductDistance
  }
  inline tableint* ll(tableint i, int level) {
    return level == 0 ? &link0[(size_t)i * (maxM0 + 1)] : &linkup[i][(size_t)(level - 1) * (maxM + 1)];
  }
  inline const tableint* ll(tableint i, int level) const {
    return level == 0 ? &link0[(size_t)i * (maxM0 + 1)] : &linkup[i][(size_t)(level - 1) * (maxM + 1)];
  }
  int random_level() {
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    double r = -log(distribution(level_generator)) * mult;
    return (int)r;
  }

  Heap search_base_layer(tableint ep, const float* q, int layer) {
    if (++tag == 0) { std::fill(visited.begin(), visited.end(), 0u); tag = 1; }
    Heap top, cand;
    float lower;
    float d0 = dist(q, vec(ep));
    top.emplace(d0, ep);
    lower = d0;
    cand.emplace(-d0, ep);
    visited[ep] = tag;
    while (!cand.empty()) {
      Pair curr = cand.top();
      if ((-curr.first) > lower && top.size() == ef_construction) break;
      cand.pop();
      const tableint* l = ll(curr.second, layer);
      size_t size = l[0];
      for (size_t j = 1; j <= size; ++j) {
        tableint c = l[j];
        if (visited[c] == tag) continue;
        visited[c] = tag;
        float d1 = dist(q, vec(c));
        if (top.size() < ef_construction || lower > d1) {
          cand.emplace(-d1, c);
          top.emplace(d1, c);
          if (top.size() > ef_construction) top.pop();
          if (!top.empty()) lower = top.top().first;
        }
      }
    }
    return top;
  }

  void heuristic(Heap& top, size_t Mlim) {
    if (top.size() < Mlim) return;
    Heap closest;
    std::vector<Pair> ret;
    while (!top.empty()) { closest.emplace(-top.top().first, top.top().second); top.pop(); }
    while (!closest.empty()) {
      if (ret.size() >= Mlim) break;
      Pair cur = closest.top();
      float dq = -cur.first;
      closest.pop();
      bool good = true;
      for (const Pair& s : ret) {
        float cd = dist(vec(s.second), vec(cur.second));
        if (cd < dq) { good = false; break; }
      }
      if (good) ret.push_back(cur);
    }
    for (const Pair& p : ret) top.emplace(-p.first, p.second);
  }

  tableint connect(tableint cur_c, Heap& top, int level) {
    size_t Mcurmax = level ? maxM : maxM0;
    heuristic(top, M);
    std::vector<tableint> sel;
    sel.reserve(M);
    while (!top.empty()) { sel.push_back(top.top().second); top.pop(); }
    tableint next_ep = sel.back();
    {
      tableint* l = ll(cur_c, level);
      l[0] = (tableint)sel.size();
      for (size_t i = 0; i < sel.size(); ++i) l[1 + i] = sel[i];
    }
    for (size_t idx = 0; idx < sel.size(); ++idx) {
      tableint* lo = ll(sel[idx], level);
      size_t sz = lo[0];
      if (sz < Mcurmax) {
        lo[1 + sz] = cur_c;
        lo[0] = (tableint)(sz + 1);
      } else {
        float d_max = dist(vec(cur_c), vec(sel[idx]));
        Heap cands;
        cands.emplace(d_max, cur_c);
        for (size_t j = 0; j < sz; ++j) cands.emplace(dist(vec(lo[1 + j]), vec(sel[idx])), lo[1 + j]);
        heuristic(cands, Mcurmax);
        int indx = 0;
        while (!cands.empty()) { lo[1 + indx] = cands.top().second; cands.pop(); indx++; }
        lo[0] = (tableint)indx;
      }
    }
    return next_ep;
  }

  int add_point(const float* x, int64_t label) {
    if (lookup.count(label)) return -2;  // update-in-place not restated
    if (cur >= max_elements) return -3;
    tableint cur_c = (tableint)cur++;
    lookup[label] = cur_c;
    int curlevel = random_level();
    int maxlevelcopy = maxlevel;
    tableint currObj = enterpoint;
    memcpy(&data[(size_t)cur_c * d], x, sizeof(float) * d);
    labels[cur_c] = label;
    levels[cur_c] = curlevel;
    if (curlevel) linkup[cur_c].assign((size_t)curlevel * (maxM + 1), 0);
    const float* q = vec(cur_c);
    if ((int)currObj != -1) {
      if (curlevel < maxlevelcopy) {
        float curdist = dist(q, vec(currObj));
        for (int level = maxlevelcopy; level > curlevel; level--) {
          bool changed = true;
          while (changed) {
            changed = false;
            const tableint* l = ll(currObj, level);
            int size = l[0];
            for (int i = 1; i <= size; ++i) {
              tableint c = l[i];
              float dd = dist(q, vec(c));
              if (dd < curdist) { curdist = dd; currObj = c; changed = true; }
            }
          }
        }
      }
      for (int level = std::min(curlevel, maxlevelcopy); level >= 0; level--) {
        Heap top = search_base_layer(currObj, q, level);
        currObj = connect(cur_c, top, level);
      }
    } else {
      enterpoint = 0;
      maxlevel = curlevel;
    }
    if (curlevel > maxlevelcopy) { enterpoint = cur_c; maxlevel = curlevel; }
    return 0;
  }

  // searchKnn + searchBaseLayerST<has_deletions=false, collect_metrics=true>
  void search(const float* q, int k, size_t ef_run, const oracle_filter* filt, float* od, int64_t* oi,
              int64_t* ndis, int64_t* hops) const {
    for (int i = 0; i < k; ++i) { od[i] = 0; oi[i] = -1; }
    int64_t nd = 0, nh = 0;
    if (cur == 0) { if (ndis) *ndis = 0; if (hops) *hops = 0; return; }
    tableint currObj = enterpoint;
    float curdist = dist(q, vec(enterpoint));
    nd++;
    for (int level = maxlevel; level > 0; level--) {
      bool changed = true;
      while (changed) {
        changed = false;
        const tableint* l = ll(currObj, level);
        int size = l[0];
        nh++; nd += size;
        for (int i = 1; i <= size; ++i) {
          tableint c = l[i];
          float dd = dist(q, vec(c));
          if (dd < curdist) { curdist = dd; currObj = c; changed = true; }
        }
      }
    }
    const size_t ef_ = std::max(ef_run, (size_t)k);
    std::vector<uint8_t> vis((size_t)cur, 0);
    Heap top, cand;
    float lower;
    if (!filt || filter_pass(filt, labels[currObj])) {
      float d0 = dist(q, vec(currObj));
      lower = d0;
      top.emplace(d0, currObj);
      cand.emplace(-d0, currObj);
    } else {
      lower = std::numeric_limits<float>::max();
      cand.emplace(-lower, currObj);
    }
    vis[currObj] = 1;
    while (!cand.empty()) {
      Pair cp = cand.top();
      if ((-cp.first) > lower && (top.size() == ef_ || !filt)) break;
      cand.pop();
      const tableint* l = ll(cp.second, 0);
      size_t size = l[0];
      nh++; nd += (int64_t)size;
      for (size_t j = 1; j <= size; ++j) {
        tableint c = l[j];
        if (vis[c]) continue;
        vis[c] = 1;
        float dd = dist(q, vec(c));
        if (top.size() < ef_ || lower > dd) {
          cand.emplace(-dd, c);
          if (!filt || filter_pass(filt, labels[c])) top.emplace(dd, c);
          if (top.size() > ef_) top.pop();
          if (!top.empty()) lower = top.top().first;
        }
      }
    }
    while (top.size() > (size_t)k) top.pop();
    // caller-side: std::priority_queue<pair<float,label>> popped back-to-front (hnsw.cc:400-419)
    std::priority_queue<std::pair<float, int64_t>> res;
    while (!top.empty()) { res.push({top.top().first, labels[top.top().second]}); top.pop(); }
    int n = (int)res.size();
    for (int i = n - 1; i >= 0; --i) { od[i] = res.top().first; oi[i] = res.top().second; res.pop(); }
    if (ndis) *ndis = nd;
    if (hops) *hops = nh;
  }
};

extern "C" {

oracle_hnsw* oracle_hnsw_create(int metric, int32_t d, int64_t max_elements, int32_t M, int32_t efc, int64_t seed) {
  oracle_hnsw* h = new oracle_hnsw();
  h->metric = metric; h->d = d; h->max_elements = max_elements;
  h->M = M; h->maxM = M; h->maxM0 = 2 * (size_t)M;
  h->ef_construction = std::max((size_t)efc, (size_t)M);
  h->mult = 1 / log(1.0 * M); h->revSize = 1.0 / h->mult;
  h->level_generator.seed((unsigned)seed);
  h->data.resize((size_t)max_elements * d);
  h->labels.resize(max_elements);
  h->levels.resize(max_elements);
  h->link0.assign((size_t)max_elements * (h->maxM0 + 1), 0);
  h->linkup.resize(max_elements);
  h->visited.assign(max_elements, 0);
  return h;
}
void oracle_hnsw_destroy(oracle_hnsw* h) { delete h; }

int oracle_hnsw_add(oracle_hnsw* h, int64_t n, const float* x, const int64_t* labels) {
  std::vector<float> tmp(h->d);
  for (int64_t i = 0; i < n; ++i) {
    const float* xi = x + i * (int64_t)h->d;
    if (h->metric == ORACLE_COSINE) { oracle_normalize_hnsw(xi, h->d, tmp.data()); xi = tmp.data(); }
    int rc = h->add_point(xi, labels[i]);
    if (rc) return rc;
  }
  return 0;
}

int oracle_hnsw_search(oracle_hnsw* h, int64_t nq, const float* xq, int32_t k, int32_t ef, const oracle_filter* filt,
                       int nthreads, float* out_dist, int64_t* out_ids, int64_t* out_ndis, int64_t* out_hops) {
  if (ef > 0) h->ef = ef;  // sticky setEf, vector_index_hnsw.cc:426-428
  const size_t ef_run = h->ef;
  parallel_for(nq, nthreads, [&](int64_t qi) {
    std::vector<float> qb(xq + qi * (int64_t)h->d, xq + (qi + 1) * (int64_t)h->d);
    if (h->metric == ORACLE_COSINE) {
      std::vector<float> t(h->d);
      oracle_normalize_hnsw(qb.data(), h->d, t.data());
      qb.swap(t);
    }
    h->search(qb.data(), k, ef_run, filt, out_dist + qi * (int64_t)k, out_ids + qi * (int64_t)k,
              out_ndis ? out_ndis + qi : nullptr, out_hops ? out_hops + qi : nullptr);
  });
  return 0;
}

// Flat export consumed by b200vs_set_trained_state(B200VS_HNSW) — layout documented in include/b200vs.h.
//   int64 hdr[8] = {magic 'HNSW', n, d, M, maxM0, maxlevel, enterpoint, metric}
//   int32 levels[n]; int64 up_off[n+1] (in uint32 units); uint32 link0[n*(maxM0+1)];
//   uint32 linkup[up_off[n]]; float data[n*d]; int64 labels[n]
int64_t oracle_hnsw_export_size(oracle_hnsw* h) {
  int64_t n = h->cur, up = 0;
  for (int64_t i = 0; i < n; ++i) up += (int64_t)h->linkup[i].size();
  int64_t sz = 8 * 8 + n * 4;
  sz = (sz + 7) / 8 * 8;
  sz += (n + 1) * 8 + n * (int64_t)(h->maxM0 + 1) * 4 + up * 4;
  sz = (sz + 7) / 8 * 8;
  sz += n * (int64_t)h->d * 4;
  sz = (sz + 7) / 8 * 8;
  sz += n * 8;
  return sz;
}
int oracle_hnsw_export(oracle_hnsw* h, void* blob, int64_t len) {
  if (len < oracle_hnsw_export_size(h)) return -1;
  int64_t n = h->cur;
  char* p = (char*)blob;
  char* base = p;
  int64_t hdr[8] = {0x57534E48, n, h->d, (int64_t)h->M, (int64_t)h->maxM0, h->maxlevel, (int64_t)h->enterpoint, h->metric};
  memcpy(p, hdr, sizeof(hdr)); p += sizeof(hdr);
  memcpy(p, h->levels.data(), n * 4); p += n * 4;
  p = base + ((p - base) + 7) / 8 * 8;
  int64_t* off = (int64_t*)p; p += (n + 1) * 8;
  int64_t acc = 0;
  for (int64_t i = 0; i < n; ++i) { off[i] = acc; acc += (int64_t)h->linkup[i].size(); }
  off[n] = acc;
  memcpy(p, h->link0.data(), n * (int64_t)(h->maxM0 + 1) * 4); p += n * (int64_t)(h->maxM0 + 1) * 4;
  for (int64_t i = 0; i < n; ++i) { memcpy(p, h->linkup[i].data(), h->linkup[i].size() * 4); p += h->linkup[i].size() * 4; }
  p = base + ((p - base) + 7) / 8 * 8;
  memcpy(p, h->data.data(), n * (int64_t)h->d * 4); p += n * (int64_t)h->d * 4;
  p = base + ((p - base) + 7) / 8 * 8;
  memcpy(p, h->labels.data(), n * 8);
  return 0;
}

// Load a graph in the export layout (a graph built elsewhere, e.g. by the engine's concurrent builder whose insertion
// order — like the reference's 16-thread build, vector_index_hnsw.cc:229-243 — is not reproducible): the oracle then
// SEARCHES exactly that graph.  Returns 0, or -1 when the blob does not fit this oracle's (d, M, max_elements).
int oracle_hnsw_import(oracle_hnsw* h, const void* blob, int64_t len) {
  if (len < 64) return -1;
  const char* base = (const char*)blob;
  const int64_t* hdr = (const int64_t*)base;
  const int64_t n = hdr[1];
  if (hdr[0] != 0x57534E48 || hdr[2] != h->d || hdr[3] != (int64_t)h->M || n > h->max_elements) return -1;
  const char* p = base + 64;
  memcpy(h->levels.data(), p, n * 4); p += n * 4;
  p = base + ((p - base) + 7) / 8 * 8;
  const int64_t* off = (const int64_t*)p; p += (n + 1) * 8;
  memcpy(h->link0.data(), p, (size_t)n * (h->maxM0 + 1) * 4); p += (size_t)n * (h->maxM0 + 1) * 4;
  for (int64_t i = 0; i < n; ++i) h->linkup[i].assign((const tableint*)p + off[i], (const tableint*)p + off[i + 1]);
  p += off[n] * 4;
  p = base + ((p - base) + 7) / 8 * 8;
  memcpy(h->data.data(), p, (size_t)n * h->d * 4); p += (size_t)n * h->d * 4;
  p = base + ((p - base) + 7) / 8 * 8;
  if ((p - base) + n * 8 > len) return -1;
  memcpy(h->labels.data(), p, n * 8);
  h->cur = n; h->maxlevel = (int)hdr[5]; h->enterpoint = (tableint)hdr[6];
  h->lookup.clear();
  for (int64_t i = 0; i < n; ++i) h->lookup[h->labels[i]] = (tableint)i;
  return 0;
}

}  // extern "C"
