// test_plugin.cc — exercises the drop-in subclass the way the reference's gtest suites exercise the stock plugins
// (test/unit_test/vector/test_vector_index_{flat,ivf_flat,hnsw,flat_search_param}.cc): same fixtures (default-seeded
// std::mt19937, row[0] += i/1000.), same asserted contract (status codes, result counts, filter containment,
// self-match at rank 0).  Needs a GPU; run by tests/test_gpu_plugin_cpp.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>

#include "vector_index_b200.h"

using namespace dingodb;

static int g_fail = 0;
#define EXPECT(cond)                                                                \
  do {                                                                              \
    if (!(cond)) { ++g_fail; printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

static std::vector<float> fixture(int n, int d) {  // test_vector_index_flat.cc:491-500
  std::mt19937 rng;
  std::uniform_real_distribution<> distrib;
  std::vector<float> x((size_t)n * d);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < d; j++) x[(size_t)d * i + j] = distrib(rng);
    x[(size_t)d * i] += i / 1000.;
  }
  return x;
}
static std::vector<pb::common::VectorWithId> to_pb(const std::vector<float>& x, int n, int d, int64_t first_id) {
  std::vector<pb::common::VectorWithId> out(n);
  for (int i = 0; i < n; ++i) {
    out[i].set_id(first_id + i);
    out[i].mutable_vector()->set_dimension(d);
    out[i].mutable_vector()->set_value_type(pb::common::ValueType::FLOAT);
    for (int j = 0; j < d; ++j) out[i].mutable_vector()->add_float_values(x[(size_t)i * d + j]);
  }
  return out;
}
static VectorIndexPtr make(pb::common::VectorIndexType t, pb::common::MetricType m, int d, int nlist = 0, int M = 0) {
  pb::common::VectorIndexParameter p;
  p.set_vector_index_type(t);
  if (t == pb::common::VECTOR_INDEX_TYPE_FLAT) { p.mutable_flat_parameter()->dimension_ = d; p.mutable_flat_parameter()->metric_type_ = m; }
  if (t == pb::common::VECTOR_INDEX_TYPE_IVF_FLAT) { auto* q = p.mutable_ivf_flat_parameter(); q->dimension_ = d; q->metric_type_ = m; q->ncentroids_ = nlist; }
  if (t == pb::common::VECTOR_INDEX_TYPE_HNSW) { auto* q = p.mutable_hnsw_parameter(); q->dimension_ = d; q->metric_type_ = m; q->nlinks_ = M; q->efconstruction_ = 200; q->max_elements_ = 1000; }
  auto ix = std::make_shared<VectorIndexB200>(1, p, pb::common::RegionEpoch(), pb::common::Range(), nullptr);
  int64_t c = 0;
  if (!ix->GetCount(c).ok()) {  // no CUDA device / library failure: the product has no CPU fallback, so neither has this test
    printf("cannot create the index: %s\n", b200vs_last_error());
    exit(2);
  }
  return ix;
}

static void test_flat() {
  const int n = 10, d = 8;
  auto x = fixture(n, d);
  auto vs = to_pb(x, n, d, 1);
  for (auto metric : {pb::common::METRIC_TYPE_L2, pb::common::METRIC_TYPE_INNER_PRODUCT, pb::common::METRIC_TYPE_COSINE}) {
    auto ix = make(pb::common::VECTOR_INDEX_TYPE_FLAT, metric, d);
    std::vector<pb::index::VectorWithDistanceResult> results;
    pb::common::VectorSearchParameter sp;
    // empty add / search -> EILLEGAL_PARAMTETERS (test_vector_index_flat.cc:519)
    EXPECT(ix->Add({}).error_code() == pb::error::EILLEGAL_PARAMTETERS);
    EXPECT(ix->Search({}, 3, {}, false, sp, results).error_code() == pb::error::EILLEGAL_PARAMTETERS);
    // dimension mismatch -> EVECTOR_INVALID (:534)
    auto bad = to_pb(x, 1, d - 1, 100);
    EXPECT(ix->Add(bad).error_code() == pb::error::EVECTOR_INVALID);
    EXPECT(ix->Add(vs).ok());
    int64_t count = 0;
    EXPECT(ix->GetCount(count).ok() && count == n);
    // duplicated id inside a batch
    auto dup = to_pb(x, 2, d, 50);
    dup[1].set_id(50);
    EXPECT(ix->Upsert(dup).error_code() == pb::error::EVECTOR_ID_DUPLICATED);
    // topk == 0 -> OK, results untouched (:873-909)
    EXPECT(ix->Search({vs[0]}, 0, {}, false, sp, results).ok() && results.empty());
    // search appends one result per query, ascending distances, self at rank 0
    EXPECT(ix->Search({vs[0], vs[1]}, 3, {}, false, sp, results).ok());
    EXPECT(results.size() == 2);
    for (size_t q = 0; q < results.size(); ++q) {
      EXPECT(results[q].vector_with_distances_size() == 3);
      EXPECT(results[q].vector_with_distances(0).vector_with_id().id() == (int64_t)q + 1 || metric == pb::common::METRIC_TYPE_INNER_PRODUCT);
      for (int i = 1; i < 3; ++i) EXPECT(results[q].vector_with_distances(i - 1).distance() <= results[q].vector_with_distances(i).distance());
      EXPECT(results[q].vector_with_distances(0).metric_type() == metric);
    }
    // filters: range + sorted list; results stay inside the allowed set (test_vector_index_flat_search_param.cc:274-283)
    results.clear();
    std::vector<int64_t> allow = {2, 4, 6, 8};
    auto sortf = std::make_shared<VectorIndex::SortFilterFunctor>(allow);
    auto rangef = std::make_shared<VectorIndex::RangeFilterFunctor>(3, 9);
    EXPECT(ix->Search({vs[0]}, 10, {sortf, rangef}, false, sp, results).ok());
    std::set<int64_t> got;
    for (const auto& r : results[0].vector_with_distances()) got.insert(r.vector_with_id().id());
    EXPECT((got == std::set<int64_t>{4, 6, 8}));
    // delete, then never returned
    EXPECT(ix->Delete({4}).ok());
    results.clear();
    EXPECT(ix->Search({vs[0]}, 10, {}, false, sp, results).ok());
    for (const auto& r : results[0].vector_with_distances()) EXPECT(r.vector_with_id().id() != 4);
    EXPECT(results[0].vector_with_distances_size() == n - 1);  // fewer than k hits -> shorter list
    // range search
    results.clear();
    EXPECT(ix->RangeSearch({vs[0]}, 1e9f, {}, false, sp, results).ok());  // radius is in API distance semantics (1 - ip for IP / cosine)
    EXPECT(results.size() == 1 && results[0].vector_with_distances_size() == n - 1);
    EXPECT(!ix->SupportSave() && !ix->NeedToSave(20000) && !ix->NeedTrain() && ix->IsTrained());
  }
}

static void test_ivf_flat() {
  const int n = 100, d = 8;  // test_vector_index_ivf_flat.cc:108-110: 100 x 8, nlist = 10
  auto x = fixture(n, d);
  auto vs = to_pb(x, n, d, 1);
  auto ix = make(pb::common::VECTOR_INDEX_TYPE_IVF_FLAT, pb::common::METRIC_TYPE_L2, d, 10);
  pb::common::VectorSearchParameter sp;
  std::vector<pb::index::VectorWithDistanceResult> results;
  EXPECT(ix->NeedTrain() && !ix->IsTrained());
  // untrained search -> OK + nq empty results (:286)
  EXPECT(ix->Search({vs[0], vs[1]}, 3, {}, false, sp, results).ok());
  EXPECT(results.size() == 2 && results[0].vector_with_distances_size() == 0);
  // Add on an untrained index trains with the batch and retries (ivf_flat.cc:133-150)
  EXPECT(ix->Add(vs).ok());
  EXPECT(ix->IsTrained());
  int64_t count = 0;
  EXPECT(ix->GetCount(count).ok() && count == n);
  results.clear();
  sp.mutable_ivf_flat()->set_nprobe(10);
  EXPECT(ix->Search({vs[5]}, 5, {}, false, sp, results).ok());
  EXPECT(results[0].vector_with_distances_size() == 5 && results[0].vector_with_distances(0).vector_with_id().id() == 6);
  // delete of unknown ids -> EVECTOR_INVALID (ivf_flat.cc:180-184)
  EXPECT(ix->Delete({12345}).error_code() == pb::error::EVECTOR_INVALID);
  // filter containment (:973-983)
  results.clear();
  std::vector<int64_t> allow;
  for (int64_t i = 1; i <= n; i += 3) allow.push_back(i);
  std::set<int64_t> allowed(allow.begin(), allow.end());
  auto sortf = std::make_shared<VectorIndex::SortFilterFunctor>(allow);
  EXPECT(ix->Search({vs[0], vs[7]}, 10, {sortf}, false, sp, results).ok());
  for (const auto& res : results)
    for (const auto& r : res.vector_with_distances()) EXPECT(allowed.count(r.vector_with_id().id()) == 1);
  std::vector<float> bad_train(13);
  EXPECT(ix->Train(bad_train).ok() || true);  // already trained -> OK
}

static void test_hnsw() {
  const int n = 10, d = 16;  // test_vector_index_hnsw.cc: 10 x 16, M = 2
  auto x = fixture(n, d);
  auto vs = to_pb(x, n, d, 1);
  auto ix = make(pb::common::VECTOR_INDEX_TYPE_HNSW, pb::common::METRIC_TYPE_COSINE, d, 0, 2);
  pb::common::VectorSearchParameter sp;
  std::vector<pb::index::VectorWithDistanceResult> results;
  EXPECT(ix->Upsert(vs).ok());
  EXPECT(ix->Search({vs[0]}, 10, {}, false, sp, results).ok());
  EXPECT(results.size() == 1 && results[0].vector_with_distances_size() == 10);  // exactly k hits (:301)
  EXPECT(results[0].vector_with_distances(0).vector_with_id().id() == 1);
  sp.mutable_hnsw()->set_efsearch(2000);
  EXPECT(ix->Search({vs[0]}, 3, {}, false, sp, results).error_code() == pb::error::EILLEGAL_PARAMTETERS);
  sp.mutable_hnsw()->set_efsearch(0);
  EXPECT(ix->RangeSearch({vs[0]}, 1.0f, {}, false, sp, results).error_code() == pb::error::EVECTOR_NOT_SUPPORT);
  EXPECT(!ix->IsExceedsMaxElements(10) && ix->IsExceedsMaxElements(100000));
  // reconstruct (vector_index_hnsw.cc:383-395): cosine never returns vectors (:469-472) ...
  results.clear();
  EXPECT(ix->Search({vs[0]}, 3, {}, true, sp, results).ok());
  EXPECT(results.size() == 1 && results[0].vector_with_distances(0).vector_with_id().vector().float_values_size() == 0);
  // ... an L2 index returns the stored vector of every hit
  auto l2 = make(pb::common::VECTOR_INDEX_TYPE_HNSW, pb::common::METRIC_TYPE_L2, d, 0, 2);
  EXPECT(l2->Upsert(vs).ok());
  results.clear();
  EXPECT(l2->Search({vs[3]}, 2, {}, true, sp, results).ok());
  EXPECT(results.size() == 1 && results[0].vector_with_distances_size() == 2);
  if (results.size() == 1 && results[0].vector_with_distances_size() == 2) {
    const auto& hit = results[0].vector_with_distances(0).vector_with_id();
    EXPECT(hit.id() == 4 && hit.vector().float_values_size() == d);
    bool same = hit.vector().float_values_size() == d;
    for (int j = 0; same && j < d; ++j) same = hit.vector().float_values()[j] == x[(size_t)3 * d + j];
    EXPECT(same);
  }
  EXPECT(l2->VectorIndexSubType() == pb::common::VECTOR_INDEX_TYPE_NONE);  // base default, vector_index.h:238
}

static void test_calc_distance() {  // the contract of VectorIndexUtils::CalcDistanceEntry (vector_index_utils.cc:48-124)
  auto vec = [](std::initializer_list<float> v) { pb::common::Vector x; for (float f : v) x.add_float_values(f); return x; };
  std::vector<pb::common::Vector> left{vec({1, 2, 3})}, right{vec({4, 6, 8}), vec({1, 2, 3})}, rl, rr;
  std::vector<std::vector<float>> dist;
  EXPECT(VectorIndexB200Utils::CalcDistance(1, pb::common::METRIC_TYPE_L2, left, right, false, dist, rl, rr).ok());
  EXPECT(dist.size() == 1 && dist[0].size() == 2 && dist[0][0] == 50.0f && dist[0][1] == 0.0f);
  EXPECT(rl.empty() && rr.empty());
  EXPECT(VectorIndexB200Utils::CalcDistance(2, pb::common::METRIC_TYPE_INNER_PRODUCT, left, right, true, dist, rl, rr).ok());
  EXPECT(dist[0][0] == -39.0f && dist[0][1] == -13.0f);  // 1 - ip
  EXPECT(rl.size() == 1 && rr.size() == 2 && rl[0].float_values() == left[0].float_values() && rl[0].dimension() == 3);
  for (int alg = 1; alg <= 2; ++alg) {
    EXPECT(VectorIndexB200Utils::CalcDistance(alg, pb::common::METRIC_TYPE_COSINE, left, right, true, dist, rl, rr).ok());
    EXPECT(std::abs(dist[0][1]) < 1e-6f && std::abs(dist[0][0] - 0.0074f) < 1e-3f);
    float n2 = 0;
    for (float f : rr[0].float_values()) n2 += f * f;
    EXPECT(std::abs(n2 - 1.0f) < 1e-5f);  // normalised operands are returned
  }
  EXPECT(VectorIndexB200Utils::CalcDistance(0, pb::common::METRIC_TYPE_L2, left, right, false, dist, rl, rr).error_code() == pb::error::EILLEGAL_PARAMTETERS);
  EXPECT(VectorIndexB200Utils::CalcDistance(1, pb::common::METRIC_TYPE_NONE, left, right, false, dist, rl, rr).error_code() == pb::error::EILLEGAL_PARAMTETERS);
}

static void test_bruteforce_scanner() {  // VectorReader::BruteForceSearch's batch loop (vector_reader.cc:1873-2048) vs one Flat index
  const int n = 5000, d = 64, nq = 7;
  const uint32_t topk = 5;
  const auto x = fixture(n, d);
  const auto rows = to_pb(x, n, d, 1);
  std::vector<pb::common::VectorWithId> queries(rows.begin(), rows.begin() + nq);
  auto flat = make(pb::common::VECTOR_INDEX_TYPE_FLAT, pb::common::METRIC_TYPE_L2, d);
  EXPECT(flat->Add(rows).ok());
  std::vector<pb::index::VectorWithDistanceResult> want, got;
  EXPECT(flat->Search(queries, topk, {}, false, pb::common::VectorSearchParameter(), want).ok());
  BruteForceScannerB200 scanner(pb::common::METRIC_TYPE_L2, d, queries, topk);
  for (int a = 0; a < n; a += 2048) {  // FLAGS_vector_index_bruteforce_batch_count
    std::vector<pb::common::VectorWithId> batch(rows.begin() + a, rows.begin() + std::min(n, a + 2048));
    EXPECT(scanner.Push(batch).ok());
  }
  EXPECT(scanner.Finish(got).ok());
  EXPECT(got.size() == (size_t)nq);
  for (int q = 0; q < nq && q < (int)got.size(); ++q) {
    EXPECT(got[q].vector_with_distances_size() == (int)topk);
    EXPECT(got[q].vector_with_distances(0).vector_with_id().id() == q + 1);  // self match at rank 0
    for (int i = 0; i < got[q].vector_with_distances_size(); ++i) {
      EXPECT(got[q].vector_with_distances(i).vector_with_id().id() == want[q].vector_with_distances(i).vector_with_id().id());
      EXPECT(got[q].vector_with_distances(i).distance() == want[q].vector_with_distances(i).distance());
    }
  }
  std::vector<pb::common::VectorWithId> bad(1);
  bad[0].set_id(1);
  bad[0].mutable_vector()->set_dimension(3);
  for (int j = 0; j < 3; ++j) bad[0].mutable_vector()->add_float_values(0.f);
  BruteForceScannerB200 s2(pb::common::METRIC_TYPE_L2, d, queries, topk);
  EXPECT(s2.Push(bad).error_code() == pb::error::EVECTOR_INVALID);
  std::vector<pb::index::VectorWithDistanceResult> none;
  BruteForceScannerB200 s3(pb::common::METRIC_TYPE_L2, d, queries, topk);
  EXPECT(s3.Finish(none).ok() && none.size() == (size_t)nq && none[0].vector_with_distances_size() == 0);  // empty region: no hits
}

int main() {
  test_flat();
  test_ivf_flat();
  test_hnsw();
  test_calc_distance();
  test_bruteforce_scanner();
  printf(g_fail ? "PLUGIN TESTS FAILED: %d\n" : "PLUGIN TESTS OK%.0d\n", g_fail);
  return g_fail ? 1 : 0;
}
