"""Parity of the CUDA Flat path (through the C ABI) with the CPU oracle: bit-exact ids AND distances, plus the
behavioural contract the reference's unit tests pin (test/unit_test/vector/test_vector_index_flat.cc)."""
import numpy as np
import pytest

import b200vs
from b200vs import COSINE, FLAT, IP, L2
from gpu_util import assert_same_results, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def stored_rows(oracle, metric, xb):
    return oracle.normalize_faiss(xb) if metric == COSINE else xb


@pytest.mark.parametrize("metric", [L2, IP, COSINE])
@pytest.mark.parametrize("n,d,nq,k", [(10, 8, 3, 3), (1000, 128, 17, 10), (5000, 768, 8, 10), (777, 100, 5, 100),
                                      (300, 13, 4, 7), (4000, 1536, 3, 10), (2000, 33, 40, 1)])
def test_flat_bit_exact_vs_oracle(oracle, metric, n, d, nq, k):
    rng = np.random.default_rng(n + d)
    xb = rng.random((n, d)).astype(np.float32)
    xq = rng.random((nq, d)).astype(np.float32)
    ids = (np.arange(n, dtype=np.int64) * 7 + 11)
    ix = b200vs.Index(FLAT, metric, d)
    ix.add(xb, ids)
    assert ix.get_count() == n
    Dg, Ig = ix.search(xq, k)
    Do, Io = oracle.flat_search(metric, stored_rows(oracle, metric, xb), ids, xq, k, nthreads=8)
    assert_same_results(Dg, Ig, Do, Io)


def test_reference_fixture_golden(oracle):
    # the reference's own Flat fixture: 10 x 8, default-seeded mt19937 (test_vector_index_flat.cc:491-500)
    xb = oracle.fixture(10, 8)
    ids = np.arange(1, 11, dtype=np.int64)
    ix = b200vs.Index(FLAT, L2, 8)
    ix.add(xb, ids)
    D, I = ix.search(xb[:2], 3)
    Do, Io = oracle.flat_search(L2, xb, ids, xb[:2], 3)
    assert_same_results(D, I, Do, Io)
    assert I[0, 0] == 1 and I[1, 0] == 2 and D[0, 0] == 0
    # SURVEY §0 KAT: L2sqr(row0,row1) = 0x401f384a
    D1, I1 = ix.search(xb[:1], 10)
    j = list(I1[0]).index(2)
    assert D1[0, j].view(np.uint32) == 0x401F384A


def test_status_codes_and_edge_cases():
    ix = b200vs.Index(FLAT, L2, 8)
    x = np.random.default_rng(0).random((10, 8)).astype(np.float32)
    # empty add / empty search -> EILLEGAL_PARAMTETERS (flat.cc:123-125, :208-210)
    with pytest.raises(b200vs.B200VSError) as e:
        ix.add(np.zeros((0, 8), np.float32), np.zeros(0, np.int64))
    assert e.value.code == b200vs.EILLEGAL_PARAMETERS
    with pytest.raises(b200vs.B200VSError) as e:
        ix.search(np.zeros((0, 8), np.float32), 3)
    assert e.value.code == b200vs.EILLEGAL_PARAMETERS
    # search on an empty index: OK, nothing found
    D, I = ix.search(x[:2], 3)
    assert (I == -1).all()
    ix.add(x, np.arange(1, 11))
    # topk == 0 -> OK, results untouched (flat.cc:212; test_vector_index_flat.cc:873-909)
    D, I = ix.search(x[:2], 0)
    assert D.shape == (2, 0)
    # duplicate ids inside one batch -> EVECTOR_ID_DUPLICATED (flat.cc:131-136)
    with pytest.raises(b200vs.B200VSError) as e:
        ix.add(x[:2], np.array([50, 50]))
    assert e.value.code == b200vs.EVECTOR_ID_DUPLICATED
    # k larger than the index -> -1 padded
    D, I = ix.search(x[:1], 20)
    assert (I[0, :10] >= 1).all() and (I[0, 10:] == -1).all()
    # deleting unknown ids is OK for Flat (flat.cc:171-203)
    assert ix.delete(np.array([12345])) == 0
    # range search is supported, HNSW-only NOT_SUPPORT is tested elsewhere
    D, I, C = ix.range_search(x[:1], 1e9, 16)
    assert C[0] == 10


def test_add_replaces_existing_ids_and_delete(oracle):
    rng = np.random.default_rng(1)
    d = 32
    xb = rng.random((500, d)).astype(np.float32)
    ids = np.arange(1, 501, dtype=np.int64)
    ix = b200vs.Index(FLAT, L2, d)
    ix.add(xb, ids)
    # re-adding existing ids replaces them (flat.cc:141-156): count unchanged, new vectors searched
    xnew = rng.random((100, d)).astype(np.float32)
    ix.add(xnew, ids[:100])
    assert ix.get_count() == 500
    cur = xb.copy()
    cur[:100] = xnew
    xq = rng.random((6, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, 10)
    Do, Io = oracle.flat_search(L2, cur, ids, xq, 10)
    assert_same_results(Dg, Ig, Do, Io)
    # delete, then search never returns deleted ids
    assert ix.delete(ids[200:450]) == 250
    assert ix.get_count() == 250
    mids = ids.copy()
    mids[200:450] = -1
    Dg, Ig = ix.search(xq, 10)
    Do, Io = oracle.flat_search(L2, cur, mids, xq, 10)
    assert_same_results(Dg, Ig, Do, Io)
    # heavy deletion triggers compaction; results unchanged
    assert ix.delete(ids[:150]) == 150
    mids[:150] = -1
    Dg, Ig = ix.search(xq, 10)
    Do, Io = oracle.flat_search(L2, cur, mids, xq, 10)
    assert_same_results(Dg, Ig, Do, Io)
    assert ix.get_count() == 100


@pytest.mark.parametrize("metric", [L2, IP])
def test_filters_match_oracle(oracle, metric):
    rng = np.random.default_rng(2)
    n, d = 3000, 64
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    xq = rng.random((5, d)).astype(np.float32)
    ix = b200vs.Index(FLAT, metric, d)
    ix.add(xb, ids)
    allow = np.sort(rng.choice(ids, 200, replace=False))
    for kw in (dict(id_range=(100, 400)), dict(sorted_ids=allow), dict(sorted_ids=allow, negate=True),
               dict(id_range=(1, 2000), sorted_ids=allow)):
        Dg, Ig = ix.search(xq, 20, **kw)
        Do, Io = oracle.flat_search(metric, xb, ids, xq, 20, **kw)
        assert_same_results(Dg, Ig, Do, Io)
    # containment, as the reference asserts (test_vector_index_flat_search_param.cc:274-283)
    Dg, Ig = ix.search(xq, 20, sorted_ids=allow)
    assert set(Ig.ravel()) <= set(allow) | {-1}


def test_ties_broken_by_id():
    xb = np.ones((6, 4), np.float32)
    ids = np.array([9, 3, 7, 1, 5, 2], np.int64)
    ix = b200vs.Index(FLAT, L2, 4)
    ix.add(xb, ids)
    D, I = ix.search(np.zeros((1, 4), np.float32), 4)
    assert list(I[0]) == [1, 2, 3, 5]


def test_large_k_and_batch(oracle):
    rng = np.random.default_rng(3)
    n, d = 6000, 24
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.int64) + 1
    ix = b200vs.Index(FLAT, L2, d)
    ix.add(xb, ids)
    xq = rng.random((3, d)).astype(np.float32)
    for k in (1024, 4096):  # RPC limit top_n <= 4096 (index_service.cc:197-211)
        Dg, Ig = ix.search(xq, k)
        Do, Io = oracle.flat_search(L2, xb, ids, xq, k, nthreads=4)
        assert_same_results(Dg, Ig, Do, Io)
    xq = rng.random((1024, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, 10)
    Do, Io = oracle.flat_search(L2, xb, ids, xq, 10, nthreads=8)
    assert_same_results(Dg, Ig, Do, Io)


def test_range_search_matches_oracle_topk(oracle):
    rng = np.random.default_rng(4)
    n, d = 2000, 16
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(FLAT, L2, d)
    ix.add(xb, ids)
    xq = rng.random((4, d)).astype(np.float32)
    Do, Io = oracle.flat_search(L2, xb, ids, xq, n)
    radius = float(Do[:, 40].mean())
    D, I, C = ix.range_search(xq, radius, 256)
    for q in range(4):
        want = Io[q][Do[q] < radius]
        assert C[q] == len(want) and list(I[q, :C[q]]) == list(want)
        assert (I[q, C[q]:] == -1).all()


def test_full_size_config1_properties(oracle):
    # BASELINE config 1: Flat L2, 100K x 128, top-10, batch=1 — size-independent properties + sampled oracle check
    rng = np.random.default_rng(1234)
    n, d = 100_000, 128
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(FLAT, L2, d)
    for a in range(0, n, 32768):  # kBuildVectorIndexBatchSize, constant.h:173
        ix.add(xb[a:a + 32768], ids[a:a + 32768])
    q = np.random.default_rng(4321).random((1, d)).astype(np.float32)
    D, I = ix.search(q, 10)
    Do, Io = oracle.flat_search(L2, xb, ids, q, 10, nthreads=1)
    assert_same_results(D, I, Do, Io)
    # self queries come back at rank 0 with distance 0; results ascending; idempotent
    sel = rng.integers(0, n, 64)
    D, I = ix.search(xb[sel], 10)
    assert (I[:, 0] == ids[sel]).all() and (D[:, 0] == 0).all() and (np.diff(D, axis=1) >= 0).all()
    D2, I2 = ix.search(xb[sel], 10)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
