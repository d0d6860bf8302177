"""IVF-PQ: CUDA LUT scan vs the oracle on a SHARED trained state (coarse centroids + PQ codebooks from the oracle).
Codes produced by the GPU encoder must equal the oracle's; distances are compared within the north-star tolerance
(1e-4 relative) and recall@k within 1e-3 — in practice they are bit-identical because both sides sum the LUT in the
same order."""
import numpy as np
import pytest

import b200vs
from b200vs import COSINE, IP, IVF_PQ, L2
from gpu_util import assert_same_results, recall, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def build(oracle, metric, n, d, nlist, M, seed=0):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    stored = oracle.normalize_faiss(xb) if metric == COSINE else xb
    cent = oracle.kmeans(metric, stored, nlist)
    asg = oracle.assign(metric, stored, cent)
    cb = oracle.pq_train(stored - cent[asg], M, niter=8)
    ix = b200vs.Index(IVF_PQ, metric, d, nlist=nlist, pq_m=M, pq_nbits=8)
    ix.set_trained_state(b200vs.ivfpq_state_blob(cent, cb, metric))
    for a in range(0, n, 5000):
        ix.add(xb[a:a + 5000], ids[a:a + 5000])
    return ix, xb, stored, ids, cent, cb, asg


@pytest.mark.parametrize("metric", [L2, IP, COSINE])
@pytest.mark.parametrize("n,d,nlist,M,nprobe,nq,k", [(12000, 64, 32, 8, 8, 20, 10), (9000, 96, 16, 12, 16, 7, 100), (6000, 128, 8, 32, 4, 64, 5)])
def test_ivfpq_matches_oracle(oracle, metric, n, d, nlist, M, nprobe, nq, k):
    ix, xb, stored, ids, cent, cb, asg = build(oracle, metric, n, d, nlist, M, seed=n + M)
    assert ix.get_count() == n
    off, _, codes, lids = ix.export_lists(nlist, with_vectors=False, code_size=M)
    # GPU encoder == oracle encoder (row by row, matched through ids)
    want = oracle.ivfpq_encode(cb, cent, stored, asg)
    pos = np.empty(n + 1, np.int64)
    pos[lids] = np.arange(n)
    assert np.array_equal(codes[pos[ids]], want)
    assert np.array_equal(np.diff(off), np.bincount(asg, minlength=nlist))
    xq = np.random.default_rng(1).standard_normal((nq, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, k, nprobe=nprobe)
    Do, Io = oracle.ivfpq_search(metric, cent, cb, off, codes, lids, xq, k, nprobe, nthreads=8)
    assert recall(Ig, Io) >= 1 - 1e-3
    valid = (Io >= 0) & (Ig == Io)
    assert np.allclose(Dg[valid], Do[valid], rtol=1e-4, atol=1e-5)
    assert_same_results(Dg, Ig, Do, Io)  # stronger than the gate: same summation order on both sides


def test_flat_fallback_when_training_set_is_small(oracle):
    # VectorIndexIvfPq::Train: fewer than max(256*nlist, 256*2^nbits) vectors -> inner Flat index (ivf_pq.cc:339-353)
    rng = np.random.default_rng(2)
    xb = rng.random((1000, 64)).astype(np.float32)  # the reference's small fixture: 1000 x 64 (test_vector_index_ivf_pq.cc)
    ids = np.arange(1, 1001, dtype=np.int64)
    ix = b200vs.Index(IVF_PQ, L2, 64, nlist=100, pq_m=8, pq_nbits=8)
    assert not ix.is_trained()
    D, I = ix.search(xb[:2], 3)  # not trained -> OK + empty (ivf_pq.cc:159-163)
    assert (I == -1).all()
    with pytest.raises(b200vs.B200VSError) as e:
        ix.add(xb, ids)
    assert e.value.code == b200vs.EVECTOR_NOT_TRAIN
    ix.train(xb)
    assert ix.is_trained()
    ix.add(xb, ids)
    xq = rng.random((5, 64)).astype(np.float32)
    D, I = ix.search(xq, 10)
    Do, Io = oracle.flat_search(L2, xb, ids, xq, 10)
    assert_same_results(D, I, Do, Io)
    assert ix.delete(np.array([424242])) == 0  # Flat semantics inside


def test_gpu_trained_ivfpq_recall(oracle):
    # full GPU training path (coarse k-means + 256-centroid sub-quantisers): needs >= 65536 training vectors
    rng = np.random.default_rng(3)
    n, d, nlist, M = 70000, 32, 64, 8
    centers = rng.standard_normal((500, d)).astype(np.float32)
    xb = (centers[rng.integers(0, 500, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    ids = np.arange(n, dtype=np.int64)
    ix = b200vs.Index(IVF_PQ, L2, d, nlist=nlist, pq_m=M, pq_nbits=8)
    ix.train(xb)
    assert ix.is_trained()
    ix.add(xb, ids)
    xq = xb[:100] + 0.01 * rng.standard_normal((100, d)).astype(np.float32)
    D, I = ix.search(xq, 10, nprobe=16)
    Df, If = oracle.flat_search(L2, xb, ids, xq, 10, nthreads=8)
    assert recall(I, If) > 0.3  # 8-byte codes on noisy 32-d data: a sanity floor, not a quality claim
    assert (I[:, 0] == ids[:100]).mean() > 0.8  # the perturbed source vector is (almost always) the nearest code


def test_upsert_delete_filters(oracle):
    ix, xb, stored, ids, cent, cb, asg = build(oracle, IP, 8000, 64, 16, 8, seed=4)
    ix.delete(ids[::5])
    xq = np.random.default_rng(5).standard_normal((6, 64)).astype(np.float32)
    off, _, codes, lids = ix.export_lists(16, with_vectors=False, code_size=8)
    assert off[-1] == ix.get_count() == 8000 - len(ids[::5])
    for kw in (dict(), dict(id_range=(100, 3000)), dict(sorted_ids=np.arange(2, 8000, 3))):
        Dg, Ig = ix.search(xq, 10, nprobe=6, **kw)
        Do, Io = oracle.ivfpq_search(IP, cent, cb, off, codes, lids, xq, 10, 6, **kw)
        assert_same_results(Dg, Ig, Do, Io)
    with pytest.raises(b200vs.B200VSError) as e:
        ix.delete(np.array([999999]))
    assert e.value.code == b200vs.EVECTOR_INVALID


@pytest.mark.parametrize("metric", [L2, IP, COSINE])
def test_range_search_on_codes(oracle, metric):
    """VectorIndexRawIvfPq::RangeSearch (vector_index_raw_ivf_pq.cc:212-278): the LUT scan with a radius instead of top-k.
    Checked against the oracle's LUT distances of every probed row (top-k with k = all candidates), thresholded on the host."""
    n, d, nlist, M, nprobe, nq = 9000, 64, 16, 8, 6, 9
    ix, xb, stored, ids, cent, cb, asg = build(oracle, metric, n, d, nlist, M, seed=11)
    off, _, codes, lids = ix.export_lists(nlist, with_vectors=False, code_size=M)
    xq = np.random.default_rng(12).standard_normal((nq, d)).astype(np.float32)
    kall = 4096
    Do, Io = oracle.ivfpq_search(metric, cent, cb, off, codes, lids, xq, kall, nprobe, nthreads=8)  # API distances, ascending
    # a radius that keeps a few dozen hits per query
    radius = float(np.sort(Do[Io >= 0])[nq * 40])
    D, I, C = ix.range_search(xq, radius, 256, nprobe=nprobe)
    for q in range(nq):
        valid = Io[q] >= 0
        # faiss range_search is strict on the RAW metric: L2 dis < radius ; IP ip > 1 - radius  <=>  1 - ip < radius up to rounding of (1 - x)
        raw = Do[q][valid] if metric == L2 else 1.0 - Do[q][valid]
        keep = raw < radius if metric == L2 else raw > np.float32(1.0) - np.float32(radius)
        want_i, want_d = Io[q][valid][keep], Do[q][valid][keep]
        assert C[q] == min(len(want_i), 256)
        assert np.array_equal(I[q, :C[q]], want_i[:C[q]])
        assert np.array_equal(D[q, :C[q]].view(np.uint32), want_d[:C[q]].view(np.uint32))
        assert (I[q, C[q]:] == -1).all()
    # filters ride along; an untrained index answers with empty results
    D2, I2, C2 = ix.range_search(xq, radius, 256, nprobe=nprobe, id_range=(1, 2000))
    assert ((I2 < 2000) | (I2 == -1)).all() and (C2 <= C).all()
    fresh = b200vs.Index(IVF_PQ, metric, d, nlist=nlist, pq_m=M, pq_nbits=8)
    D3, I3, C3 = fresh.range_search(xq, radius, 16)
    assert (C3 == 0).all() and (I3 == -1).all()
