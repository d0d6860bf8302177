"""Save / Load round trip (VectorIndex::Save/Load, src/vector/vector_index.h:168-170): a loaded index answers exactly like
the saved one.  Own container format."""
import os

import numpy as np
import pytest

import b200vs
from b200vs import COSINE, FLAT, HNSW, IVF_FLAT, IVF_PQ, L2
from gpu_util import assert_same_results, require_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,metric", [(FLAT, L2), (FLAT, COSINE), (IVF_FLAT, COSINE), (IVF_FLAT, L2), (HNSW, COSINE)])
def test_save_load_round_trip(tmp_path, kind, metric):
    require_gpu()
    rng = np.random.default_rng(kind * 10 + metric)
    n, d = 6000, 48
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64) * 3
    kw = dict(nlist=16, hnsw_m=8, hnsw_efc=100, max_elements=2 * n)
    a = b200vs.Index(kind, metric, d, **kw)
    if kind == IVF_FLAT:
        a.train(xb)
    a.add(xb, ids)
    a.delete(ids[::9])
    xq = rng.random((40, d)).astype(np.float32)
    Da, Ia = a.search(xq, 10, nprobe=6, efsearch=64)
    path = os.path.join(tmp_path, "idx.b2vs")
    a.save(path)
    b = b200vs.Index(kind, metric, d, **kw)
    b.load(path)
    assert b.get_count() == a.get_count() or kind == HNSW  # HNSW keeps tombstoned nodes in the graph
    Db, Ib = b.search(xq, 10, nprobe=6, efsearch=64)
    if kind == HNSW:
        assert not (set(Ib.ravel()) & set(ids[::9]))
    else:
        assert_same_results(Da, Ia, Db, Ib)
    with pytest.raises(b200vs.B200VSError):
        b.load(path)  # load into a non-empty index is refused


@pytest.mark.parametrize("metric", [L2, b200vs.IP])
@pytest.mark.parametrize("n_train", [3000, 70000])
def test_ivf_pq_save_load_round_trip(tmp_path, metric, n_train):
    """IVF-PQ: 3000 training vectors -> the inner Flat index (vector_index_ivf_pq.cc:339-353); 70000 >= 256 * 2^nbits ->
    real PQ lists.  Codes are stored and restored as they are, so the reloaded index answers identically."""
    require_gpu()
    rng = np.random.default_rng(n_train + metric)
    d, M, nlist = 32, 8, 16
    xb = rng.random((n_train, d)).astype(np.float32)
    ids = np.arange(1, n_train + 1, dtype=np.int64) * 2
    kw = dict(nlist=nlist, pq_m=M, pq_nbits=8)
    a = b200vs.Index(IVF_PQ, metric, d, **kw)
    a.train(xb)
    a.add(xb, ids)
    a.delete(ids[::7])
    xq = rng.random((50, d)).astype(np.float32)
    Da, Ia = a.search(xq, 10, nprobe=8)
    path = os.path.join(tmp_path, "pq.b2vs")
    a.save(path)
    b = b200vs.Index(IVF_PQ, metric, d, **kw)
    assert not b.is_trained()
    b.load(path)
    assert b.is_trained() and b.get_count() == a.get_count()
    Db, Ib = b.search(xq, 10, nprobe=8)
    assert_same_results(Da, Ia, Db, Ib)
    # the reloaded index keeps working as an index: upsert + delete, same answers on both sides
    extra = rng.random((100, d)).astype(np.float32)
    eid = np.arange(10**6, 10**6 + 100, dtype=np.int64)
    a.add(extra, eid)
    b.add(extra, eid)
    Da, Ia = a.search(xq, 10, nprobe=8)
    Db, Ib = b.search(xq, 10, nprobe=8)
    assert_same_results(Da, Ia, Db, Ib)
    with pytest.raises(b200vs.B200VSError):
        b.load(path)  # load into a trained index is refused


def test_ivf_pq_save_untrained(tmp_path):
    require_gpu()
    ix = b200vs.Index(IVF_PQ, L2, 32, nlist=4, pq_m=4, pq_nbits=8)
    path = os.path.join(tmp_path, "empty.b2vs")
    ix.save(path)
    other = b200vs.Index(IVF_PQ, L2, 32, nlist=4, pq_m=4, pq_nbits=8)
    other.load(path)
    assert not other.is_trained() and other.get_count() == 0
