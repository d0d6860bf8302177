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


def test_save_not_supported_for_pq():
    require_gpu()
    ix = b200vs.Index(IVF_PQ, L2, 32, nlist=4, pq_m=4, pq_nbits=8)
    with pytest.raises(b200vs.B200VSError) as e:
        ix.save("/tmp/never.b2vs")
    assert e.value.code == b200vs.EVECTOR_NOT_SUPPORT
