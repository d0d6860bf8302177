"""Streaming brute-force scan (b200vs_scan_*), the GPU counterpart of VectorReader::BruteForceSearch
(src/vector/vector_reader.cc:1873-2048): per-tile Flat search + running top-k merge must equal ONE exact Flat search
over the concatenation of the tiles — the oracle's flat_search — bit for bit."""
import numpy as np
import pytest

import b200vs
import oracle_lib
from gpu_util import require_gpu

pytestmark = pytest.mark.gpu

OM = {b200vs.L2: oracle_lib.L2, b200vs.IP: oracle_lib.IP, b200vs.COSINE: oracle_lib.COSINE}


def _stored(o, metric, xb):
    # Flat stores normalised rows for cosine (vector_index_flat.cc:155)
    return o.normalize_faiss(xb) if metric == b200vs.COSINE else xb


@pytest.mark.parametrize("metric", [b200vs.L2, b200vs.IP, b200vs.COSINE])
@pytest.mark.parametrize("tiles", [[2048, 2048, 777], [5, 1, 300], [4096]])
def test_scan_equals_flat_over_concatenation(metric, tiles):
    require_gpu()
    o = oracle_lib.load()
    d, nq, k = 96, 37, 10
    n = sum(tiles)
    rng = np.random.default_rng(n + metric)
    xb = rng.random((n, d), dtype=np.float32)
    ids = rng.permutation(n).astype(np.int64) + 1
    xb[n - 1] = xb[0]  # a duplicate row across tiles: tie broken by id
    xq = rng.random((nq, d), dtype=np.float32)
    sc = b200vs.BruteForceScan(metric, d, xq, k)
    a = 0
    for t in tiles:
        sc.push(xb[a:a + t], ids[a:a + t])
        a += t
    gd, gi = sc.finish()
    wd, wi = o.flat_search(OM[metric], _stored(o, metric, xb), ids, xq, k)
    assert np.array_equal(gi, wi)
    assert np.array_equal(gd.view(np.uint32), wd.view(np.uint32))


def test_scan_large_batch_tile_path_and_filters():
    """A batch big enough for the tensor-core tile scan inside each tile, with an id range and a negated id list."""
    require_gpu()
    o = oracle_lib.load()
    d, nq, k, n = 128, 256, 20, 6000
    rng = np.random.default_rng(5)
    xb = rng.random((n, d), dtype=np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    xq = rng.random((nq, d), dtype=np.float32)
    deny = np.sort(rng.choice(ids, 500, replace=False))
    for kw, of in (({"id_range": (100, 5000)}, {"id_range": (100, 5000)}), ({"sorted_ids": deny, "negate": True}, {"sorted_ids": deny, "negate": True})):
        sc = b200vs.BruteForceScan(b200vs.L2, d, xq, k, **kw)
        for a in range(0, n, 2048):
            sc.push(xb[a:a + 2048], ids[a:a + 2048])
        gd, gi = sc.finish()
        wd, wi = o.flat_search(oracle_lib.L2, xb, ids, xq, k, **of)
        assert np.array_equal(gi, wi)
        assert np.array_equal(gd.view(np.uint32), wd.view(np.uint32))


def test_scan_fewer_rows_than_k_and_no_rows():
    require_gpu()
    d, k = 16, 8
    xq = np.random.default_rng(0).random((3, d), dtype=np.float32)
    sc = b200vs.BruteForceScan(b200vs.L2, d, xq, k)
    gd, gi = sc.finish()  # nothing scanned: the reference returns no hits (vector_reader.cc:1908-1911)
    assert (gi == -1).all()
    sc = b200vs.BruteForceScan(b200vs.L2, d, xq, k)
    xb = np.random.default_rng(1).random((5, d), dtype=np.float32)
    sc.push(xb[:2], np.array([7, 9], np.int64))
    sc.push(xb[2:], np.array([1, 2, 3], np.int64))
    gd, gi = sc.finish()
    assert ((gi >= 0).sum(axis=1) == 5).all() and (gi[:, 5:] == -1).all()
    assert (np.diff(gd[:, :5], axis=1) >= 0).all()
