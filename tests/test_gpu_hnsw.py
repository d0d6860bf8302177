"""HNSW: the graph built by the engine's host builder equals the oracle's graph (same insertion order, seed 100), and
the GPU search (per-hop candidate-distance batches in the exact AVX-512 order) walks it exactly like the CPU search:
ids and distance bit patterns identical.  Contract checks mirror test/unit_test/vector/test_vector_index_hnsw.cc."""
import numpy as np
import pytest

import b200vs
import oracle_lib
from b200vs import COSINE, HNSW, IP, L2
from gpu_util import assert_same_results, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def pair(oracle, metric, n, d, M=16, efc=200, seed=0, batches=3):
    rng = np.random.default_rng(seed)
    xb = rng.random((n, d)).astype(np.float32)
    labels = np.arange(1000, 1000 + n, dtype=np.int64)
    ix = b200vs.Index(HNSW, metric, d, hnsw_m=M, hnsw_efc=efc, max_elements=n * 2)
    step = (n + batches - 1) // batches
    for a in range(0, n, step):
        ix.add(xb[a:a + step], labels[a:a + step])
    h = oracle_lib.OracleHnsw(oracle, metric, d, n, M, efc)
    h.add(xb, labels)
    return ix, h, xb, labels


@pytest.mark.parametrize("metric", [L2, IP, COSINE])
@pytest.mark.parametrize("n,d,M,k,ef", [(10, 16, 2, 10, 0), (3000, 32, 16, 10, 128), (2000, 128, 8, 5, 64), (1500, 100, 16, 30, 16)])
def test_graph_and_search_equal_oracle(oracle, metric, n, d, M, k, ef):
    ix, h, xb, labels = pair(oracle, metric, n, d, M=M, seed=n + d)
    assert ix.get_count() == n
    assert np.array_equal(ix.get_trained_state(), h.export()), "host-built graph differs from the oracle's"
    xq = np.random.default_rng(7).random((40, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, k, efsearch=ef)
    Do, Io, nd, nh = h.search(xq, k, ef=ef, nthreads=8)
    assert_same_results(Dg, Ig, Do, Io)
    if n >= k:
        assert (Ig >= 1000).all()  # exactly k hits (test_vector_index_hnsw.cc:301)


def test_load_oracle_graph_then_search(oracle):
    rng = np.random.default_rng(1)
    n, d = 4000, 64
    xb = rng.random((n, d)).astype(np.float32)
    labels = np.arange(n, dtype=np.int64) * 2 + 1
    h = oracle_lib.OracleHnsw(oracle, COSINE, d, n, 16, 200)
    h.add(xb, labels)
    ix = b200vs.Index(HNSW, COSINE, d, hnsw_m=16, hnsw_efc=200, max_elements=n)
    ix.set_trained_state(h.export())
    xq = rng.random((64, d)).astype(np.float32)
    for kw in (dict(), dict(id_range=(100, 3000)), dict(sorted_ids=np.arange(1, 8000, 6))):
        Dg, Ig = ix.search(xq, 10, efsearch=128, **kw)
        Do, Io, _, _ = h.search(xq, 10, ef=128, nthreads=8, **kw)
        assert_same_results(Dg, Ig, Do, Io)
    # self query: itself at rank 0 for cosine (test_vector_index_recall_flat.cc:222-229 analogue)
    Dg, Ig = ix.search(xb[:20], 3, efsearch=64)
    assert (Ig[:, 0] == labels[:20]).all()


def test_status_codes_sticky_ef_and_deletes(oracle):
    ix, h, xb, labels = pair(oracle, L2, 1200, 24, seed=3)
    xq = np.random.default_rng(9).random((10, 24)).astype(np.float32)
    for bad in (-1, 1025):  # hnsw.cc:332-336
        with pytest.raises(b200vs.B200VSError) as e:
            ix.search(xq, 5, efsearch=bad)
        assert e.value.code == b200vs.EILLEGAL_PARAMETERS
    with pytest.raises(b200vs.B200VSError) as e:
        ix.range_search(xq, 1.0, 10)  # hnsw.cc:487-493
    assert e.value.code == b200vs.EVECTOR_NOT_SUPPORT
    with pytest.raises(b200vs.B200VSError) as e:
        ix.search(np.zeros((0, 24), np.float32), 5)
    assert e.value.code == b200vs.EILLEGAL_PARAMETERS
    # ef is sticky: efsearch=0 keeps the last value (hnsw.cc:426-428)
    D1, I1 = ix.search(xq, 5, efsearch=200)
    D2, I2 = ix.search(xq, 5, efsearch=0)
    Do, Io, _, _ = h.search(xq, 5, ef=200)
    assert_same_results(D2, I2, Do, Io)
    # markDelete: deleted labels are traversed but never returned
    dead = labels[::3]
    assert ix.delete(dead) == len(dead)
    assert ix.get_count() == 1200 - len(dead) and ix.get_deleted_count() == len(dead)
    D3, I3 = ix.search(xq, 10, efsearch=100)
    assert not (set(I3.ravel()) & set(dead))
    assert (I3 >= 0).all()


@pytest.mark.parametrize("metric", [L2, COSINE])
@pytest.mark.parametrize("M", [24, 40, 64])
def test_wide_link_lists(oracle, metric, M):
    """nlinks > 16 (level-0 lists of up to 2 * nlinks = 128 neighbours): the reference constructor takes any nlinks
    (vector_index_hnsw.cc:135-184); the kernel walks a list 32 neighbours at a time."""
    ix, h, xb, labels = pair(oracle, metric, 2500, 48, M=M, efc=120, seed=M)
    assert np.array_equal(ix.get_trained_state(), h.export())
    xq = np.random.default_rng(M).random((33, 48)).astype(np.float32)
    for k, ef in ((10, 64), (50, 0), (1, 300)):
        Dg, Ig = ix.search(xq, k, efsearch=ef)
        Do, Io, _, _ = h.search(xq, k, ef=ef, nthreads=8)
        assert_same_results(Dg, Ig, Do, Io)


def test_reconstruct_returns_stored_vectors(oracle):
    """hnswlib getDataByLabel behind Search(reconstruct = true), vector_index_hnsw.cc:383-395."""
    ix, h, xb, labels = pair(oracle, L2, 800, 20, seed=5)
    D, I = ix.search(xb[:5], 3, efsearch=50)
    vec, found = ix.reconstruct(I.ravel())
    assert found.all()
    pos = I.ravel() - 1000
    assert np.array_equal(vec, xb[pos])
    ix.delete(labels[:1])
    _, found = ix.reconstruct(np.array([labels[0], 424242], dtype=np.int64))
    assert not found.any()
    # cosine stores the hnsw-normalised rows (the reference never reconstructs for cosine, :469-472)
    ic, hc, xc, lc = pair(oracle, COSINE, 300, 16, seed=6)
    vec, found = ic.reconstruct(lc[:4])
    assert found.all() and np.allclose(np.linalg.norm(vec, axis=1), 1.0, atol=1e-5)


def test_concurrent_build_graph_is_searched_exactly(oracle):
    """hnsw_build_threads > 1 = the reference's concurrent addPoint (vector_index_hnsw.cc:229-243): the graph depends on thread
    timing, so parity is defined on the graph itself — the oracle adopts it (oracle_hnsw_import) and both sides must walk it
    identically; the graph must also be a sound HNSW (every live row reachable in practice: self queries come back first)."""
    rng = np.random.default_rng(8)
    n, d, M, efc = 20000, 64, 16, 100
    xb = rng.random((n, d)).astype(np.float32)
    labels = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(HNSW, L2, d, hnsw_m=M, hnsw_efc=efc, max_elements=n, hnsw_build_threads=16)
    for a in range(0, n, 5000):
        ix.add(xb[a:a + 5000], labels[a:a + 5000])
    assert ix.get_count() == n
    h = oracle_lib.OracleHnsw(oracle, L2, d, n, M, efc)
    h.load(ix.get_trained_state())
    xq = rng.random((100, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, 10, efsearch=128)
    Do, Io, _, _ = h.search(xq, 10, ef=128, nthreads=8)
    assert_same_results(Dg, Ig, Do, Io)
    Dg, Ig = ix.search(xb[:200], 1, efsearch=64)
    assert (Ig[:, 0] == labels[:200]).mean() > 0.9
    Df, If = oracle.flat_search(oracle_lib.L2, xb, labels, xq, 10, nthreads=8)
    Dg, Ig = ix.search(xq, 10, efsearch=200)
    assert np.mean([len(set(a) & set(b)) / 10 for a, b in zip(Ig, If)]) > 0.9  # a healthy graph: recall vs brute force
