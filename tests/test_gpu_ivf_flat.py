"""Parity of the CUDA IVF-Flat path with the CPU oracle on a SHARED trained state (centroids from the oracle's
k-means are loaded into the GPU index; SURVEY.md §7 "non-deterministic reference builds")."""
import numpy as np
import pytest

import b200vs
from b200vs import COSINE, IP, IVF_FLAT, L2
from gpu_util import assert_same_results, recall, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def build_pair(oracle, metric, n, d, nlist, seed=0, batches=3):
    rng = np.random.default_rng(seed)
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    stored = oracle.normalize_faiss(xb) if metric == COSINE else xb
    cent = oracle.kmeans(metric, stored, nlist)
    ix = b200vs.Index(IVF_FLAT, metric, d, nlist=nlist)
    ix.set_trained_state(b200vs.ivf_state_blob(cent, metric))
    assert ix.is_trained()
    step = (n + batches - 1) // batches
    for a in range(0, n, step):
        ix.add(xb[a:a + step], ids[a:a + step])
    return ix, xb, stored, ids, cent


@pytest.mark.parametrize("metric", [L2, IP, COSINE])
@pytest.mark.parametrize("n,d,nlist,nprobe,nq,k", [(100, 8, 10, 3, 4, 5), (20000, 128, 64, 8, 33, 10),
                                                   (30000, 768, 32, 4, 16, 10), (5000, 48, 16, 16, 9, 100)])
def test_ivfflat_bit_exact_vs_oracle(oracle, metric, n, d, nlist, nprobe, nq, k):
    ix, xb, stored, ids, cent = build_pair(oracle, metric, n, d, nlist, seed=n + d)
    assert ix.get_count() == n
    # the GPU's add-time assignment equals the oracle's
    off, lx, _, lids = ix.export_lists(nlist)
    asg = oracle.assign(metric, stored, cent)
    assert np.array_equal(np.diff(off), np.bincount(asg, minlength=nlist))
    for l in range(nlist):
        assert set(lids[off[l]:off[l + 1]]) == set(ids[asg == l])
    xq = np.random.default_rng(99).random((nq, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, k, nprobe=nprobe)
    Do, Io = oracle.ivfflat_search(metric, cent, off, lx, lids, xq, k, nprobe, nthreads=8)
    assert_same_results(Dg, Ig, Do, Io)


def test_nprobe_default_and_clamp(oracle):
    ix, xb, stored, ids, cent = build_pair(oracle, L2, 4000, 32, 100, seed=7)
    off, lx, _, lids = ix.export_lists(100)
    xq = np.random.default_rng(1).random((5, 32)).astype(np.float32)
    D0, I0 = ix.search(xq, 10)  # nprobe <= 0 -> 80 (constant.h:178)
    Do, Io = oracle.ivfflat_search(L2, cent, off, lx, lids, xq, 10, 80)
    assert_same_results(D0, I0, Do, Io)
    D1, I1 = ix.search(xq, 10, nprobe=5000)  # clamped to nlist (ivf_flat.cc:234) == exhaustive
    Df, If = oracle.flat_search(L2, xb, ids, xq, 10)
    assert_same_results(D1, I1, Df, If)


def test_untrained_behaviour_and_status_codes():
    ix = b200vs.Index(IVF_FLAT, L2, 8, nlist=10)
    assert not ix.is_trained()
    x = np.random.default_rng(0).random((100, 8)).astype(np.float32)
    # untrained search -> OK + empty results (ivf_flat.cc:224-227; test_vector_index_ivf_flat.cc:286)
    D, I = ix.search(x[:3], 5)
    assert (I == -1).all()
    # untrained add -> EVECTOR_NOT_TRAIN (the C++ wrapper trains with the batch and retries, ivf_flat.cc:133-150)
    with pytest.raises(b200vs.B200VSError) as e:
        ix.add(x, np.arange(100))
    assert e.value.code == b200vs.EVECTOR_NOT_TRAIN
    # untrained delete -> OK (ivf_flat.cc:174-177)
    assert ix.delete(np.array([1, 2])) == 0
    ix.train(x)
    assert ix.is_trained()
    ix.add(x, np.arange(100))
    # delete of ids that are not there -> EVECTOR_INVALID (ivf_flat.cc:180-184)
    with pytest.raises(b200vs.B200VSError) as e:
        ix.delete(np.array([555]))
    assert e.value.code == b200vs.EVECTOR_INVALID
    with pytest.raises(b200vs.B200VSError) as e:
        ix.search(np.zeros((0, 8), np.float32), 3)
    assert e.value.code == b200vs.EILLEGAL_PARAMETERS


def test_nlist_degenerates_to_one_when_data_is_small(oracle):
    # ivf_flat.cc:676-680; reference fixture is 100 x 8 with nlist = 10 (test_vector_index_ivf_flat.cc:108-110)
    ix = b200vs.Index(IVF_FLAT, L2, 8, nlist=2048)
    x = oracle.fixture(100, 8)
    ids = np.arange(1, 101, dtype=np.int64)
    ix.train(x)
    ix.add(x, ids)
    D, I = ix.search(x[:4], 5, nprobe=1)
    Do, Io = oracle.flat_search(L2, x, ids, x[:4], 5)
    assert_same_results(D, I, Do, Io)


def test_upsert_delete_and_filters(oracle):
    ix, xb, stored, ids, cent = build_pair(oracle, L2, 6000, 40, 24, seed=3)
    rng = np.random.default_rng(5)
    xnew = rng.random((500, 40)).astype(np.float32)
    ix.upsert(xnew, ids[:500])  # remove_ids then add (ivf_flat.cc:115-121)
    assert ix.get_count() == 6000
    cur = xb.copy()
    cur[:500] = xnew
    assert ix.delete(ids[1000:4500]) == 3500  # also crosses the compaction threshold
    mids = ids.copy()
    mids[1000:4500] = -1
    xq = rng.random((8, 40)).astype(np.float32)
    off, lx, _, lids = ix.export_lists(24)
    assert ix.get_count() == 2500 and off[-1] == 2500
    for kw in (dict(), dict(id_range=(100, 900)), dict(sorted_ids=np.arange(1, 6000, 3)), dict(sorted_ids=np.arange(1, 6000, 3), negate=True)):
        Dg, Ig = ix.search(xq, 10, nprobe=6, **kw)
        Do, Io = oracle.ivfflat_search(L2, cent, off, lx, lids, xq, 10, 6, **kw)
        assert_same_results(Dg, Ig, Do, Io)
    # exhaustive probe equals flat over the surviving rows
    Dg, Ig = ix.search(xq, 10, nprobe=24)
    Do, Io = oracle.flat_search(L2, cur, mids, xq, 10)
    assert_same_results(Dg, Ig, Do, Io)


def test_gpu_training_quality(oracle):
    rng = np.random.default_rng(11)
    n, d, nlist = 20000, 32, 64
    centers = rng.random((nlist, d)).astype(np.float32)
    xb = (centers[rng.integers(0, nlist, n)] + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    ids = np.arange(n, dtype=np.int64)
    ix = b200vs.Index(IVF_FLAT, L2, d, nlist=nlist)
    ix.train(xb)
    ix.add(xb, ids)
    state = ix.get_trained_state()
    cent_gpu = state[32:].view(np.float32).reshape(nlist, d)
    cent_cpu = oracle.kmeans(L2, xb, nlist)
    def objective(c):
        a = oracle.assign(L2, xb, c)
        return ((xb - c[a]) ** 2).sum(1).mean()
    og, oc = objective(cent_gpu), objective(cent_cpu)
    assert og < 1.15 * oc, (og, oc)
    # recall@10 of the GPU-trained index vs exact, nprobe 8
    xq = xb[:200] + 0.01 * rng.standard_normal((200, d)).astype(np.float32)
    Dg, Ig = ix.search(xq, 10, nprobe=8)
    Df, If = oracle.flat_search(L2, xb, ids, xq, 10, nthreads=8)
    assert recall(Ig, If) > 0.9


def test_range_search(oracle):
    ix, xb, stored, ids, cent = build_pair(oracle, L2, 3000, 16, 12, seed=9)
    xq = np.random.default_rng(2).random((3, 16)).astype(np.float32)
    off, lx, _, lids = ix.export_lists(12)
    Do, Io = oracle.ivfflat_search(L2, cent, off, lx, lids, xq, 3000, 12)
    radius = float(Do[:, 30].mean())
    D, I, C = ix.range_search(xq, radius, 128, nprobe=12)
    for q in range(3):
        want = Io[q][(Do[q] < radius) & (Io[q] >= 0)]
        assert C[q] == len(want) and list(I[q, :C[q]]) == list(want)
