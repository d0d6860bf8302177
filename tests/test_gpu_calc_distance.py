"""b200vs_calc_distance == the oracle's restatement of VectorIndexUtils::CalcDistanceEntry
(src/vector/vector_index_utils.cc:48-124, :193-419), bit for bit, for both algorithm flavours and all three metrics."""
import numpy as np
import pytest

import b200vs
import oracle_lib
from gpu_util import require_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algorithm", [b200vs.ALGORITHM_FAISS, b200vs.ALGORITHM_HNSWLIB])
@pytest.mark.parametrize("metric", [b200vs.L2, b200vs.IP, b200vs.COSINE])
@pytest.mark.parametrize("d,nl,nr", [(8, 3, 5), (1, 2, 2), (100, 7, 33), (768, 40, 65), (1027, 4, 9)])
def test_calc_distance_bit_exact(algorithm, metric, d, nl, nr):
    require_gpu()
    o = oracle_lib.load()
    left = o.fixture(nl, d)  # the reference's own fixture generator (mt19937, default seed)
    right = np.random.default_rng(d * 7 + nr).random((nr, d), dtype=np.float32) * 3 - 1
    right[0] = left[0]  # a zero distance / unit cosine
    want, wl, wr = o.calc_distance(algorithm, {b200vs.L2: oracle_lib.L2, b200vs.IP: oracle_lib.IP, b200vs.COSINE: oracle_lib.COSINE}[metric], left, right)
    got, gl, gr = b200vs.calc_distance(algorithm, metric, left, right, return_normalized=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(gl.view(np.uint32), wl.view(np.uint32))
    assert np.array_equal(gr.view(np.uint32), wr.view(np.uint32))
    if metric != b200vs.COSINE:  # is_return_normlize hands the operands back unchanged
        assert np.array_equal(gl, left) and np.array_equal(gr, right)


def test_calc_distance_known_values():
    """Hand-checkable values: (1,2,3) vs (4,6,8): L2 = 9+16+25 = 50, ip = 4+12+24 = 40 -> 1-ip = -39."""
    require_gpu()
    a = np.array([[1, 2, 3]], np.float32)
    b = np.array([[4, 6, 8], [1, 2, 3]], np.float32)
    assert b200vs.calc_distance(b200vs.ALGORITHM_FAISS, b200vs.L2, a, b).tolist() == [[50.0, 0.0]]
    assert b200vs.calc_distance(b200vs.ALGORITHM_HNSWLIB, b200vs.IP, a, b).tolist() == [[-39.0, -13.0]]
    c = b200vs.calc_distance(b200vs.ALGORITHM_FAISS, b200vs.COSINE, a, b)
    assert abs(c[0, 1]) < 1e-6 and abs(c[0, 0] - (1 - 40 / np.sqrt(14 * 116))) < 1e-6


def test_calc_distance_errors_and_empty():
    require_gpu()
    a = np.zeros((2, 4), np.float32)
    with pytest.raises(b200vs.B200VSError):
        b200vs.calc_distance(0, b200vs.L2, a, a)  # ALGORITHM_NONE
    with pytest.raises(b200vs.B200VSError):
        b200vs.calc_distance(b200vs.ALGORITHM_FAISS, 0, a, a)  # METRIC_TYPE_NONE
    out = b200vs.calc_distance(b200vs.ALGORITHM_FAISS, b200vs.L2, a, np.zeros((0, 4), np.float32))
    assert out.shape == (2, 0)
