import numpy as np
import pytest


def require_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def assert_same_results(Dg, Ig, Do, Io, exact=True, rtol=1e-4):
    """GPU (Dg, Ig) vs oracle (Do, Io).  exact: ids and distance bit patterns identical."""
    if exact:
        bad = np.argwhere(Ig != Io)
        assert bad.size == 0, f"id mismatch at {bad[:5].tolist()}: gpu {Ig[tuple(bad[0])]} oracle {Io[tuple(bad[0])]}"
        assert np.array_equal(Dg.view(np.uint32), Do.view(np.uint32)), "distance bit patterns differ"
    else:
        valid = Io >= 0
        assert np.allclose(Dg[valid], Do[valid], rtol=rtol, atol=1e-6)


def recall(Ia, Ib):
    k = Ia.shape[1]
    return float(np.mean([len(set(a[a >= 0]) & set(b[b >= 0])) / max(1, (b >= 0).sum()) for a, b in zip(Ia, Ib)]))
