"""Search is called concurrently from many host threads (the reference uses a 16-thread pool, src/server/server.cc:868-873)
while Raft apply keeps writing through the same index object.  Concurrent searches run on separate lanes of the library;
writers take the write lock and wait for in-flight device work."""
import threading

import numpy as np
import pytest

import b200vs
from b200vs import FLAT, IVF_FLAT, L2
from gpu_util import assert_same_results, require_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [FLAT, IVF_FLAT])
def test_concurrent_searches_and_far_away_writes(kind):
    require_gpu()
    rng = np.random.default_rng(0)
    n, d = 30000, 128
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(kind, L2, d, nlist=32)
    if kind == IVF_FLAT:
        ix.train(xb)
    ix.add(xb, ids)
    queries = [rng.random((nq, d)).astype(np.float32) for nq in (64, 8, 128, 33)]  # TC path and exact path mixed
    want = [ix.search(q, 10, nprobe=8) for q in queries]
    errors = []
    stop = threading.Event()

    def reader(t):
        try:
            for it in range(30):
                j = (t + it) % len(queries)
                D, I = ix.search(queries[j], 10, nprobe=8)
                assert_same_results(D, I, want[j][0], want[j][1])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
        finally:
            stop.set() if t == 0 else None

    def writer():
        far = (rng.random((256, d)) + 100.0).astype(np.float32)  # never among the neighbours of the queries
        base = 10_000_000
        i = 0
        try:
            while not stop.is_set() and i < 40:
                ix.upsert(far, np.arange(base, base + 256, dtype=np.int64))
                ix.delete(np.arange(base, base + 128, dtype=np.int64))
                i += 1
        except Exception as e:  # noqa: BLE001
            errors.append("writer " + repr(e))

    ths = [threading.Thread(target=reader, args=(t,)) for t in range(6)] + [threading.Thread(target=writer)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors[:3]
    D, I = ix.search(queries[0], 10, nprobe=8)
    assert_same_results(D, I, want[0][0], want[0][1])
