"""Step-by-step smoke of the tensor-core path with verbose output (run on the GPU box when test_gpu_tc fails)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dingo-store_b200", "python"))
import b200vs

def run(kind, metric, n, d, nq, k, nlist=0, nprobe=0):
    rng = np.random.default_rng(1)
    xb = rng.random((n, d)).astype(np.float32); xq = rng.random((nq, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(kind, metric, d, nlist=nlist)
    if kind == b200vs.IVF_FLAT: ix.train(xb)
    ix.add(xb, ids)
    print(f"[{kind} m{metric} n{n} d{d} nq{nq} k{k}] built", flush=True)
    t = time.time(); D2, I2 = ix.search(xq, k, nprobe=nprobe, exact_only=True); print("  exact ok %.3fs" % (time.time() - t), flush=True)
    ix.set_profiling(True)
    t = time.time(); D1, I1 = ix.search(xq, k, nprobe=nprobe); st = ix.stats(); print("  tc ok %.3fs stats %s" % (time.time() - t, st), flush=True)
    same = np.array_equal(I1, I2) and np.array_equal(D1.view(np.uint32), D2.view(np.uint32))
    print("  identical:", same, flush=True)
    if not same:
        bad = np.argwhere(I1 != I2)
        print("  mismatching rows:", len(set(bad[:, 0])), "of", nq)
        q = bad[0, 0]
        print("  q", q, "tc   ", I1[q], D1[q]); print("  q", q, "exact", I2[q], D2[q])
    return same

ok = True
ok &= run(b200vs.FLAT, b200vs.L2, 4096, 64, 16, 5)
ok &= run(b200vs.FLAT, b200vs.L2, 20000, 128, 64, 10)
ok &= run(b200vs.FLAT, b200vs.IP, 20000, 768, 100, 10)
ok &= run(b200vs.IVF_FLAT, b200vs.L2, 50000, 768, 256, 10, nlist=64, nprobe=8)
print("ALL OK" if ok else "MISMATCH", flush=True)
