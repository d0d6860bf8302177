"""Regenerates tests/golden/oracle_regression.json: outputs of the CPU ORACLE's search restatements (flat, k-means,
IVF-Flat, PQ, IVF-PQ, HNSW, the distance matrix) on small seeded inputs.

The reference's own tests hold no numeric goldens for these paths (faiss and hnswlib are not vendored),
so these values do NOT pin the oracle to the reference — tests/golden/simd_kat.json does that for the arithmetic.  They
freeze the restated algorithms: GPU parity is defined against the oracle, so an accidental change of the oracle
(tie rule, k-means subsampling, HNSW level draw ...) must show up as a diff here.
    python tests/golden/make_oracle_regression.py          # rewrites the file
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib  # noqa: E402
from oracle_lib import COSINE, IP, L2  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_regression.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel().tolist()


def compute():
    o = oracle_lib.load()
    out = {}
    xb = o.fixture(300, 16)  # the reference's fixture generator
    ids = np.arange(1, 301, dtype=np.int64) * 5
    xq = np.random.default_rng(7).random((4, 16), dtype=np.float32)
    for name, m in (("l2", L2), ("ip", IP), ("cosine", COSINE)):
        stored = o.normalize_faiss(xb) if m == COSINE else xb
        D, I = o.flat_search(m, stored, ids, xq, 6)
        out[f"flat_{name}"] = {"ids": I.ravel().tolist(), "dist_bits": bits(D)}
    D, I = o.flat_search(L2, xb, ids, xq, 6, id_range=(100, 900), sorted_ids=ids[::3], negate=True)
    out["flat_l2_filtered"] = {"ids": I.ravel().tolist(), "dist_bits": bits(D)}
    # k-means + IVF-Flat
    xt = np.random.default_rng(11).random((2000, 8), dtype=np.float32)
    for name, m in (("l2", L2), ("ip", IP)):
        cent = o.kmeans(m, xt, 16, niter=10, max_pts=64, seed=1234, nthreads=1)
        asg = o.assign(m, xt, cent, nthreads=1)
        order = np.argsort(asg, kind="stable")
        off = np.zeros(17, np.int64)
        off[1:] = np.cumsum(np.bincount(asg, minlength=16))
        q = np.random.default_rng(12).random((5, 8), dtype=np.float32)
        D, I = o.ivfflat_search(m, cent, off, xt[order], (order + 1).astype(np.int64), q, 5, 4)
        out[f"ivfflat_{name}"] = {"centroids_sha": sha(cent), "assign_sha": sha(asg), "ids": I.ravel().tolist(), "dist_bits": bits(D)}
    # PQ + IVF-PQ (by residual)
    xp = np.random.default_rng(13).random((3000, 16), dtype=np.float32)
    cent = o.kmeans(L2, xp, 8, niter=10, max_pts=256, seed=1234, nthreads=1)
    asg = o.assign(L2, xp, cent, nthreads=1)
    resid = xp - cent[asg]
    cb = o.pq_train(resid, 4, nbits=8, niter=5, seed=1234, nthreads=1)
    codes = o.ivfpq_encode(cb, cent, xp, asg, nthreads=1)
    order = np.argsort(asg, kind="stable")
    off = np.zeros(9, np.int64)
    off[1:] = np.cumsum(np.bincount(asg, minlength=8))
    q = np.random.default_rng(14).random((4, 16), dtype=np.float32)
    for name, m in (("l2", L2), ("ip", IP)):
        D, I = o.ivfpq_search(m, cent, cb, off, codes[order], (order + 1).astype(np.int64), q, 8, 3)
        out[f"ivfpq_{name}"] = {"codebooks_sha": sha(cb), "codes_sha": sha(codes), "ids": I.ravel().tolist(), "dist_bits": bits(D)}
    # HNSW
    xh = np.random.default_rng(15).random((400, 12), dtype=np.float32)
    for name, m in (("l2", L2), ("cosine", COSINE)):
        h = oracle_lib.OracleHnsw(o, m, 12, 1000, 8, 40)
        h.add(xh, np.arange(1, 401, dtype=np.int64))
        q = np.random.default_rng(16).random((5, 12), dtype=np.float32)
        D, I, nd, nh = h.search(q, 6, ef=32)
        out[f"hnsw_{name}"] = {"graph_sha": sha(h.export()), "ids": I.ravel().tolist(), "dist_bits": bits(D), "ndis": nd.tolist(), "hops": nh.tolist()}
        h.close()
    # distance matrix
    left, right = o.fixture(3, 10), np.random.default_rng(17).random((4, 10), dtype=np.float32)
    for alg in (1, 2):
        for name, m in (("l2", L2), ("ip", IP), ("cosine", COSINE)):
            d, lo, ro = o.calc_distance(alg, m, left, right)
            out[f"calc_{alg}_{name}"] = {"dist_bits": bits(d), "left_sha": sha(lo), "right_sha": sha(ro)}
    return out


if __name__ == "__main__":
    res = compute()
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", len(res), "cases to", OUT)
