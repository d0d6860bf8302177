"""ctypes access to the CPU ORACLE (oracle/_build/liboracle.so) and to the reference's own simd kernels
(oracle/_ref/libdingo_simd_ref.so).  TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libdingo_simd_ref.so")

L2, IP, COSINE = 1, 2, 3


class Filter(ctypes.Structure):
    _fields_ = [("has_range", ctypes.c_int), ("range_min", ctypes.c_int64), ("range_max", ctypes.c_int64),
                ("sorted_ids", ctypes.c_void_p), ("n_ids", ctypes.c_int64), ("negate", ctypes.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Oracle:
    def __init__(self, path):
        L = ctypes.CDLL(path)
        vp, i32, i64, f32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
        for n in ("oracle_fvec_L2sqr", "oracle_fvec_inner_product", "oracle_fvec_L2sqr_seq", "oracle_fvec_inner_product_seq"):
            f = getattr(L, n)
            f.restype = f32
            f.argtypes = [vp, vp, sz]
        L.oracle_normalize_faiss.argtypes = [vp, i32]
        L.oracle_normalize_hnsw.argtypes = [vp, i32, vp]
        L.oracle_fixture_mt19937.argtypes = [i64, i32, vp]
        fp = ctypes.POINTER(Filter)
        L.oracle_flat_search.argtypes = [ctypes.c_int, i32, i64, vp, vp, i64, vp, i32, fp, ctypes.c_int, vp, vp]
        L.oracle_kmeans.argtypes = [ctypes.c_int, i32, i64, vp, i32, i32, i32, i64, ctypes.c_int, vp]
        L.oracle_assign.argtypes = [ctypes.c_int, i32, i64, vp, i32, vp, ctypes.c_int, vp]
        L.oracle_ivfflat_search.argtypes = [ctypes.c_int, i32, i32, vp, vp, vp, vp, i64, vp, i32, i32, fp, ctypes.c_int, vp, vp]
        L.oracle_pq_train.argtypes = [i32, i32, i32, i64, vp, i32, i64, ctypes.c_int, vp]
        L.oracle_pq_encode.argtypes = [i32, i32, i32, vp, i64, vp, ctypes.c_int, vp]
        L.oracle_ivfpq_encode.argtypes = [i32, i32, i32, vp, i32, vp, i64, vp, vp, ctypes.c_int, vp]
        L.oracle_ivfpq_search.argtypes = [ctypes.c_int, i32, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp, i32, i32, fp, ctypes.c_int, vp, vp]
        L.oracle_calc_distance.argtypes = [ctypes.c_int, ctypes.c_int, i32, i64, vp, i64, vp, vp, vp, vp]
        L.oracle_hnsw_create.argtypes = [ctypes.c_int, i32, i64, i32, i32, i64]
        L.oracle_hnsw_create.restype = vp
        L.oracle_hnsw_destroy.argtypes = [vp]
        L.oracle_hnsw_add.argtypes = [vp, i64, vp, vp]
        L.oracle_hnsw_search.argtypes = [vp, i64, vp, i32, i32, fp, ctypes.c_int, vp, vp, vp, vp]
        L.oracle_hnsw_export_size.argtypes = [vp]
        L.oracle_hnsw_export_size.restype = i64
        L.oracle_hnsw_export.argtypes = [vp, vp, i64]
        L.oracle_hnsw_import.argtypes = [vp, vp, i64]
        L.oracle_version.restype = ctypes.c_char_p
        L.oracle_parallel_copy.argtypes = [vp, vp, sz, ctypes.c_int]
        self.L = L

    def numa_spread(self, a, nthreads):
        """Copy of `a` whose pages are first-touched by `nthreads` workers (spread over the NUMA nodes)."""
        a = np.ascontiguousarray(a)
        out = np.empty_like(a)
        self.L.oracle_parallel_copy(out.ctypes.data, a.ctypes.data, a.nbytes, nthreads)
        return out

    # ---- primitives ----
    def l2sqr(self, x, y):
        x, y = _f32(x), _f32(y)
        return self.L.oracle_fvec_L2sqr(x.ctypes.data, y.ctypes.data, x.size)

    def ip(self, x, y):
        x, y = _f32(x), _f32(y)
        return self.L.oracle_fvec_inner_product(x.ctypes.data, y.ctypes.data, x.size)

    def fixture(self, n, d):
        out = np.zeros((n, d), dtype=np.float32)
        self.L.oracle_fixture_mt19937(n, d, out.ctypes.data)
        return out

    def normalize_faiss(self, x):
        x = _f32(x).copy()
        for row in x.reshape(-1, x.shape[-1]):
            self.L.oracle_normalize_faiss(row.ctypes.data, row.size)
        return x

    def normalize_hnsw(self, x):
        x = _f32(x)
        out = np.zeros_like(x)
        xr, orr = x.reshape(-1, x.shape[-1]), out.reshape(-1, x.shape[-1])
        for i in range(xr.shape[0]):
            self.L.oracle_normalize_hnsw(xr[i].ctypes.data, xr.shape[1], orr[i].ctypes.data)
        return out

    @staticmethod
    def _filter(id_range=None, sorted_ids=None, negate=False):
        if id_range is None and sorted_ids is None:
            return None, None
        f = Filter()
        keep = None
        if id_range is not None:
            f.has_range, f.range_min, f.range_max = 1, int(id_range[0]), int(id_range[1])
        if sorted_ids is not None:
            keep = _i64(sorted_ids)
            f.sorted_ids, f.n_ids, f.negate = keep.ctypes.data, keep.size, int(bool(negate))
        return f, keep

    # ---- searches ----
    def calc_distance(self, algorithm, metric, left, right):
        """(distances [nl, nr], left_out, right_out) of VectorIndexUtils::CalcDistanceEntry."""
        left, right = _f32(left), _f32(right)
        nl, nr, d = left.shape[0], right.shape[0], left.shape[1]
        out = np.zeros((nl, nr), np.float32)
        lo, ro = np.zeros_like(left), np.zeros_like(right)
        rc = self.L.oracle_calc_distance(algorithm, metric, d, nl, left.ctypes.data, nr, right.ctypes.data, out.ctypes.data, lo.ctypes.data, ro.ctypes.data)
        assert rc == 0
        return out, lo, ro

    def flat_search(self, metric, xb, ids, xq, k, nthreads=1, **filt):
        xb, ids, xq = _f32(xb), _i64(ids), _f32(xq)
        n, d = xb.shape if xb.ndim == 2 else (0, xq.shape[1])
        nq = xq.shape[0]
        D = np.zeros((nq, k), np.float32)
        I = np.full((nq, k), -1, np.int64)
        f, keep = self._filter(**filt)
        rc = self.L.oracle_flat_search(metric, d, n, xb.ctypes.data, ids.ctypes.data, nq, xq.ctypes.data, k,
                                       ctypes.byref(f) if f else None, nthreads, D.ctypes.data, I.ctypes.data)
        assert rc == 0
        return D, I

    def kmeans(self, metric, x, k, niter=10, max_pts=256, seed=1234, nthreads=8):
        x = _f32(x)
        c = np.zeros((k, x.shape[1]), np.float32)
        rc = self.L.oracle_kmeans(metric, x.shape[1], x.shape[0], x.ctypes.data, k, niter, max_pts, seed, nthreads, c.ctypes.data)
        assert rc == 0, rc
        return c

    def assign(self, metric, x, centroids, nthreads=8):
        x, c = _f32(x), _f32(centroids)
        out = np.zeros(x.shape[0], np.int32)
        self.L.oracle_assign(metric, x.shape[1], x.shape[0], x.ctypes.data, c.shape[0], c.ctypes.data, nthreads, out.ctypes.data)
        return out

    def ivfflat_search(self, metric, centroids, list_off, xb, ids, xq, k, nprobe, nthreads=1, **filt):
        c, off, xb, ids, xq = _f32(centroids), _i64(list_off), _f32(xb), _i64(ids), _f32(xq)
        nq, d = xq.shape
        D = np.zeros((nq, k), np.float32)
        I = np.full((nq, k), -1, np.int64)
        f, keep = self._filter(**filt)
        rc = self.L.oracle_ivfflat_search(metric, d, c.shape[0], c.ctypes.data, off.ctypes.data, xb.ctypes.data, ids.ctypes.data,
                                          nq, xq.ctypes.data, k, nprobe, ctypes.byref(f) if f else None, nthreads,
                                          D.ctypes.data, I.ctypes.data)
        assert rc == 0
        return D, I

    def pq_train(self, x, M, nbits=8, niter=25, seed=1234, nthreads=8):
        x = _f32(x)
        d = x.shape[1]
        cb = np.zeros((M, 1 << nbits, d // M), np.float32)
        rc = self.L.oracle_pq_train(d, M, nbits, x.shape[0], x.ctypes.data, niter, seed, nthreads, cb.ctypes.data)
        assert rc == 0, rc
        return cb

    def ivfpq_encode(self, codebooks, centroids, x, assign, nthreads=8):
        cb, c, x = _f32(codebooks), _f32(centroids), _f32(x)
        a = np.ascontiguousarray(assign, dtype=np.int32)
        M = cb.shape[0]
        codes = np.zeros((x.shape[0], M), np.uint8)
        rc = self.L.oracle_ivfpq_encode(x.shape[1], M, 8, cb.ctypes.data, c.shape[0], c.ctypes.data, x.shape[0], x.ctypes.data,
                                        a.ctypes.data, nthreads, codes.ctypes.data)
        assert rc == 0
        return codes

    def ivfpq_search(self, metric, centroids, codebooks, list_off, codes, ids, xq, k, nprobe, nthreads=1, **filt):
        c, cb, off, ids, xq = _f32(centroids), _f32(codebooks), _i64(list_off), _i64(ids), _f32(xq)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        nq, d = xq.shape
        D = np.zeros((nq, k), np.float32)
        I = np.full((nq, k), -1, np.int64)
        f, keep = self._filter(**filt)
        rc = self.L.oracle_ivfpq_search(metric, d, c.shape[0], cb.shape[0], 8, c.ctypes.data, cb.ctypes.data, off.ctypes.data,
                                        codes.ctypes.data, ids.ctypes.data, nq, xq.ctypes.data, k, nprobe,
                                        ctypes.byref(f) if f else None, nthreads, D.ctypes.data, I.ctypes.data)
        assert rc == 0
        return D, I


class OracleHnsw:
    def __init__(self, o, metric, d, max_elements, M, efc, seed=100):
        self.o, self.d = o, d
        self.h = o.L.oracle_hnsw_create(metric, d, max_elements, M, efc, seed)

    def add(self, x, labels):
        x, labels = _f32(x), _i64(labels)
        rc = self.o.L.oracle_hnsw_add(self.h, x.shape[0], x.ctypes.data, labels.ctypes.data)
        assert rc == 0, rc

    def search(self, xq, k, ef=0, nthreads=1, **filt):
        xq = _f32(xq)
        nq = xq.shape[0]
        D = np.zeros((nq, k), np.float32)
        I = np.full((nq, k), -1, np.int64)
        nd = np.zeros(nq, np.int64)
        nh = np.zeros(nq, np.int64)
        f, keep = Oracle._filter(**filt)
        rc = self.o.L.oracle_hnsw_search(self.h, nq, xq.ctypes.data, k, ef, ctypes.byref(f) if f else None, nthreads,
                                         D.ctypes.data, I.ctypes.data, nd.ctypes.data, nh.ctypes.data)
        assert rc == 0
        return D, I, nd, nh

    def load(self, blob):
        """Adopt a graph in the export layout (e.g. the engine's b200vs_get_trained_state): searches then run on that graph."""
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        rc = self.o.L.oracle_hnsw_import(self.h, b.ctypes.data, b.nbytes)
        assert rc == 0, rc

    def export(self):
        n = self.o.L.oracle_hnsw_export_size(self.h)
        buf = np.zeros(n, np.uint8)
        rc = self.o.L.oracle_hnsw_export(self.h, buf.ctypes.data, n)
        assert rc == 0
        return buf

    def close(self):
        if self.h:
            self.o.L.oracle_hnsw_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _oracle = Oracle(ORACLE_SO)
    return _oracle


def load_ref():
    if not os.path.exists(REF_SO):
        return None
    L = ctypes.CDLL(REF_SO)
    for n in ("ref_fvec_L2sqr", "ref_fvec_inner_product", "ref_fvec_L2sqr_avx512", "ref_fvec_inner_product_avx512",
              "ref_fvec_L2sqr_avx", "ref_fvec_inner_product_avx", "ref_fvec_L2sqr_sse", "ref_fvec_inner_product_sse",
              "ref_fvec_L2sqr_ref", "ref_fvec_inner_product_ref"):
        f = getattr(L, n)
        f.restype = ctypes.c_float
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.ref_simd_type.restype = ctypes.c_char_p
    return L
