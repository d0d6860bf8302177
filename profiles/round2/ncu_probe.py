"""Small invocations of the kernels whose `ncu --set full` captures are committed under profiles/round2 (run under ncu with -k)."""
import sys, numpy as np, torch
sys.path.insert(0, "dingo-store_b200/python"); sys.path.insert(0, "tests")
import b200vs, oracle_lib
which = sys.argv[1]
o = oracle_lib.load()
rng = np.random.default_rng(0)
if which == "pq":
    n, d, M, nlist, nq, k = 1_000_000, 768, 96, 2048, 1024, 100
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    ix = b200vs.Index(b200vs.IVF_PQ, b200vs.IP, d, nlist=nlist, pq_m=M, pq_nbits=8)
    ix.train(torch.randn((256 * nlist, d), generator=g, device="cuda").cpu().numpy())
    for a in range(0, n, 131072):
        m = min(131072, n - a)
        ix.add(torch.randn((m, d), generator=g, device="cuda").cpu().numpy(), np.arange(a + 1, a + 1 + m, dtype=np.int64))
    xq = torch.randn((nq, d), generator=g, device="cuda").cpu().numpy()
    ix.search(xq, k, nprobe=64); ix.search(xq, k, nprobe=64)
elif which == "hnsw":
    n, d = 100_000, 768
    xb = rng.random((n, d)).astype(np.float32)
    ix = b200vs.Index(b200vs.HNSW, b200vs.COSINE, d, hnsw_m=16, hnsw_efc=200, max_elements=n, hnsw_build_threads=64)
    ix.add(xb, np.arange(1, n + 1, dtype=np.int64))
    xq = rng.random((512, d)).astype(np.float32)
    ix.search(xq, 10, efsearch=128); ix.search(xq, 10, efsearch=128)
elif which == "flat1":
    n, d = 100_000, 128
    ix = b200vs.Index(b200vs.FLAT, b200vs.L2, d)
    ix.add(rng.random((n, d)).astype(np.float32), np.arange(1, n + 1, dtype=np.int64))
    xq = rng.random((1, d)).astype(np.float32)
    for _ in range(3): ix.search(xq, 10)
torch.cuda.synchronize()
