#!/usr/bin/env python
"""bench_configs.py — the OTHER BASELINE.json configs (bench.py measures the headline one), one JSON line each.

  cfg1  Flat L2, 100K x 128, top-10, batch = 1            (the reference's own CPU-runnable case)
  cfg2  IVF-Flat L2, 1M x 768, nlist 1024, nprobe 32, batch 256, top-10
  cfg3  IVF-PQ IP, 10M x 768, M = 96, nbits 8, nlist 2048, nprobe 64, batch 1024, top-100   (--pq-n)
  cfg4  HNSW cosine, 1M x 768, M = 16, efConstruction 200, efSearch 128, batch 512          (--hnsw-n; concurrent host build)
  cfg5  IVF-Flat L2, (rows-per-GPU x world) x 1536, nlist 2048 per GPU, nprobe 64, batch 4096, top-10, list-sharded over
        the GPUs of the box through b200vs_shard_* — run under torchrun (BASELINE: 12.5 M rows per GPU x 8 GPUs = 100 M)
Each line: device-resident QPS (CUDA events), e2e QPS through the host-pointer C ABI, parity with the oracle on a
bounded sample, the dominant kernel's time from the library's profiling mode and the algorithmic figure of
SURVEY.md §8(d).  Scaled sizes are stated in the line ("scaled_from").  Not part of the driver contract.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dingo-store_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import b200vs  # noqa: E402
import oracle_lib  # noqa: E402
import b200vs.shard as shard_host  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_search(ix, xq, k, steps=20, warmup=5, **kw):
    dev = torch.device("cuda", 0)
    nq = xq.shape[0]
    q = torch.from_numpy(xq).to(dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sp, keep = b200vs.make_search_params(**kw)
    st = torch.cuda.Stream(device=dev)
    for _ in range(warmup):
        ix.search_device(nq, q.data_ptr(), k, od.data_ptr(), oi.data_ptr(), stream=st.cuda_stream, sp=sp)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        ix.search_device(nq, q.data_ptr(), k, od.data_ptr(), oi.data_ptr(), stream=st.cuda_stream, sp=sp)
    e1.record(st)
    st.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    qh = torch.from_numpy(xq).pin_memory()
    hd = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    hi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    for _ in range(warmup):
        ix.search_raw(nq, qh.data_ptr(), k, hd.data_ptr(), hi.data_ptr(), sp=sp)
    t = time.perf_counter()
    for _ in range(steps):
        ix.search_raw(nq, qh.data_ptr(), k, hd.data_ptr(), hi.data_ptr(), sp=sp)
    e2e_ms = (time.perf_counter() - t) / steps * 1e3
    ix.set_profiling(True)
    ix.search_device(nq, q.data_ptr(), k, od.data_ptr(), oi.data_ptr(), stream=st.cuda_stream, sp=sp)
    st.synchronize()
    stats = ix.stats()
    ix.set_profiling(False)
    return dev_ms, e2e_ms, stats, hd.numpy().copy(), hi.numpy().copy()


def rnd(n, d, seed, dist="uniform"):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    out = np.empty((n, d), np.float32)
    for a in range(0, n, 262144):
        m = min(262144, n - a)
        x = torch.rand((m, d), generator=g, device="cuda") if dist == "uniform" else torch.randn((m, d), generator=g, device="cuda")
        out[a:a + m] = x.cpu().numpy()
    return out


def cfg1(o, cores):
    n, d, k = 100_000, 128, 10
    xb = rnd(n, d, 1234)
    ids = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(b200vs.FLAT, b200vs.L2, d)
    for a in range(0, n, 32768):
        ix.add(xb[a:a + 32768], ids[a:a + 32768])
    xq = rnd(1, d, 4321)
    dev_ms, e2e_ms, st, D, I = time_search(ix, xq, k, steps=200, warmup=20)
    t = time.perf_counter()
    reps = 20
    for _ in range(reps):
        Do, Io = o.flat_search(oracle_lib.L2, xb, ids, xq, k, nthreads=1)
    cpu_ms = (time.perf_counter() - t) / reps * 1e3
    bytes_q = n * d * 4
    return {"config": "cfg1 Flat L2 100Kx128 top-10 batch=1", "qps_device": 1e3 / dev_ms, "latency_ms_device": dev_ms, "qps_e2e": 1e3 / e2e_ms,
            "ids_bit_exact": bool(np.array_equal(I, Io)), "dist_bit_exact": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32))),
            "algorithmic_bytes_per_query": bytes_q, "achieved_gbs_whole_call": bytes_q / (dev_ms * 1e-3) / 1e9, "hbm_peak_gbs": peak(),
            "cpu_baseline": {"qps": 1e3 / cpu_ms, "cores": 1, "kind": "port", "sample": "the same single query, one thread (the reference runs one query per task)"},
            "path": "exact FP32 scan (batch < 16)", "search_stats": st}


def cfg2(o, cores):
    n, d, nlist, nprobe, nq, k = 1_000_000, 768, 1024, 32, 256, 10
    xb = rnd(n, d, 1234)
    ids = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(b200vs.IVF_FLAT, b200vs.L2, d, nlist=nlist)
    ix.train(xb[:nlist * 256])
    for a in range(0, n, 32768):
        ix.add(xb[a:a + 32768], ids[a:a + 32768])
    xq = rnd(nq, d, 4321)
    dev_ms, e2e_ms, st, D, I = time_search(ix, xq, k, nprobe=nprobe)
    off, lx, _, lids = ix.export_lists(nlist)
    lx = o.numa_spread(lx, cores)
    cent = ix.get_trained_state()[32:].view(np.float32).reshape(nlist, d)
    t = time.perf_counter()
    Do, Io = o.ivfflat_search(oracle_lib.L2, cent, off, lx, lids, xq, k, nprobe, nthreads=cores)
    cpu_s = time.perf_counter() - t
    rows = st[4]
    return {"config": "cfg2 IVF-Flat L2 1Mx768 nlist=1024 nprobe=32 batch=256 top-10", "qps_device": nq / dev_ms * 1e3, "ms_per_batch_device": dev_ms,
            "qps_e2e": nq / e2e_ms * 1e3, "ids_bit_exact": bool(np.array_equal(I, Io)), "dist_bit_exact": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32))),
            "recall_at_10_vs_oracle": float(np.mean([len(set(a) & set(b)) / k for a, b in zip(I, Io)])),
            "roofline": {"bound": "hbm", "kernel": "tc_scan_kernel capture pass", "kernel_ms": st[3] / 1e6, "algorithmic_bytes": rows * (d * 4 + 8),
                         "achieved": rows * (d * 4 + 8) / max(st[3], 1) , "unit": "GB/s", "peak": peak(), "frac": rows * (d * 4 + 8) / max(st[3], 1) / peak()},
            "cpu_baseline": {"qps": nq / cpu_s, "cores": cores, "kind": "port", "sample": f"all {nq} queries, {cores} threads, one query per task"},
            "fallback_queries": st[2], "search_stats": st}


def cfg3(o, cores, n):
    d, M, nlist, nprobe, nq, k = 768, 96, 2048, 64, 1024, 100
    ix = b200vs.Index(b200vs.IVF_PQ, b200vs.IP, d, nlist=nlist, pq_m=M, pq_nbits=8)
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    t0 = time.time()
    ntrain = max(256 * nlist, 65536)
    train = torch.randn((ntrain, d), generator=g, device="cuda").cpu().numpy()
    ix.train(train)
    t_train = time.time() - t0
    keep = None
    t0 = time.time()
    nid = 1
    for a in range(0, n, 131072):
        m = min(131072, n - a)
        x = torch.randn((m, d), generator=g, device="cuda").cpu().numpy()
        if keep is None:
            keep = x[:20000].copy()
        for b in range(0, m, 32768):
            mm = min(32768, m - b)
            ix.add(x[b:b + mm], np.arange(nid, nid + mm, dtype=np.int64))
            nid += mm
    t_add = time.time() - t0
    xq = rnd(nq, d, 4321, "normal")
    dev_ms, e2e_ms, st, D, I = time_search(ix, xq, k, steps=5, warmup=2, nprobe=nprobe)
    # parity on a sample: oracle LUT scan over the exported codes with the GPU-trained state
    blob = ix.get_trained_state()
    cent = blob[48:48 + nlist * d * 4].view(np.float32).reshape(nlist, d)
    cb = blob[48 + nlist * d * 4:].view(np.float32).reshape(M, 256, d // M)
    off, _, codes, lids = ix.export_lists(nlist, with_vectors=False, code_size=M)
    ns = 32
    t = time.perf_counter()
    Do, Io = o.ivfpq_search(oracle_lib.IP, cent, cb, off, codes, lids, xq[:ns], k, nprobe, nthreads=cores)
    cpu_s = time.perf_counter() - t
    lookups = float(st[4]) * M * 0 + 0  # per-batch lookups need the per-query probe sizes: use avg list len
    avg_len = n / nlist
    lookups = nq * nprobe * avg_len * M
    return {"config": f"cfg3 IVF-PQ IP {n}x768 M=96 nbits=8 nlist=2048 nprobe=64 batch=1024 top-100", "scaled_from": "10M x 768" if n < 10_000_000 else None,
            "qps_device": nq / dev_ms * 1e3, "ms_per_batch_device": dev_ms, "qps_e2e": nq / e2e_ms * 1e3,
            "ids_bit_exact_sample": bool(np.array_equal(I[:ns], Io)), "recall_at_100_vs_oracle": float(np.mean([len(set(a) & set(b)) / k for a, b in zip(I[:ns], Io)])),
            "max_rel_dist_err": float(np.max(np.abs(D[:ns] - Do) / np.maximum(np.abs(Do), 1e-6))),
            "roofline": {"bound": "shared-memory LUT", "kernel": "pq_scan_select_kernel", "kernel_ms": st[3] / 1e6, "lookups_per_batch": lookups,
                         "achieved_lookups_per_s": lookups / max(st[3] * 1e-9, 1e-9), "bytes_scan_equiv_gbs": lookups / max(st[3], 1), "hbm_peak_gbs": peak()},
            "cpu_baseline": {"qps": ns / cpu_s, "cores": cores, "kind": "port", "sample": f"first {ns} queries, {cores} threads"},
            "train_seconds": t_train, "add_seconds": t_add, "search_stats": st}


def cfg4(o, cores, n, build_threads=64):
    """HNSW cosine.  The graph is built by the engine with `build_threads` concurrent host writers — what the reference does
    with its thread pool (vector_index_hnsw.cc:229-243; such a graph is not reproducible run to run) — and the oracle adopts
    that very graph (oracle_hnsw_import), so both sides search the SAME index, as for the IVF types."""
    d, M, efc, ef, nq, k = 768, 16, 200, 128, 512, 10
    xb = rnd(n, d, 1234)
    labels = np.arange(1, n + 1, dtype=np.int64)
    ix = b200vs.Index(b200vs.HNSW, b200vs.COSINE, d, hnsw_m=M, hnsw_efc=efc, max_elements=n * 2, hnsw_build_threads=build_threads)
    t0 = time.time()
    for a in range(0, n, 65536):
        ix.add(xb[a:a + 65536], labels[a:a + 65536])
    t_build = time.time() - t0
    xq = rnd(nq, d, 4321)
    dev_ms, e2e_ms, st, D, I = time_search(ix, xq, k, steps=5, warmup=2, efsearch=ef)
    h = oracle_lib.OracleHnsw(o, oracle_lib.COSINE, d, n, M, efc)
    t0 = time.time()
    h.load(ix.get_trained_state())
    t_import = time.time() - t0
    t = time.perf_counter()
    Do, Io, nd, nh = h.search(xq, k, ef=ef, nthreads=cores)
    cpu_s = time.perf_counter() - t
    t = time.perf_counter()
    h.search(xq[:64], k, ef=ef, nthreads=min(16, cores))
    cpu16_s = time.perf_counter() - t
    # graph quality: recall of the search against the exact answer on a sample
    ns = 64
    xn = xb / (np.linalg.norm(xb, axis=1, keepdims=True) + 1e-30)
    qn = xq[:ns] / (np.linalg.norm(xq[:ns], axis=1, keepdims=True) + 1e-30)
    exact = np.argsort(-(qn @ xn.T), axis=1)[:, :k] + 1
    rec = float(np.mean([len(set(a) & set(b)) / k for a, b in zip(I[:ns], exact)]))
    bytes_batch = float(nd.sum()) * d * 4 + float(nh.sum()) * (4 + 2 * M * 4)
    return {"config": f"cfg4 HNSW cosine {n}x768 M=16 efC=200 efSearch=128 batch=512 top-10", "scaled_from": None if n >= 1_000_000 else "1M x 768",
            "qps_device": nq / dev_ms * 1e3, "ms_per_batch_device": dev_ms, "qps_e2e": nq / e2e_ms * 1e3,
            "graph": f"built by the engine with {build_threads} concurrent writers (reference: 16-thread pool); the oracle searches the same graph",
            "ids_bit_exact": bool(np.array_equal(I, Io)), "dist_bit_exact": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32))),
            "recall_at_10_vs_exact_cosine": rec,
            "roofline": {"bound": "hbm random gathers", "kernel": "hnsw_search_kernel", "kernel_ms": st[3] / 1e6, "algorithmic_bytes": bytes_batch,
                         "ndis_per_query": float(nd.mean()), "hops_per_query": float(nh.mean()), "achieved": bytes_batch / max(st[3], 1), "unit": "GB/s",
                         "peak": peak(), "frac": bytes_batch / max(st[3], 1) / peak()},
            "cpu_baseline": {"qps": nq / cpu_s, "cores": cores, "kind": "port", "sample": f"all {nq} queries, {cores} threads",
                             "reference_shape_16_threads_qps": 64 / cpu16_s},
            "build_seconds_engine": t_build, "oracle_import_seconds": t_import, "search_stats": st}

def cfg5(o, cores, rows_per_gpu, steps=10, warmup=3, dim=1536, nlist_per_gpu=2048, nprobe=64, nq=4096, k=10, in_flight=2, verify=256, oracle_q=8):
    """BASELINE config 5: one logical IVF-Flat index list-sharded over the world's GPUs behind b200vs_shard_*.
    Build = distributed training, a planning pass that pre-sizes every owned list, then the add pass (rows generated on the
    device, routed to their list owner over NCCL).  Every rank prints nothing; rank 0 returns the JSON line."""
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    d, n = dim, rows_per_gpu
    nlist = nlist_per_gpu * world
    t0 = time.time()
    ix = b200vs.Index(b200vs.IVF_FLAT, b200vs.L2, d, nlist=nlist, device=local)
    idb = None
    if world > 1:
        t = torch.from_numpy(b200vs.Shard.unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(t, 0)
        idb = t.cpu().numpy()
    sh = b200vs.Shard(ix, rank, world, idb, lanes=in_flight)
    chunk = 65536

    def chunks():
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        for a in range(0, n, chunk):
            m = min(chunk, n - a)
            yield a, torch.rand((m, d), generator=g, device=dev, dtype=torch.float32)

    g0 = torch.Generator(device=dev)
    g0.manual_seed(99 + rank)
    ntrain = min(n, nlist_per_gpu * 64)
    sh.train(torch.rand((ntrain, d), generator=g0, device=dev, dtype=torch.float32).cpu().numpy())
    t_train = time.time() - t0
    t1 = time.time()
    for a, x in chunks():  # pass 1: assignment only -> per-list row counts
        torch.cuda.synchronize()
        sh.plan_add_device(x.shape[0], x.data_ptr())
    sh.plan_commit()       # all-reduce of the counts, ONE arena allocation per rank
    t_plan = time.time() - t1
    t2 = time.time()
    for a, x in chunks():  # pass 2: the same rows again, now stored on their list owner
        gid = torch.arange(a + 1, a + 1 + x.shape[0], dtype=torch.int64, device=dev) + rank * n
        torch.cuda.synchronize()
        sh.add_device(x.shape[0], x.data_ptr(), gid.data_ptr())
    t_add = time.time() - t2
    del x
    torch.cuda.empty_cache()
    mem = ix.get_memory_size()
    cnt = torch.tensor([ix.get_count()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    # ---- timed search: device-resident queries, `in_flight` batches in flight ----
    gq = torch.Generator(device=dev)
    gq.manual_seed(4321)
    qs = [torch.rand((nq, d), generator=gq, device=dev, dtype=torch.float32) for _ in range(2)]
    L = in_flight
    od = [torch.empty((nq, k), dtype=torch.float32, device=dev) for _ in range(L)]
    oi = [torch.empty((nq, k), dtype=torch.int64, device=dev) for _ in range(L)]
    sp, _keep = b200vs.make_search_params(nprobe=nprobe)
    main = torch.cuda.Stream(device=dev)
    sts = [torch.cuda.Stream(device=dev) for _ in range(L)]
    torch.cuda.set_stream(main)
    seq = [0]

    host_ms = []

    def step(i, lanes):
        ln = i % lanes
        th = time.perf_counter()
        sh.search_device(nq, qs[i % 2].data_ptr(), k, od[ln].data_ptr(), oi[ln].data_ptr(), stream=sts[ln].cuda_stream, sp=sp, seq=seq[0])
        host_ms.append((time.perf_counter() - th) * 1e3)
        seq[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, lanes):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in sts[:lanes]:
            st.wait_event(e0)
        for i in range(nsteps):
            step(i, lanes)
        for st in sts[:lanes]:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    for i in range(max(warmup, 2 * L)):
        step(i, L)
    barrier()
    ms = timed(steps, L)
    del host_ms[:]
    ms1 = timed(steps, 1)
    host_enqueue_ms = float(np.median(host_ms)) if host_ms else None
    # e2e through the host-pointer collective call
    qh = [q.cpu().pin_memory() for q in qs]
    hd = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    hi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    for i in range(2):
        sh.search_raw(nq, qh[i % 2].data_ptr(), k, hd.data_ptr(), hi.data_ptr(), sp=sp, seq=seq[0]); seq[0] += 1
    barrier()
    tt = time.perf_counter()
    for i in range(steps):
        sh.search_raw(nq, qh[i % 2].data_ptr(), k, hd.data_ptr(), hi.data_ptr(), sp=sp, seq=seq[0]); seq[0] += 1
    barrier()
    e2e_ms = (time.perf_counter() - tt) * 1e3
    tm = torch.tensor([ms, ms1, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, ms1, e2e_ms = tm.tolist()
    # roofline of the list scan on this rank (profiling pass, never part of the timed numbers)
    ix.set_profiling(True)
    sh.search_device(nq, qs[0].data_ptr(), k, od[0].data_ptr(), oi[0].data_ptr(), stream=sts[0].cuda_stream, sp=sp, seq=seq[0]); seq[0] += 1
    torch.cuda.synchronize()
    st = ix.stats()
    ph = ix.phase_times()
    ix.set_profiling(False)
    alg = st[4] * (d * 4 + 8)
    roof = torch.tensor([alg / max(st[3], 1), st[3] / 1e6, float(st[2])], dtype=torch.float64, device=dev)
    roofs = [torch.zeros_like(roof) for _ in range(world)]
    phs = [None] * world
    if world > 1:
        dist.all_gather(roofs, roof)
        dist.all_gather_object(phs, {a: round(v, 3) for a, v in ph.items()})
    else:
        roofs = [roof]
        phs = [ph]
    # ---- verification 1: tile path vs the exact FP32 scan path (different kernels) on `verify` queries, full scale ----
    xq = qh[0].numpy()[:verify].copy()
    Dt, It = sh.search(xq, k, seq=seq[0], nprobe=nprobe); seq[0] += 1
    De, Ie = sh.search(xq, k, seq=seq[0], nprobe=nprobe, exact_only=True); seq[0] += 1
    # ---- verification 2: the CPU oracle on the lists the first `oracle_q` queries probe (exported list by list) ----
    cent = ix.get_trained_state()[32:].view(np.float32).reshape(nlist, d)
    xo = xq[:oracle_q]
    # lists worth exporting: the exact top-(nprobe + 8) centroids of each query (a few extra: rounding at the boundary is harmless)
    dd = ((xo[:, None, :].astype(np.float64) - cent[None, :, :].astype(np.float64)) ** 2).sum(-1)
    pls = set(np.unique(np.argsort(dd, axis=1, kind="stable")[:, :nprobe + 8]).tolist())
    b, e = sh.list_range()
    off = np.zeros(nlist + 1, np.int64)
    vs, idl = [], []
    for l in range(nlist):
        if b <= l < e and l in pls:
            v, i_ = ix.export_list(l)
            vs.append(v); idl.append(i_)
            off[l + 1] = off[l] + v.shape[0]
        else:
            off[l + 1] = off[l]
    lx = np.concatenate(vs, 0) if vs else np.zeros((0, d), np.float32)
    lids = np.concatenate(idl, 0) if idl else np.zeros(0, np.int64)
    threads = max(1, cores // world)
    tt = time.time()
    Do, Io = o.ivfflat_search(oracle_lib.L2, cent, off, lx, lids, xo, k, nprobe, nthreads=min(threads, oracle_q))
    cpu_s = time.time() - tt
    if world > 1:
        gd = [torch.empty((oracle_q, k), dtype=torch.float32, device=dev) for _ in range(world)]
        gi = [torch.empty((oracle_q, k), dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(gd, torch.from_numpy(Do).to(dev))
        dist.all_gather(gi, torch.from_numpy(Io).to(dev))
        Dm, Im = shard_host.merge_topk(torch.stack(gd).cpu().numpy(), torch.stack(gi).cpu().numpy(), k)
        cs = torch.tensor([cpu_s], dtype=torch.float64, device=dev)
        dist.all_reduce(cs, op=dist.ReduceOp.MAX)
        cpu_s = float(cs.item())
    else:
        Dm, Im = Do, Io
    line = None
    if rank == 0:
        fr = [float(r[0]) / peak() for r in roofs]
        line = {"config": f"cfg5 IVF-Flat L2 {n * world}x{d} ({n} rows per GPU x {world} GPUs), nlist={nlist} ({nlist_per_gpu} per GPU), nprobe={nprobe}, batch={nq}, top-{k}, list-sharded (b200vs_shard_*)",
                "scaled_from": None if (n == 12_500_000 and world == 8) else "100M x 1536 over 8 GPUs (12.5 M rows per GPU)",
                "assumptions": "nlist / nprobe are not given in BASELINE.json: 16384 / 64 at 8 GPUs (SURVEY 8d); synthetic U[0,1) rows generated on the device",
                "n_gpus": world, "rows_total": int(cnt.item()), "index_bytes_per_gpu": mem,
                "qps_device": nq * steps / (ms / 1e3), "ms_per_batch_device": ms / steps, "batches_in_flight": L,
                "qps_single_stream": nq * steps / (ms1 / 1e3), "qps_e2e": nq * steps / (e2e_ms / 1e3), "ms_per_batch_e2e": e2e_ms / steps,
                "roofline_per_rank": {"bound": "hbm", "kernel": "tc_scan_kernel capture pass", "achieved_gbs": [float(r[0]) for r in roofs], "kernel_ms": [float(r[1]) for r in roofs],
                                      "peak": peak(), "frac": fr, "frac_min": min(fr), "frac_max": max(fr), "fallback_queries": [int(r[2]) for r in roofs],
                                      "algorithmic_bytes_rank0": alg},
                "phase_ms_rank0": {a: round(v, 4) for a, v in ph.items()}, "phase_ms_all_ranks": phs, "host_enqueue_ms_per_batch_median": host_enqueue_ms,
                "verify_tile_vs_exact_scan": {"queries": int(verify), "ids_bit_exact": bool(np.array_equal(It, Ie)),
                                              "dist_bit_exact": bool(np.array_equal(Dt.view(np.uint32), De.view(np.uint32)))},
                "verify_vs_oracle": {"queries": int(oracle_q), "ids_bit_exact": bool(np.array_equal(It[:oracle_q], Im)),
                                     "dist_bit_exact": bool(np.array_equal(Dt[:oracle_q].view(np.uint32), Dm.view(np.uint32))),
                                     "recall_at_10": float(np.mean([len(set(a_) & set(b_)) / k for a_, b_ in zip(It[:oracle_q], Im)])),
                                     "how": "CPU oracle on every rank's probed lists (exported list by list), per-rank top-k merged with the MergeSearchResults rule"},
                "cpu_baseline": {"qps": oracle_q / cpu_s, "cores": threads * world, "kind": "port",
                                 "sample": f"{oracle_q} queries; each rank's oracle scans only the probed lists that rank owns ({threads} threads per rank, max over ranks)"},
                "build_seconds": {"train": t_train, "plan_pass": t_plan, "add_pass": t_add}, "search_stats": st}
    sh.close()
    return line


def cfg_calc(o, cores):
    """the UtilService distance matrix (VectorCalcDistance), 1024 x 1024 x 768, host pointers in and out."""
    nl = nr = 1024
    d = 768
    left, right = rnd(nl, d, 11), rnd(nr, d, 12)
    out = {}
    for name, metric in (("l2", b200vs.L2), ("cosine", b200vs.COSINE)):
        b200vs.calc_distance(b200vs.ALGORITHM_FAISS, metric, left, right)  # warm-up (context, allocations)
        t = time.time()
        reps = 5
        for _ in range(reps):
            got = b200vs.calc_distance(b200vs.ALGORITHM_FAISS, metric, left, right)
        gpu_s = (time.time() - t) / reps
        ns = 64  # bounded CPU sample: 64 left rows against all right rows
        t = time.time()
        want, _, _ = o.calc_distance(1, oracle_lib.L2 if metric == b200vs.L2 else oracle_lib.COSINE, left[:ns], right)
        cpu_s = (time.time() - t) * nl / ns
        out[name] = {"pairs_per_s_e2e": nl * nr / gpu_s, "ms_e2e": gpu_s * 1e3, "bit_exact_vs_oracle": bool(np.array_equal(got[:ns].view(np.uint32), want.view(np.uint32))),
                     "cpu_pairs_per_s_1thread": nl * nr / cpu_s}
    return {"config": "calc_distance", "workload": f"VectorCalcDistance {nl} x {nr} x {d} f32, ALGORITHM_FAISS, host buffers (H2D + D2H inside)", **out}


def cfg_brute(o, cores):
    """BruteForceSearch over 200K x 768 rows streamed in 2048-row tiles (FLAGS_vector_index_bruteforce_batch_count)."""
    n, d, nq, k, tile = 200_000, 768, 256, 10, 2048
    xb, xq = rnd(n, d, 21), rnd(nq, d, 22)
    ids = np.arange(1, n + 1, dtype=np.int64)

    def run(t):
        sc = b200vs.BruteForceScan(b200vs.L2, d, xq, k)
        for a in range(0, n, t):
            sc.push(xb[a:a + t], ids[a:a + t])
        return sc.finish()

    run(tile)
    res = {}
    for t in (tile, 32768):
        t0 = time.time()
        gd, gi = run(t)
        res[f"tile_{t}"] = {"seconds": time.time() - t0, "rows_per_s": n / (time.time() - t0), "queries_x_rows_per_s": nq * n / (time.time() - t0)}
    ns = 16
    t0 = time.time()
    wd, wi = o.flat_search(oracle_lib.L2, xb, ids, xq[:ns], k, nthreads=cores)
    cpu_s = (time.time() - t0) * nq / ns
    return {"config": "bruteforce_scan", "workload": f"BruteForceSearch L2 {n} x {d} streamed from host, batch {nq}, top-{k}", **res,
            "ids_bit_exact_vs_oracle": bool(np.array_equal(gi[:ns], wi)), "dist_bit_exact_vs_oracle": bool(np.array_equal(gd[:ns].view(np.uint32), wd.view(np.uint32))),
            "cpu_baseline_seconds_scaled": cpu_s, "cpu_cores": cores}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--pq-n", type=int, default=10_000_000)
    ap.add_argument("--hnsw-n", type=int, default=1_000_000)
    ap.add_argument("--hnsw-build-threads", type=int, default=64)
    ap.add_argument("--cfg5-rows", type=int, default=12_500_000, help="cfg5: database rows per GPU")
    ap.add_argument("--cfg5-batch", type=int, default=4096)
    ap.add_argument("--cfg5-steps", type=int, default=10)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    o = oracle_lib.load()
    cores = os.cpu_count() or 1
    fns = {"1": lambda: cfg1(o, cores), "2": lambda: cfg2(o, cores), "3": lambda: cfg3(o, cores, a.pq_n), "4": lambda: cfg4(o, cores, a.hnsw_n, a.hnsw_build_threads),
           "calc": lambda: cfg_calc(o, cores), "brute": lambda: cfg_brute(o, cores),
           "5": lambda: cfg5(o, cores, a.cfg5_rows, steps=a.cfg5_steps, nq=a.cfg5_batch)}
    rank = int(os.environ.get("RANK", 0))
    for c in a.configs.split(","):
        t0 = time.time()
        try:
            line = fns[c]()
        except Exception as e:  # keep going: one line per config
            import traceback
            line = {"config": f"cfg{c}", "error": repr(e), "traceback": traceback.format_exc()[-1500:]}
        if line is None or rank != 0:  # multi-rank configs: rank 0 prints
            continue
        line["wall_seconds"] = time.time() - t0
        s = json.dumps(line)
        print(s, flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(s + "\n")


if __name__ == "__main__":
    main()
