"""N > 1 host path on CPU: world_size-2 `gloo` run of the list-sharded search plumbing — row exchange to list
owners, per-shard top-k, ONE all_gather, merge — checked against a single-process brute force."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200vs import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _brute(xb, ids, xq, k):
    d = ((xq[:, None, :].astype(np.float64) - xb[None, :, :].astype(np.float64)) ** 2).sum(-1).astype(np.float32)
    out_d = np.zeros((xq.shape[0], k), np.float32)
    out_i = np.full((xq.shape[0], k), -1, np.int64)
    for q in range(xq.shape[0]):
        o = np.lexsort((ids, d[q]))[:k]
        out_d[q, :len(o)] = d[q][o]
        out_i[q, :len(o)] = ids[o]
    return out_d, out_i


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n, d, nlist_per_rank, nq, k = 600, 8, 4, 16, 5
    nlist = nlist_per_rank * world
    cent = np.random.default_rng(7).random((nlist, d)).astype(np.float32)  # replicated centroids
    xb = rng.random((n, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.int64) + rank * n + 1
    asg = ((xb[:, None, :] - cent[None, :, :]) ** 2).sum(-1).argmin(1)
    order, counts = shard.split_rows_by_owner(asg, nlist_per_rank, world)
    # exchange rows to their list owner (all_gather of variable-size chunks keeps the test backend-agnostic)
    send = [(xb[order][counts[:r].sum():counts[:r + 1].sum()], ids[order][counts[:r].sum():counts[:r + 1].sum()]) for r in range(world)]
    gathered = [None] * world
    dist.all_gather_object(gathered, send)
    mine_x = np.concatenate([g[rank][0] for g in gathered], 0)
    mine_i = np.concatenate([g[rank][1] for g in gathered], 0)
    # every row this rank now holds belongs to a list it owns
    a2 = ((mine_x[:, None, :] - cent[None, :, :]) ** 2).sum(-1).argmin(1)
    assert (shard.list_owner(a2, nlist_per_rank) == rank).all()
    xq = np.random.default_rng(9).random((nq, d)).astype(np.float32)  # same batch on every rank
    ld, li = _brute(mine_x, mine_i, xq, k)  # exhaustive probe of the local shard
    td, ti = torch.from_numpy(ld), torch.from_numpy(li)
    gd = [torch.empty_like(td) for _ in range(world)]
    gi = [torch.empty_like(ti) for _ in range(world)]
    dist.all_gather(gd, td)
    dist.all_gather(gi, ti)
    md, mi = shard.merge_topk(np.stack([t.numpy() for t in gd]), np.stack([t.numpy() for t in gi]), k)
    # ground truth over the union of all shards
    allx = [None] * world
    dist.all_gather_object(allx, (xb, ids))
    gx = np.concatenate([a[0] for a in allx], 0)
    gids = np.concatenate([a[1] for a in allx], 0)
    wd, wi = _brute(gx, gids, xq, k)
    ok = bool(np.array_equal(mi, wi) and np.allclose(md, wd, rtol=1e-6))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_list_sharded_search_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_merge_rule_ties_and_padding():
    pd = np.array([[[0.5, 1.0, 0.0]], [[0.5, 0.7, 2.0]]], np.float32)  # [nparts=2, nq=1, k=3]
    pi = np.array([[[9, 4, -1]], [[3, 8, 1]]], np.int64)
    d, i = shard.merge_topk(pd, pi, 4)
    assert list(i[0]) == [3, 9, 8, 4] and np.allclose(d[0], [0.5, 0.5, 0.7, 1.0])
    d, i = shard.merge_topk(pd, pi, 6)
    assert list(i[0]) == [3, 9, 8, 4, 1, -1]
