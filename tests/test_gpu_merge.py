"""The CUDA k-way merge (b200vs_merge_topk_device) follows the same rule as b200vs.shard.merge_topk."""
import numpy as np
import pytest

import b200vs
from b200vs import shard
from gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def test_merge_kernel_matches_numpy_rule():
    require_gpu()
    import torch
    rng = np.random.default_rng(0)
    for nparts, nq, k in ((2, 7, 5), (8, 64, 10), (4, 3, 100)):
        pd = np.sort(rng.random((nparts, nq, k)).astype(np.float32), axis=2)
        pd[:, :, :2] = np.round(pd[:, :, :2], 1)  # force some ties
        pd = np.sort(pd, axis=2)
        pi = rng.permutation(nparts * nq * k).reshape(nparts, nq, k).astype(np.int64)
        pi[0, 0, k - 1] = -1
        td, ti = torch.from_numpy(pd).cuda(), torch.from_numpy(pi).cuda()
        od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        b200vs.merge_topk_device(0, nparts, nq, k, td.data_ptr(), ti.data_ptr(), od.data_ptr(), oi.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        wd, wi = shard.merge_topk(pd, pi, k)
        assert np.array_equal(oi.cpu().numpy(), wi)
        assert np.array_equal(od.cpu().numpy(), wd)
