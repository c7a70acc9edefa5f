"""Row (e) on real NCCL: item-sharded tokenisation (all-gather of ids) and all-reduced k-means over 2 GPUs must equal
the single-GPU result.  Needs >= 2 GPUs (`gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`); skipped otherwise."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import inputs as I

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, D, K, L = 5001, 768, 256, 3          # odd N: ragged shards


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from rq_vae_recommender_b200 import parallel
    x, cbs = I.rq_problem(N, D, K, L, seed=31)
    lo, hi = parallel.shard_bounds(N, world, rank)
    xs = torch.from_numpy(x[lo:hi]).cuda()
    tok = parallel.CorpusTokenizer([torch.from_numpy(c).cuda() for c in cbs])
    ids = tok.tokenize_sharded(xs, N)
    usage = parallel.codebook_usage(ids[lo:hi].contiguous(), K)
    xk = torch.from_numpy(I.randn(32, N, 32)[lo:hi]).cuda()
    np.random.seed(9); torch.manual_seed(10)
    w = torch.zeros(64, 32, device="cuda")
    parallel.sharded_kmeans_init_(w, xk, N, max_iters=10)
    q.put((rank, ids.cpu().numpy(), usage.cpu().numpy(), w.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_tokenize_and_kmeans_over_nccl_match_single_gpu():
    from rq_vae_recommender_b200 import ops, parallel
    from oracle import rq_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    x, cbs = I.rq_problem(N, D, K, L, seed=31)
    single = ops.rq_tokenize_tc(torch.from_numpy(x).cuda(), [torch.from_numpy(c).cuda() for c in cbs]).cpu().numpy()
    xk = torch.from_numpy(I.randn(32, N, 32)).cuda()
    np.random.seed(9); torch.manual_seed(10)
    w1 = torch.zeros(64, 32, device="cuda")
    parallel.sharded_kmeans_init_(w1, xk, N, max_iters=10)
    for rank, ids, usage, w in outs:
        assert ids.shape == (N, L) and np.array_equal(ids, single)          # corpus order, same ids as one GPU
        assert np.array_equal(usage, O.codebook_usage(single, K))           # all-reduced usage counts
        assert np.allclose(w, w1.cpu().numpy(), atol=1e-6)                  # every rank: the single-GPU centroids
    assert np.array_equal(outs[0][3], outs[1][3])
