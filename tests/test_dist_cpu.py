"""world_size-2 gloo test (CPU) of the gallery-sharded retrieval logic with the oracle as the local scorer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcr_b200 import dist as ddist
from dcr_b200 import synthetic
from oracle import similarity as osim


def _oracle_local(q, g, k, base):
    v, i = osim.sim_topk(q.numpy(), g.numpy(), k)
    return torch.from_numpy(v), torch.from_numpy(i + base)


def _oracle_merge(s, i, k):
    v, j = osim.merge_topk(s.numpy(), i.numpy(), k)
    return torch.from_numpy(v), torch.from_numpy(j)


def _worker(rank, world, port, nq, ng, d, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, g = synthetic.descriptors(nq, ng, d, seed=5)
    qlo, qhi = ddist.shard_bounds(nq, rank, world)
    glo, ghi = ddist.shard_bounds(ng, rank, world)
    v, i = ddist.sharded_topk(q[qlo:qhi], g[glo:ghi], k, glo, _oracle_local, _oracle_merge)
    if rank == 1:       # a non-zero rank reports: every rank must hold the full answer
        np.savez(out, v=v.numpy(), i=i.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_topk_equals_single(tmp_path):
    nq, ng, d, k = 37, 501, 64, 5
    out = str(tmp_path / "r1.npz")
    mp.spawn(_worker, args=(2, _free_port(), nq, ng, d, k, out), nprocs=2, join=True)
    got = np.load(out)
    q, g = synthetic.descriptors(nq, ng, d, seed=5)
    v, i = osim.sim_topk(q.numpy(), g.numpy(), k)
    assert np.array_equal(got["i"], i)
    np.testing.assert_array_equal(got["v"], v)


def test_shard_bounds_cover():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            b = [ddist.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
