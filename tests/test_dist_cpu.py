"""world_size-2 gloo test of the row-sharding host logic (runs on CPU).

The search callable here is the ORACLE (tests may use it as the checker): what is under test
is the partition / broadcast / gather plumbing that bench.py uses with NCCL on the GPU box.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pq_huffman_b200 import dist_knn, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, k, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_ffi
    orc = oracle_ffi.Oracle()
    if rank == 0:
        x = torch.from_numpy(synth.sift_like(n, d, seed=77))
    else:
        x = torch.empty((n, d), dtype=torch.float32)

    def search(xt, kk, lo, hi):
        a, dd = orc.knn_rows(xt.numpy(), None, kk, lo, hi, n_thread=2)
        return torch.from_numpy(a), torch.from_numpy(dd)

    ids, dists = dist_knn.sharded_self_join(x, k, search)
    if rank == 0:
        np.savez(out_path, ids=ids.numpy(), dists=dists.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 64])
def test_two_rank_sharded_search_equals_single(tmp_path, oracle, n):
    d, k = 24, 7
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, d, k, out), nprocs=2, join=True)
    got = np.load(out)
    ea, ed = oracle.knn(synth.sift_like(n, d, seed=77), None, k)
    assert np.array_equal(got["ids"], ea)
    assert np.array_equal(got["dists"].view(np.uint32), ed.view(np.uint32))


def test_row_slices_cover_exactly():
    for n in (0, 1, 7, 8, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                lo, hi, per = dist_knn.row_slice(n, world, r)
                assert lo == min(n, r * per) and lo <= hi <= n
                seen += hi - lo
            assert seen == n
