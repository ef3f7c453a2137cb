"""Query-row sharding of the kNN job over ranks (one process per GPU).

The path shards by query rows: every rank holds the whole base, owns a
contiguous slice of query rows and searches it independently; the only exchanges are the
broadcast of the base from the rank that loaded it and the gather of the per-rank neighbour
lists.  These helpers are backend-agnostic (NCCL on the GPU box, gloo in the CPU tests) and do
no arithmetic themselves: the search is a callable the caller supplies.
"""
import torch
import torch.distributed as dist


def row_slice(n, world, rank):
    """Rows [lo, hi) owned by `rank`; every rank's slot holds `per` rows (the last is ragged)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    hi = min(n, (rank + 1) * per)
    return lo, hi, per


def broadcast_base(x, src=0):
    """Base vectors travel once from the rank that loaded them to every other rank."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(x, src=src)
    return x


def gather_rows(local_a, local_d, n, k, dst=0):
    """Gather per-rank [rows x k] id / distance slices into [n x k] tensors on `dst`.

    Slices are padded to the common slot size so a plain gather works for ragged n.
    Returns (ids, dists) on dst, (None, None) elsewhere.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return local_a[:n], local_d[:n]
    lo, hi, per = row_slice(n, world, rank)
    rows = hi - lo
    pad_a = torch.zeros((per, k), dtype=local_a.dtype, device=local_a.device)
    pad_d = torch.zeros((per, k), dtype=local_d.dtype, device=local_d.device)
    pad_a[:rows].copy_(local_a[:rows])
    pad_d[:rows].copy_(local_d[:rows])
    if rank == dst:
        ga = [torch.empty_like(pad_a) for _ in range(world)]
        gd = [torch.empty_like(pad_d) for _ in range(world)]
    else:
        ga = gd = None
    dist.gather(pad_a, ga, dst=dst)
    dist.gather(pad_d, gd, dst=dst)
    if rank != dst:
        return None, None
    return torch.cat(ga, 0)[:n], torch.cat(gd, 0)[:n]


def sharded_self_join(x, k, search_rows, src=0, dst=0):
    """Whole job: broadcast x from `src`, search this rank's rows, gather on `dst`.

    `search_rows(x, k, lo, hi) -> (ids[rows x k], dists[rows x k])` runs the search for the
    rank's slice (pqh_knn_search_device on the GPU box).
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    broadcast_base(x, src)
    n = x.shape[0]
    lo, hi, _ = row_slice(n, world, rank)
    a, d = search_rows(x, k, lo, hi)
    return gather_rows(a, d, n, k, dst)
