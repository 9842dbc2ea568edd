"""Multi-GPU plumbing for multi-pair matching: one process per GPU, pairs sharded across ranks,
ONE broadcast of the shared descriptor/keypoint set, no collective on the data path; per-rank results go to
the host and only small per-pair summaries are gathered on rank 0.

torch.distributed is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests); the matching
itself is the C ABI (`Context.match_pairs`) or, in CPU tests, any `compute_fn` with the same signature.
"""
import numpy as np


def shard_pairs(pairs, rank, world, mode="contiguous"):
    """Rows of `pairs` (n x 2) owned by `rank`.  Image pairs are independent, so any partition is valid:
    'contiguous' keeps a rank's pairs adjacent in image index (best L2 / H2D locality for sliding-window lists),
    'strided' (round-robin) balances ragged keypoint counts."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    n = pairs.shape[0]
    if mode == "contiguous":
        lo, hi = (n * rank) // world, (n * (rank + 1)) // world
        idx = np.arange(lo, hi)
    elif mode == "strided":
        idx = np.arange(rank, n, world)
    else:
        raise ValueError(mode)
    return idx, np.ascontiguousarray(pairs[idx])


def broadcast_image_set(image_set, src, device, group=None):
    """Broadcast {offsets int64[n+1], sizes int32[n,2], desc uint8[N,32], kp float32[N,2]} from rank `src`.

    image_set is the dict on `src` (numpy arrays or torch tensors) and None elsewhere.  Returns a dict of torch
    tensors on `device` (desc, kp) plus numpy offsets/sizes on every rank.  Exactly one collective per array."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        off = np.ascontiguousarray(image_set["offsets"], dtype=np.int64)
        sizes = np.ascontiguousarray(image_set["sizes"], dtype=np.int32).reshape(-1, 2)
        meta = [(off, sizes)]
    dist.broadcast_object_list(meta, src=src, group=group)
    off, sizes = meta[0]
    total = int(off[-1])
    if rank == src:
        desc = torch.as_tensor(image_set["desc"]).to(device=device, dtype=torch.uint8).reshape(total, 32).contiguous()
        kp = torch.as_tensor(image_set["kp"]).to(device=device, dtype=torch.float32).reshape(total, 2).contiguous()
    else:
        desc = torch.empty((total, 32), dtype=torch.uint8, device=device)
        kp = torch.empty((total, 2), dtype=torch.float32, device=device)
    dist.broadcast(desc, src=src, group=group)
    dist.broadcast(kp, src=src, group=group)
    return dict(offsets=off, sizes=sizes, desc=desc, kp=kp)


def gather_pair_summaries(local_idx, local_summary, n_pairs, dst=0, group=None):
    """Gather per-pair int32 summaries (e.g. n_inliers, best_hyp) onto rank `dst` in global pair order.
    local_summary: int array [len(local_idx), k].  Returns int32 [n_pairs, k] on dst, None elsewhere."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local_summary = np.ascontiguousarray(local_summary, dtype=np.int32).reshape(len(local_idx), -1)
    parts = [None] * world if rank == dst else None
    dist.gather_object((np.asarray(local_idx), local_summary), parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.zeros((n_pairs, local_summary.shape[1]), np.int32)
    for idx, summ in parts:
        out[idx] = summ
    return out


def match_pairs_sharded(compute_fn, image_set, pairs, shard_mode="contiguous", group=None):
    """compute_fn(image_set, local_pairs) -> dict(n_inliers, best_hyp, ...) for the rank's own pairs (per-rank
    match lists / masks stay on that rank's host); returns (local_idx, local_result, summaries_on_rank0)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    idx, local_pairs = shard_pairs(pairs, rank, world, shard_mode)
    res = compute_fn(image_set, local_pairs)
    summ = np.stack([np.asarray(res["n_inliers"], np.int32), np.asarray(res["best_hyp"], np.int32)], 1)
    return idx, res, gather_pair_summaries(idx, summ, pairs.shape[0], 0, group)
