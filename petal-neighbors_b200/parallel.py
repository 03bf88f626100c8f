"""Multi-GPU drivers: one process per GPU, torch.distributed for the plumbing.

* Query sharding (default whenever the flattened tree fits one GPU): the tree is replicated, the
  queries are split into contiguous slices, there is NO data-path collective.
* Point sharding by subtree ("points larger than one GPU's HBM"): rank r keeps subtree r at depth
  log2(world) of the ball tree (pn_build_opts.shard_depth / shard_index, the reference's
  mid = (start+end)/2 split, src/ball_tree.rs:535-537); every rank answers ALL queries on its shard
  with global indices; the per-shard sorted top-k lists are all-gathered (the one exchange step)
  and merged by (distance, index) with pn_merge_topk_dev.

Nothing here computes distances or selections on the CPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import BallTree, merge_topk_dev


def query_slice(n_queries: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the query batch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_queries, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_depth(world: int) -> int:
    d = world.bit_length() - 1
    if world < 1 or (1 << d) != world:
        raise ValueError("point sharding by subtree needs a power-of-two number of ranks")
    return d


def allgather_lists(local_idx: torch.Tensor, local_dist: torch.Tensor, group=None):
    """The exchange step: every rank contributes its [nq, k] sorted lists, every rank receives
    [world, nq, k].  Works on CUDA tensors (NCCL) and, for the host-logic tests, CPU tensors (gloo)."""
    world = dist.get_world_size(group)
    nq = local_idx.shape[0]
    rest = tuple(local_idx.shape[1:])
    gi = torch.empty((world * nq,) + rest, dtype=local_idx.dtype, device=local_idx.device)  # concatenation along dim 0
    gd = torch.empty((world * nq,) + rest, dtype=local_dist.dtype, device=local_dist.device)
    dist.all_gather_into_tensor(gi, local_idx.contiguous(), group=group)
    dist.all_gather_into_tensor(gd, local_dist.contiguous(), group=group)
    return gi.view((world, nq) + rest), gd.view((world, nq) + rest)


class ShardedBallTree:
    """Ball tree whose points are sharded by subtree over the ranks of a process group."""

    def __init__(self, points, group=None, device=None, **opts):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.cuda.current_device() if device is None else device
        self.tree = BallTree.euclidean(points, device=self.device, shard_depth=shard_depth(self.world),
                                       shard_index=self.rank, **opts)
        self.dtype = self.tree.dtype
        self.dim = self.tree.dim

    def query_batch_dev(self, q_dev: torch.Tensor, k: int):
        """q_dev: [nq, d] CUDA tensor (all queries, identical on every rank).  Returns
        (idx [nq, k] int64, dist [nq, k]) CUDA tensors holding the merged global result."""
        nq = q_dev.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        li = torch.empty((nq, k), dtype=torch.int64, device=q_dev.device)
        ld = torch.empty((nq, k), dtype=q_dev.dtype, device=q_dev.device)
        self.tree.query_knn_dev(q_dev.data_ptr(), nq, q_dev.stride(0), k, li.data_ptr(), ld.data_ptr(), stream=stream, sync=True)
        gi, gd = allgather_lists(li, ld, self.group)
        oi = torch.empty_like(li)
        od = torch.empty_like(ld)
        merge_topk_dev(self.dtype, self.device, gi.data_ptr(), gd.data_ptr(), self.world, nq, k, oi.data_ptr(), od.data_ptr(),
                       stream=stream, sync=True)
        return oi, od

    def query_batch(self, Q, k: int):
        q_dev = torch.from_numpy(np.ascontiguousarray(Q, dtype=self.dtype)).cuda(self.device)
        oi, od = self.query_batch_dev(q_dev, k)
        return oi.cpu().numpy().astype(np.uint64), od.cpu().numpy()
