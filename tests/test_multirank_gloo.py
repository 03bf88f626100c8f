"""Host-side logic of the N > 1 paths with world_size = 2 over gloo on the CPU (no GPU):
query slicing, subtree shard membership from the host builder, the all-gather exchange, and a
merge of the gathered lists.  The per-shard answers come from the oracle (test infrastructure);
the merge here is a numpy restatement of pn_merge_topk_dev's contract, used only as the checker."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_numpy(gi, gd, k):
    """k smallest of the union of the lists by (distance, index); padding (2^64-1, inf) sorts last."""
    world, nq, kk = gi.shape
    ai = gi.transpose(1, 0, 2).reshape(nq, world * kk)
    ad = gd.transpose(1, 0, 2).reshape(nq, world * kk)
    order = np.lexsort((ai, ad), axis=1)[:, :k]
    return np.take_along_axis(ai, order, 1), np.take_along_axis(ad, order, 1)


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import petal_neighbors_b200 as pn
    from petal_neighbors_b200 import parallel, synth
    from oracle import pyoracle

    n, d, nq, k = 6000, 8, 101, 10
    pts = synth.uniform(n, d, 61, np.float32)
    Q = synth.uniform(nq, d, 62, np.float32)
    # --- query sharding: slices tile the batch, results concatenate to the full answer
    lo, hi = parallel.query_slice(nq, rank, world)
    li, ld = pyoracle.brute_knn(pts, Q[lo:hi], k)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    assert sum(int(s) for s in sizes) == nq
    # --- point sharding by subtree: shard membership from the host builder (no device needed)
    shard = pn.BallTree.euclidean(pts, host_only=True, shard_depth=parallel.shard_depth(world), shard_index=rank)
    ids = np.sort(shard.layout()["ids"].astype(np.int64))
    local_pts = pts[ids]
    si, sd = pyoracle.brute_knn(local_pts, Q, k)                       # shard-local answer ...
    gidx = np.where(si == np.iinfo(np.uintp).max, si, ids[np.minimum(si, len(ids) - 1).astype(np.int64)].astype(np.uintp))
    ti = torch.from_numpy(gidx.astype(np.int64))                        # ... with GLOBAL indices
    td = torch.from_numpy(sd)
    gi, gd = parallel.allgather_lists(ti, td)                           # the exchange step, over gloo
    assert gi.shape == (world, nq, k)
    mi, md = _merge_numpy(gi.numpy().astype(np.uint64), gd.numpy(), k)
    oi, od = pyoracle.brute_knn(pts, Q, k)
    ok = np.array_equal(mi, oi.astype(np.uint64)) and np.array_equal(md.view(np.uint32), od.view(np.uint32))
    n_all = torch.tensor([len(ids)])
    dist.all_reduce(n_all)
    ok = ok and int(n_all) == n
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.fail"), "w").close()
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    from oracle import pyoracle
    pyoracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]


def test_query_slice_and_depth():
    import petal_neighbors_b200  # noqa: F401
    from petal_neighbors_b200 import parallel
    for nq in (0, 1, 7, 100, 1000003):
        for world in (1, 2, 4, 8):
            parts = [parallel.query_slice(nq, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == nq
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert [parallel.shard_depth(w) for w in (1, 2, 4, 8)] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        parallel.shard_depth(3)
