"""torchrun --nproc-per-node N scripts/dist_nccl_check.py : NCCL all-gather + merge kernel for points
sharded by subtree, and query sharding with a replicated tree, both checked against the oracle."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import parallel, synth
from oracle import pyoracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (n, d, nq, k) in ((200_000, 16, 4096, 10), (50_000, 64, 1000, 10), (20_000, 3, 3000, 5)):
    pts = synth.uniform(n, d, 71, np.float32)
    Q = synth.uniform(nq, d, 72, np.float32)
    oi, od = pyoracle.brute_knn(pts, Q, k) if rank == 0 else (None, None)
    # point sharding by subtree + all-gather + merge
    st = parallel.ShardedBallTree(pts, device=local)
    t0 = time.perf_counter()
    mi, md = st.query_batch(Q, k)
    dt = time.perf_counter() - t0
    n_local = torch.tensor([st.tree.info()["n_points"]], device="cuda")
    dist.all_reduce(n_local)
    if rank == 0:
        good = np.array_equal(mi, oi.astype(np.uint64)) and np.array_equal(md.view(np.uint32), od.view(np.uint32)) and int(n_local) == n
        print(f"[subtree-sharded x{world}] n={n} d={d} nq={nq} k={k}: parity={'OK' if good else 'FAIL'} ({dt*1e3:.1f} ms)", flush=True)
        ok &= good
    # query sharding: replicated tree, slices concatenate
    full = pn.BallTree.euclidean(pts, device=local)
    lo, hi = parallel.query_slice(nq, rank, world)
    li, ld = full.query_batch(Q[lo:hi], k)
    gi = [None] * world
    dist.all_gather_object(gi, (lo, hi, li, ld))
    if rank == 0:
        ci = np.concatenate([g[2] for g in sorted(gi, key=lambda g: g[0])])
        cd = np.concatenate([g[3] for g in sorted(gi, key=lambda g: g[0])])
        good = np.array_equal(ci, oi.astype(np.uint64)) and np.array_equal(cd.view(np.uint32), od.view(np.uint32))
        print(f"[query-sharded   x{world}] n={n} d={d} nq={nq} k={k}: parity={'OK' if good else 'FAIL'}", flush=True)
        ok &= good
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
