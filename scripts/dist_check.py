"""torchrun check of the in-library multi-GPU path (one process per GPU): NCCL id shipped through torch.distributed,
sharded k-NN (both exchange modes) and replication against the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py"""
import os
import sys

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
comm = parallel.Comm.from_torch(device=local)
pyoracle.build()
n, d, nq, k = 400000, 64, 200000, 10
pts = synth.fast_gaussian_mixture(n, d, 9, n_centers=256, sigma=0.1, clip=True)
Q = synth.fast_gaussian_mixture(nq, d, 10, n_centers=256, sigma=0.1, clip=True)
sample = np.arange(0, nq, nq // 300)[:300]
oi, od = pyoracle.brute_knn(pts, Q[sample], k)
st = parallel.ShardedBallTree(pts, comm)
qd = torch.from_numpy(Q).cuda()
ok = True
for mode in (parallel.PN_EXCHANGE_ALLGATHER, parallel.PN_EXCHANGE_SLICE):
    gi, gd = st.query_batch_dev(qd, k, exchange=mode)
    gi, gd = gi.cpu().numpy(), gd.cpu().numpy()
    if mode == parallel.PN_EXCHANGE_ALLGATHER:
        good = np.array_equal(gi[sample].astype(np.uint64), oi.astype(np.uint64)) and np.array_equal(gd[sample].view(np.uint32), od.view(np.uint32))
    else:
        lo, hi = parallel.query_slice(nq, rank, world)
        m = (sample >= lo) & (sample < hi)
        good = np.array_equal(gi[sample[m] - lo].astype(np.uint64), oi[m].astype(np.uint64)) and np.array_equal(gd[sample[m] - lo].view(np.uint32), od[m].view(np.uint32))
    print(f"rank {rank} mode {mode}: parity {good} stats {st.stats}", flush=True)
    ok &= good
full = pn.BallTree.euclidean(pts, device=local) if rank == 0 else None
rep = parallel.replicate(full, comm, 0)
ri, rd = rep.query_batch(Q[sample], k)
good = np.array_equal(ri, oi.astype(np.uint64)) and np.array_equal(rd.view(np.uint32), od.view(np.uint32))
print(f"rank {rank} replicate: parity {good} info {rep.info()}", flush=True)
ok &= good
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print("ALL OK" if int(t) == world else "FAILED", flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
