"""torchrun --nproc-per-node G scripts/dist_sharded_bench.py [n_total] [nq] [d]
C5-shaped run (BASELINE config 5 at reduced size): points sharded by subtree over G GPUs, every rank answers
all queries on its shard, per-shard top-k all-gathered over NCCL and merged by the merge kernel."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn  # noqa: E402
from petal_neighbors_b200 import parallel, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
k = 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.perf_counter()
pts = synth.gaussian_mixture(n, d, 9, n_centers=4096, sigma=0.1, center_seed=11, clip=True) if d >= 64 else synth.uniform(n, d, 9)
Q = synth.gaussian_mixture(nq, d, 10, n_centers=4096, sigma=0.1, center_seed=11, clip=True) if d >= 64 else synth.uniform(nq, d, 10)
gen_s = time.perf_counter() - t0
t0 = time.perf_counter()
st = parallel.ShardedBallTree(pts, device=local, host_threads=max(1, (os.cpu_count() or 8) // world))
build_s = time.perf_counter() - t0
q_dev = torch.from_numpy(Q).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for _ in range(2):
    st.query_batch_dev(q_dev, k)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record(stream)
for _ in range(reps):
    oi, od = st.query_batch_dev(q_dev, k)
e1.record(stream)
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
scan_ms = st.tree.counters()["scan_ms"]
if rank == 0:
    from oracle import pyoracle
    s = np.arange(0, nq, max(1, nq // 32))[:32]
    bi, bd = pyoracle.brute_knn(pts, Q[s], k)
    ok = bool(np.array_equal(oi.cpu().numpy()[s].astype(np.uint64), bi.astype(np.uint64))
              and np.array_equal(od.cpu().numpy()[s].view(np.uint32), bd.view(np.uint32)))
    print(json.dumps({"config": f"C5-shaped: BallTree {n} x {d} f32 mixture sharded by depth-{parallel.shard_depth(world)} subtree over {world} GPUs, "
                                f"{nq} queries (all ranks), k={k}, NCCL all-gather + merge kernel",
                      "ms_per_batch_max_over_ranks": float(ms), "queries_per_s": nq / (float(ms) * 1e-3), "local_scan_ms_rank0": scan_ms,
                      "shard_points_rank0": st.tree.info()["n_points"], "parity_sample": ok, "gen_s": gen_s, "build_s": build_s,
                      "allgather_bytes_per_rank": nq * k * 12}), flush=True)
dist.barrier()
dist.destroy_process_group()
