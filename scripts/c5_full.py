"""BASELINE config 5 at FULL size: BallTree 100M x 128 f32 "SIFT-shaped" points (mixture of 4096 centres, sigma 0.1,
clipped to [0,1)), sharded by the eight depth-3 subtrees over 8 B200s, 10M queries, k = 10, per-shard lists exchanged
over NCCL inside the library and merged.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/c5_full.py > profiles/r02_c5_full.json

Every rank generates ALL points in its own HBM (51 GB; counter-based generator, so identical everywhere), the library
applies the first three split levels to all of them on the device and keeps subtree `rank` (pn_build_opts.shard_depth /
shard_index), then builds that shard's tree.  Parity: every rank checks sampled queries of its slice against a
brute-force evaluation in torch over all 100M points -- the sequential fold over the 128 dimensions as separate
multiply and add kernels (bit-identical to Euclidean::distance, src/distance.rs:26-35), selection on the
(distance bits, index) key.  Environment: C5_N, C5_Q, C5_SAMPLE shrink the run for dry runs."""
import json
import os
import sys
import time

if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"   # keep NCCL's version banner off stdout (the JSON line)
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn  # noqa: F401
from petal_neighbors_b200 import parallel, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = parallel.Comm.from_torch(device=local)
N = int(os.environ.get("C5_N", 100_000_000)); Q = int(os.environ.get("C5_Q", 10_000_000)); S = int(os.environ.get("C5_SAMPLE", 256))
d, k = 128, 10
mix = dict(n_centers=4096, sigma=0.1, center_seed=4, clip=True)

t0 = time.perf_counter()
pts = synth.gaussian_mixture_torch(N, d, 9, **mix)
torch.cuda.synchronize()
gen_s = time.perf_counter() - t0
t0 = time.perf_counter()
st = parallel.ShardedBallTree(pts, comm)
torch.cuda.synchronize()
build_s = time.perf_counter() - t0
info = st.tree.info()
q = synth.gaussian_mixture_torch(Q, d, 10, **mix)
torch.cuda.synchronize(); dist.barrier()

# warm-up on a slice of the batch, then ONE timed pass over all Q queries per exchange mode
for mode in (parallel.PN_EXCHANGE_SLICE, parallel.PN_EXCHANGE_ALLGATHER):   # also sets up NCCL's channels for both patterns
    st.query_batch_dev(q[: min(Q, 200_000)], k, exchange=mode)
res = {}
out = None
for name, mode in (("slice", parallel.PN_EXCHANGE_SLICE), ("allgather", parallel.PN_EXCHANGE_ALLGATHER)):
    if name == "allgather" and os.environ.get("C5_ALLGATHER", "1") == "0":
        continue
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    gi, gd = st.query_batch_dev(q, k, exchange=mode)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    s = dict(st.stats)
    t = torch.tensor([s["total_ms"], s["scan_ms"], s["exchange_ms"], s["merge_ms"], wall * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    b = torch.tensor([s["nccl_bytes_sent"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(b)
    res[name] = {"ms_per_batch": float(t[0]), "queries_per_s": Q / (float(t[0]) * 1e-3), "scan_ms": float(t[1]), "exchange_ms": float(t[2]),
                 "merge_ms": float(t[3]), "wall_ms": float(t[4]), "nccl_bytes_all_ranks": int(b[0]), "chunks": s["n_chunks"],
                 "nccl_calls_per_rank": s["nccl_calls"]}
    if name == "slice":
        out = (gi, gd)
    else:
        lo, hi = parallel.query_slice(Q, rank, world)
        same = bool(torch.equal(gi[lo:hi], out[0]) and torch.equal(gd[lo:hi], out[1]))
        res[name]["equals_slice_mode"] = same
    del gi, gd

# parity on sampled queries of this rank's slice: brute force over ALL points in torch
lo, hi = parallel.query_slice(Q, rank, world)
samp = torch.arange(lo, hi, max(1, (hi - lo) // max(1, S // world)), device="cuda")[: max(1, S // world)]
qs = q[samp]                                     # [s, d]
best = torch.full((samp.numel(), k), torch.iinfo(torch.int64).max, dtype=torch.int64, device="cuda")
CH = 1 << 20
for c0 in range(0, N, CH):
    P = pts[c0:c0 + CH]                          # [m, d]
    acc = torch.zeros((samp.numel(), P.shape[0]), dtype=torch.float32, device="cuda")
    for j in range(d):                           # sequential fold, separate multiply and add (no FMA in eager mode)
        diff = qs[:, j:j + 1] - P[:, j][None, :]
        acc = acc + diff * diff
    dd = torch.sqrt(acc)
    key = (dd.view(torch.int32).to(torch.int64) << 32) | torch.arange(c0, c0 + P.shape[0], device="cuda", dtype=torch.int64)[None, :]
    best = torch.topk(torch.cat([best, key], dim=1), k, dim=1, largest=False).values
ref_idx = best & 0xFFFFFFFF
ref_dist = (best >> 32).to(torch.int32).view(torch.float32)
gi, gd = out
ok = bool(torch.equal(gi[samp - lo], ref_idx) and torch.equal(gd[samp - lo].view(torch.int32), ref_dist.view(torch.int32)))
okt = torch.tensor([1 if ok else 0, samp.numel()], device="cuda")
dist.all_reduce(okt)
if rank == 0:
    print(json.dumps({"config": f"BallTree {N} x {d} f32 SIFT-shaped mixture sharded by subtree over {world} B200, {Q} queries, k={k} (BASELINE config 5)",
                      "n_gpus": world, "points_per_rank": info["n_points"], "generate_seconds": gen_s, "shard_build_seconds": build_s,
                      "device_bytes_per_rank": info["device_bytes"], "tree_levels": info["n_levels"],
                      "parity": {"ranks_ok": int(okt[0]), "ranks": world, "sampled_queries": int(okt[1]),
                                 "checker": "torch brute force over all points: sequential non-FMA fold, (distance, index) keys"},
                      **res}), flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
