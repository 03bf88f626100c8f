"""One process driving all GPUs of the box through pn_multi_* (include/petal_b200.h): BallTree 10M x 128 f32 sharded by subtree
over N devices, ONE batch of 1M queries, k = 10, host buffers in and out.  PEER exchange (merge kernels read the other
devices' lists over NVLink, no collective) and, for comparison, the NCCL slice exchange.  Prints one JSON line.
    python scripts/multi_arm.py N [n_points] [n_queries]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"
import torch
import petal_neighbors_b200 as pn  # noqa: F401
from petal_neighbors_b200 import parallel, synth

N = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
d, k = 128, 10
pts = synth.uniform_torch(n, d, 2, torch.float32).cpu().numpy()
Q = synth.uniform_torch(nq, d, 3, torch.float32).cpu().numpy()
torch.cuda.empty_cache()
t0 = time.perf_counter()
m = parallel.MultiGpuBallTree(pts, list(range(N)), mode=parallel.PN_SHARD_BY_SUBTREE)
create_s = time.perf_counter() - t0
out = {"workload": f"pn_multi: BallTree {n} x {d} f32 sharded by subtree over {N} devices from ONE process, one batch of {nq} queries, k={k}, host buffers",
       "n_gpus": N, "create_seconds": create_s}
ref = None
for name, ex in (("peer", None), ("nccl_slice", parallel.PN_EXCHANGE_SLICE)):
    if ex is not None:
        m.set_exchange(ex)
    m.query_batch(Q, k)   # warm-up at full size (buffers of every rank at their final size, NCCL channels)
    t0 = time.perf_counter()
    idx, dist = m.query_batch(Q, k)
    wall = time.perf_counter() - t0
    st = m.stats()
    out[name] = {"wall_ms": wall * 1e3, "value": nq / wall, "unit": "queries/s", "device_ms_max": max(s["total_ms"] for s in st),
                 "scan_ms_max": max(s["scan_ms"] for s in st), "merge_ms_max": max(s["merge_ms"] for s in st),
                 "exchange_ms_max": max(s["exchange_ms"] for s in st), "nccl_calls": sum(s["nccl_calls"] for s in st),
                 "nccl_bytes": sum(s["nccl_bytes_sent"] for s in st), "peer_mib_read": sum(s["peer_mib"] for s in st),
                 "chunks": st[0]["n_chunks"]}
    if ref is None:
        ref = (idx, dist)
    else:
        out[name]["equals_peer_result"] = bool(np.array_equal(idx, ref[0]) and np.array_equal(dist, ref[1]))
m.close()
print(json.dumps(out), flush=True)
