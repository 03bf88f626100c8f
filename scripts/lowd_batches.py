"""Low-dimensional trees (the pruned SIMT scan): wall, device and scan time of host-buffer k-NN calls over batch sizes.
usage: python scripts/lowd_batches.py [n] [d] [k]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
pts = synth.uniform(n, d, 2, np.float32)
bt = pn.BallTree.euclidean(pts)
for nq in (1, 64, 512, 4096, 32768, 262144):
    Q = synth.uniform(nq, d, 3, np.float32)
    bt.query_batch(Q, k)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        bt.query_batch(Q, k)
        best = min(best, time.perf_counter() - t0)
    c = bt.counters()
    print(json.dumps({"n": n, "d": d, "k": k, "nq": nq, "wall_ms": round(best * 1e3, 3), "device_ms": round(c["device_ms"], 3), "scan_ms": round(c["scan_ms"], 3),
                      "launches": c["kernel_launches"], "pairs_per_query": c["pairs"] / nq, "node_visits": c["node_visits"]}), flush=True)
