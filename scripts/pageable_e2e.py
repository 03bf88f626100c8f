"""Host-buffer k-NN calls with PAGEABLE numpy buffers (what a Python caller passes): wall time per call.
usage: python scripts/pageable_e2e.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

for n, d, nq, k in ((10_000_000, 3, 1_000_000, 10), (1_000_000, 16, 1_000_000, 10)):
    pts = synth.uniform(n, d, 2, np.float32)
    bt = pn.BallTree.euclidean(pts)
    Q = synth.uniform(nq, d, 3, np.float32)
    walls, devs = [], []
    for it in range(4):
        t0 = time.perf_counter()
        idx, dist = bt.query_batch(Q, k)     # fresh np.empty outputs every call, as a caller would
        walls.append(round((time.perf_counter() - t0) * 1e3, 2))
        devs.append(round(bt.counters()["device_ms"], 2))
    print(json.dumps({"n": n, "d": d, "nq": nq, "k": k, "wall_ms": walls, "device_ms": devs, "scan_ms": bt.counters()["scan_ms"],
                      "queries_per_s": nq / (min(walls[1:]) * 1e-3), "chk": int(idx.sum() & 0xFFFFFFFF)}), flush=True)
