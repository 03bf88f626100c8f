"""Host-buffer k-NN call (pinned buffers) against the device-resident call on one shape: wall clock and device time per call.
usage: python scripts/e2e_chunks.py n d nq k [reps]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import _ffi, synth

n, d, nq, k = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
pts = synth.uniform_torch(n, d, 2, torch.float32)
tree = pn.BallTree.euclidean(pts)
del pts
q_dev = synth.uniform_torch(nq, d, 3, torch.float32)
idx_dev = torch.empty((nq, k), dtype=torch.int64, device="cuda")
dist_dev = torch.empty((nq, k), dtype=torch.float32, device="cuda")
q_pin = q_dev.cpu().pin_memory()
idx_pin = torch.empty((nq, k), dtype=torch.int64).pin_memory()
dist_pin = torch.empty((nq, k), dtype=torch.float32).pin_memory()
fn = _ffi.lib().pn_balltree_query_f32
out = {"n": n, "d": d, "nq": nq, "k": k, "halves": os.environ.get("PN_HOST_HALVES", "0"), "dev_ms": [], "host_wall_ms": [], "host_device_ms": []}
for it in range(reps):
    tree.query_knn_dev(q_dev.data_ptr(), nq, d, k, idx_dev.data_ptr(), dist_dev.data_ptr(), sync=True)
    out["dev_ms"].append(round(tree.counters()["device_ms"], 2))
    t0 = time.perf_counter()
    rc = fn(tree._h, q_pin.data_ptr(), nq, d, k, idx_pin.data_ptr(), dist_pin.data_ptr())
    out["host_wall_ms"].append(round((time.perf_counter() - t0) * 1e3, 2))
    assert rc == 0
    out["host_device_ms"].append(round(tree.counters()["device_ms"], 2))
out["same"] = bool(torch.equal(idx_pin, idx_dev.cpu()) and torch.equal(dist_pin, dist_dev.cpu()))
print(json.dumps(out), flush=True)
