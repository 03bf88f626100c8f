"""SIMT vs tensor path on small batches (C2-shaped tree): end-to-end wall time per batch through the host-buffer API."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

d = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
pts = synth.uniform(n, d, 2, np.float32)
trees = {"simt": pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_SIMT), "tensor": pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)}
for nq in (1, 32, 128, 256, 512, 1024, 2048):
    Q = synth.uniform(nq, d, 3, np.float32)
    out = []
    for name, bt in trees.items():
        bt.query_batch(Q, 10)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); r = bt.query_batch(Q, 10); best = min(best, time.perf_counter() - t0)
        out.append(f"{name} {best*1e3:.3f} ms (scan {bt.counters()['scan_ms']:.3f})")
    print(f"d={d} n={n} nq={nq}: " + " | ".join(out), flush=True)
