"""GPU probe: filter hit statistics and timing of the tensor path on config-2-shaped data."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
d = int(sys.argv[3]) if len(sys.argv) > 3 else 16
pts = synth.uniform(n, d, 2, np.float32)
Q = synth.uniform(nq, d, 3, np.float32)
for algo, name in ((pn.PN_ALGO_TENSOR, "tensor"),):
    bt = pn.BallTree.euclidean(pts, algo=algo)
    for k in (10, 1):
        bt.query_batch(Q, k)
        t0 = time.perf_counter()
        idx, dist = bt.query_batch(Q, k)
        dt = time.perf_counter() - t0
        c = bt.counters()
        print(f"{name} n={n} nq={nq} d={d} k={k}: wall {dt*1e3:.1f} ms scan {c['scan_ms']:.1f} ms "
              f"hits/query {c['rerank_pairs']/nq:.1f} kth-dist mean {dist[:, -1].mean():.4f}", flush=True)
