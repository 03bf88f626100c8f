"""Tensor-path latency over batch sizes (C2-shaped tree): device-resident scan time per batch."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

d = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
pts = synth.uniform(n, d, 2, np.float32)
bt = pn.BallTree.euclidean(pts)
for nq in (2048, 4096, 16384, 65536, 100000, 262144, 1000000):
    Q = synth.uniform(nq, d, 3, np.float32)
    bt.query_batch(Q, 10)
    t0 = time.perf_counter(); bt.query_batch(Q, 10); wall = time.perf_counter() - t0
    c = bt.counters()
    print(f"d={d} n={n} nq={nq}: scan {c['scan_ms']:.2f} ms, wall {wall*1e3:.2f} ms, {nq/wall/1e6:.3f} M q/s, reranks/query {c['rerank_pairs']/nq:.0f}", flush=True)
