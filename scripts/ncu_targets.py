"""Small single-launch workloads for ncu captures (round 2): python scripts/ncu_targets.py c2|t128|c3|c4
  c2   : BASELINE config 2 at the bench's own launch size (1M x 16, 1M queries, k=10), two device-resident calls
  t128 : north-star shape 10M x 128, 75 776 queries (two whole waves of 148 CTAs x 256 queries), k=10, two calls
  c3   : VantagePointTree 1M x 64 mixture, 303 104 queries, query_nearest, two calls
  c4   : BallTree::query_radius 10M x 3, 262 144 queries, r = 0.01, two calls
  c4knn: BallTree::query on the same tree, 1M device-resident queries, k = 10 (the warp-per-query scan), two calls
  c1   : BASELINE config 1: BallTree 10k x 3 f64, every point a query (query_self), k = 10, two calls"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

what = sys.argv[1]
if what in ("c2", "t128"):
    n, d, nq = (1_000_000, 16, 1_000_000) if what == "c2" else (10_000_000, 128, 75_776)
    pts = synth.uniform_torch(n, d, 2, torch.float32)
    q = synth.uniform_torch(nq, d, 3, torch.float32)
    bt = pn.BallTree.euclidean(pts)
    oi = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); od = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    for _ in range(2):
        bt.query_knn_dev(q.data_ptr(), nq, d, 10, oi.data_ptr(), od.data_ptr(), sync=True)
    print(what, bt.counters(), bt.info())
elif what == "c3":
    n, nq, d = 1_000_000, 303_104, 64
    pts = synth.fast_gaussian_mixture(n, d, 5, n_centers=1024, sigma=0.05, center_seed=4)
    q = synth.gaussian_mixture_torch(nq, d, 6, n_centers=1024, sigma=0.05, center_seed=4)
    vp = pn.VantagePointTree.euclidean(pts)
    oi = torch.empty((nq, 1), dtype=torch.int64, device="cuda"); od = torch.empty((nq, 1), dtype=torch.float32, device="cuda")
    for _ in range(2):
        vp.query_knn_dev(q.data_ptr(), nq, d, 1, oi.data_ptr(), od.data_ptr(), sync=True)
    print(what, vp.counters(), vp.info())
elif what == "c4":
    n, nq = 10_000_000, 262_144
    bt = pn.BallTree.euclidean(synth.uniform_torch(n, 3, 7, torch.float32))
    Q = synth.fast_uniform(nq, 3, 8, np.float32)
    for _ in range(2):
        offs, ind = bt.query_radius_batch(Q, np.float32(0.01))
    print(what, bt.counters(), int(offs[-1]))
elif what == "c4knn":
    n, nq, d = 10_000_000, 1_000_000, 3
    bt = pn.BallTree.euclidean(synth.uniform_torch(n, d, 7, torch.float32))
    q = synth.uniform_torch(nq, d, 8, torch.float32)
    oi = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); od = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    for _ in range(2):
        bt.query_knn_dev(q.data_ptr(), nq, d, 10, oi.data_ptr(), od.data_ptr(), sync=True)
    print(what, bt.counters(), bt.info())
elif what == "c1":
    pts = synth.uniform(10_000, 3, 1, np.float64)
    bt = pn.BallTree.euclidean(pts)
    for _ in range(2):
        idx, dist = bt.query_self(10)
    print(what, bt.counters(), int(idx.sum()))
