"""One pruned k-NN call on the C3 data with a ball tree (for ncu launch lists)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
n = nq = int(os.environ.get("N", 1_000_000)); d = int(os.environ.get("D", 64)); k = int(os.environ.get("K", 10))
pts = synth.gaussian_mixture_torch(n, d, 5, n_centers=1024, sigma=0.05, center_seed=4)
q = synth.gaussian_mixture_torch(nq, d, 6, n_centers=1024, sigma=0.05, center_seed=4)
bt = pn.BallTree.euclidean(pts, prune=pn.PN_PRUNE_ON)
oi = torch.empty((nq, k), dtype=torch.int64, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
for it in range(2):
    bt.query_knn_dev(q.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), sync=True)
print(bt.counters())
