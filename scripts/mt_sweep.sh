export PN_B200_LIB=$PWD/petal-neighbors_b200/lib/libpetal_b200_prof.so
for M in 2 4; do
PN_TC_MT=$M timeout 600 python - <<PY 2>&1 | grep "scan"
import sys, numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
for d, n, nq, k in ((64, 1000000, 303104, 1), (64, 1000000, 303104, 10), (96, 1000000, 303104, 10), (128, 1000000, 303104, 10), (40, 1000000, 303104, 10)):
    pts = synth.uniform(n, d, 2, np.float32); Q = synth.uniform(nq, d, 3, np.float32)
    bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
    bt.query_batch(Q, k); bt.query_batch(Q, k); a = bt.counters()['scan_ms']; bt.query_batch(Q, k); b = bt.counters()['scan_ms']
    print(f"MT=$M d={d} k={k}: scan {min(a,b):.2f} ms", flush=True)
PY
done
