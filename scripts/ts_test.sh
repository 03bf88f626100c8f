# TS-form MMA experiment (A operand in tensor memory): needs lib/libpetal_b200_prof.so (-DPN_TC_PROFILE)
export PN_B200_LIB=$PWD/petal-neighbors_b200/lib/libpetal_b200_prof.so
for TS in 0 1; do
PN_TC_TS=$TS timeout 300 python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
from oracle import pyoracle
pyoracle.build()
for n, d, nq, k in ((20000, 128, 600, 10), (20000, 100, 300, 1)):
    pts = synth.uniform(n, d, 5, np.float32); Q = synth.uniform(nq, d, 6, np.float32)
    bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
    idx, dist = bt.query_batch(Q, k)
    oi, od = pyoracle.brute_knn(pts, Q, k)
    ok = np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    print("TS=$TS parity", n, d, nq, k, ok, "reranks/q", bt.counters()["rerank_pairs"]/nq, flush=True)
for d, n, nq, k in ((128, 2000000, 151552, 10), (128, 2000000, 151552, 1), (100, 1000000, 151552, 10)):
    pts = synth.uniform(n, d, 2, np.float32); Q = synth.uniform(nq, d, 3, np.float32)
    bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
    bt.query_batch(Q, k); bt.query_batch(Q, k); a = bt.counters()['scan_ms']; bt.query_batch(Q, k); b = bt.counters()['scan_ms']
    print(f"TS=$TS d={d} n={n} nq={nq} k={k}: scan {min(a,b):.2f} ms", flush=True)
PY
done
