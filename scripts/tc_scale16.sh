# d = 16 only: per-role cycle accounting and leg isolation of the narrow-row configuration (profile build),
# for the half-tile-stage kernel (PN_TC_HALF unset) and the previous one-stage kernel (PN_TC_HALF=0)
export PN_B200_LIB=$PWD/petal-neighbors_b200/lib/libpetal_b200_prof.so
for half in 1 0; do
for dbg in ${DBGS:-0 1 2 3}; do
PN_TC_HALF=$half PN_TC_DEBUG=$dbg timeout 300 python - <<PY 2>&1 | grep -E "profile|scan|Error|error" | sed "s/^/[half=$half dbg=$dbg] /"
import sys, numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
d, n, nq = 16, 1000000, 75776
pts = synth.uniform(n, d, 2, np.float32)
bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
Q = synth.uniform(nq, d, 3, np.float32)
for k in (1, 10):
    bt.query_batch(Q, k)
    sys.stderr.flush()
    print(f"--- d={d} nq={nq} k={k}", file=sys.stderr, flush=True)
    bt.query_batch(Q, k)
    c = bt.counters(); print(f"d={d} n={n} nq={nq} k={k} scan {c['scan_ms']:.2f} ms exact evals/query {c['rerank_pairs']/nq:.0f}", file=sys.stderr, flush=True)
PY
done
done
