# Timeline trace of CTA 0 of the tensor filter (profile build): 64 tiles from PN_TC_TRACE_T0, per role and event.
export PN_B200_LIB=$PWD/petal-neighbors_b200/lib/libpetal_b200_prof.so
for k in 1 10; do
PN_TC_TRACE=gpurun_out/trace_d${D:-16}_k$k.bin timeout 300 python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
d = ${D:-16}
pts = synth.uniform(1000000, d, 2, np.float32)
bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
Q = synth.uniform(37888, d, 3, np.float32)
bt.query_batch(Q, $k); bt.query_batch(Q, $k)
print("scan ms", bt.counters()["scan_ms"])
PY
done
