# A/B of engine builds in one GPU session: petal-neighbors_b200/lib/variants/*.so (built with different -D knobs or
# from different commits), interleaved over ROUNDS rounds; the same box, so differences of 0.5 % are meaningful.
for round in $(seq 1 ${ROUNDS:-1}); do
for so in petal-neighbors_b200/lib/variants/*.so; do
export PN_B200_LIB=$PWD/$so
echo "== $so (round $round)"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  C2 ms/step', round(d['ms_per_step'],2), 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'reranks/q', d['roofline']['rerank_pairs']/1e6)"
timeout 300 python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
for d, n, nq, k in ${SHAPES:-((128, 2000000, 151552, 10), (64, 1000000, 303104, 1), (32, 1000000, 303104, 10), (16, 1000000, 2048, 10))}:
    pts = synth.uniform(n, d, 2, np.float32); Q = synth.uniform(nq, d, 3, np.float32)
    bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
    bt.query_batch(Q, k); bt.query_batch(Q, k); a = bt.counters()['scan_ms']; bt.query_batch(Q, k); b = bt.counters()['scan_ms']
    print(f"  d={d} n={n} nq={nq} k={k}: scan {min(a,b):.2f} ms", flush=True)
PY
done
done
