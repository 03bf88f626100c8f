"""Times the device-resident k-NN call (pn_tree_query_knn_dev) on a few shapes; one JSON line per shape.
usage: python scripts/ab_shapes.py [tag]   (environment toggles of diagnostic experiments apply)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

tag = sys.argv[1] if len(sys.argv) > 1 else ""
SHAPES = [  # n, d, nq, k
    (1_000_000, 16, 1_000_000, 10), (1_000_000, 16, 1_000_000, 1), (1_000_000, 32, 500_000, 10),
    (1_000_000, 64, 500_000, 1), (1_000_000, 64, 500_000, 10),
]
if os.environ.get("AB_SHAPES"):
    SHAPES = [tuple(int(x) for x in s.split("x")) for s in os.environ["AB_SHAPES"].split(",")]
for n, d, nq, k in SHAPES:
    pts = synth.uniform(n, d, 2, np.float32)
    Q = synth.uniform(nq, d, 3, np.float32)
    bt = pn.BallTree.euclidean(pts, device=0)
    st = torch.cuda.Stream()
    qd = torch.from_numpy(Q).cuda()
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ms = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        bt.query_knn_dev(qd.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), stream=st.cuda_stream, sync=False)
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    bt.query_knn_dev(qd.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), stream=st.cuda_stream, sync=True)
    c = bt.counters()
    print(json.dumps(dict(tag=tag, n=n, d=d, nq=nq, k=k, ms_best=min(ms[1:]), ms_all=ms, scan_ms=c["scan_ms"],
                          rerank_per_query=c["rerank_pairs"] / nq, pairs_over_NQ=c["pairs"] / (float(n) * nq),
                          chk=int(oi.sum().item()) & 0xFFFFFFFF)), flush=True)
    del bt
