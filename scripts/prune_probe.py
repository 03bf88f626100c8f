"""Pruned vs dense tensor scan on the C3 data (1M x 64 f32 mixture) with a ball tree: time, pairs/(N*Q), reranks."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
n = nq = int(os.environ.get("N", 1_000_000)); d = int(os.environ.get("D", 64))
pts = synth.gaussian_mixture_torch(n, d, 5, n_centers=1024, sigma=0.05, center_seed=4)
q = synth.gaussian_mixture_torch(nq, d, 6, n_centers=1024, sigma=0.05, center_seed=4)
st = torch.cuda.Stream()
for k in (1, 10):
    for name, pr in (("off", pn.PN_PRUNE_OFF), ("auto", pn.PN_PRUNE_AUTO), ("on", pn.PN_PRUNE_ON)):
        bt = pn.BallTree.euclidean(pts, prune=pr)
        oi = torch.empty((nq, k), dtype=torch.int64, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        ms = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); bt.query_knn_dev(q.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), stream=st.cuda_stream, sync=False); e1.record(st)
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        bt.query_knn_dev(q.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), stream=st.cuda_stream, sync=True)
        c = bt.counters()
        print(json.dumps(dict(k=k, prune=name, ms=min(ms), scan_ms=c["scan_ms"], pairs_over_NQ=c["pairs"] / (float(n) * nq), rerank_per_query=c["rerank_pairs"] / nq,
                              launches=c["kernel_launches"], chk=int(oi.sum().item()) & 0xFFFFFFFF, build_s=bt.info()["build_seconds"])), flush=True)
        del bt
