"""Reference vs two-means partition under the pruned tensor scan on the C3 data (1M x 64 f32 mixture, ball handle):
time, pairs/(N*Q), reranks, the build-time estimates and what building the second partition costs."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
n = nq = int(os.environ.get("N", 1_000_000)); d = int(os.environ.get("D", 64))
nc = int(os.environ.get("CENTERS", 1024)); sigma = float(os.environ.get("SIGMA", 0.05))
pts = synth.gaussian_mixture_torch(n, d, 5, n_centers=nc, sigma=sigma, center_seed=4)
q = synth.gaussian_mixture_torch(nq, d, 6, n_centers=nc, sigma=sigma, center_seed=4)
st = torch.cuda.Stream()
chk = {}
only = os.environ.get("RUNS"); reps = int(os.environ.get("REPS", 3))
for k in [int(x) for x in os.environ.get("KS", "1,10").split(",")]:
    for name, opts in (("reference/auto", dict(partition=pn.PN_PARTITION_REFERENCE)),
                       ("reference/tiles", dict(partition=pn.PN_PARTITION_REFERENCE, prune=pn.PN_PRUNE_ON)),
                       ("two_means/auto", dict(partition=pn.PN_PARTITION_TWO_MEANS)),
                       ("two_means/tiles", dict(partition=pn.PN_PARTITION_TWO_MEANS, prune=pn.PN_PRUNE_ON)),
                       ("auto/auto", dict())):
        if only and name not in only.split(","):
            continue
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bt = pn.BallTree.euclidean(pts, **opts)
        torch.cuda.synchronize(); build_wall = time.perf_counter() - t0
        oi = torch.empty((nq, k), dtype=torch.int64, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        ms = []
        for it in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); bt.query_knn_dev(q.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), stream=st.cuda_stream, sync=False); e1.record(st)
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        bt.query_knn_dev(q.data_ptr(), nq, d, k, oi.data_ptr(), od.data_ptr(), stream=st.cuda_stream, sync=True)
        c = bt.counters(); inf = bt.info()
        s = (int(oi.sum().item()) & 0xFFFFFFFF, float(od.double().sum().item()))
        chk.setdefault(k, s)
        print(json.dumps(dict(k=k, run=name, ms=min(ms), scan_ms=c["scan_ms"], pairs_over_NQ=c["pairs"] / (float(n) * nq),
                              rerank_per_query=c["rerank_pairs"] / nq, launches=c["kernel_launches"], same_answer=(s == chk[k]),
                              partition=inf["tensor_partition"], seeded=inf["prune_seeded"], tiles=inf["prune_tiles"],
                              est_tile_frac=inf["est_tile_frac"], est_group_tile_frac=inf["est_group_tile_frac"],
                              est_seed_candidates=inf["est_seed_candidates"], build_s=inf["build_seconds"], build_wall_s=build_wall,
                              device_mb=inf["device_bytes"] / 1e6)), flush=True)
        del bt
