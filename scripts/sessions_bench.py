"""Concurrent callers: T host threads, each issuing small query batches, (a) all on one handle (calls serialise on its
mutex) and (b) each on its own session (pn_tree_session: own stream and workspaces, the tree's arrays shared).
One JSON line per shape.    python scripts/sessions_bench.py [threads]"""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SHAPES = [  # n, d, dtype, batch, k, calls per thread
    (10_000, 3, np.float64, 512, 10, 200),       # BASELINE config 1's tree, small batches
    (1_000_000, 16, np.float32, 1024, 10, 40),   # config 2's tree, 1024-query batches
    (10_000_000, 3, np.float32, 4096, 10, 100),  # config 4's tree, k-NN
]
for n, d, dtype, batch, k, calls in SHAPES:
    pts = synth.uniform(n, d, 2, dtype)
    tree = pn.BallTree.euclidean(pts)
    Qs = [synth.uniform(batch, d, 10 + t, dtype) for t in range(T)]
    ref = [tree.query_batch(Q, k) for Q in Qs]
    res = {}
    for mode in ("one_handle", "sessions"):
        handles = [tree] * T if mode == "one_handle" else [tree.session() for _ in range(T)]
        for h, Q in zip(handles, Qs):
            h.query_batch(Q, k)   # warm-up: workspaces
        ok = [True] * T

        def work(t):
            for _ in range(calls):
                i, dd = handles[t].query_batch(Qs[t], k)
            ok[t] = bool(np.array_equal(i, ref[t][0]) and np.array_equal(dd, ref[t][1]))

        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        dt = time.perf_counter() - t0
        res[mode] = {"queries_per_s": T * calls * batch / dt, "calls_per_s": T * calls / dt, "identical": all(ok)}
        if mode == "sessions":
            for h in handles:
                h.close()
    print(json.dumps({"n": n, "d": d, "dtype": np.dtype(dtype).name, "threads": T, "batch": batch, "k": k, **res,
                      "speedup": res["sessions"]["queries_per_s"] / res["one_handle"]["queries_per_s"]}), flush=True)
    tree.close()
