"""Runs the named BASELINE configs once on cuda:0: GPU timing + counters, exact parity against the
oracle on a seeded sample, and the reference-port CPU timing on a bounded sample.
usage: python scripts/run_configs.py c1|c3|c4|t128 [...]   -> one JSON line per config on stdout"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
from oracle import pyoracle

pyoracle.build()
TH = pyoracle.max_threads()


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def cpu_ball(pts, Q, k, budget=15.0):
    t0 = time.perf_counter()
    ref = pyoracle.BallTree.euclidean(pts)
    build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref.query_batch(Q[:32], k, n_threads=TH)
    per = (time.perf_counter() - t0) / 32
    m = int(max(32, min(len(Q), budget / per)))
    t0 = time.perf_counter()
    ref.query_batch(Q[:m], k, n_threads=TH)
    allc = m / (time.perf_counter() - t0)
    m1 = int(max(8, min(m, 4.0 / (per * TH))))
    t0 = time.perf_counter()
    ref.query_batch(Q[:m1], k, n_threads=1)
    one = m1 / (time.perf_counter() - t0)
    return dict(cpu_qps_all_cores=allc, cpu_qps_1_thread=one, cpu_cores=TH, cpu_build_s=build, cpu_sample=m), ref


def c1():
    pts = synth.uniform(10_000, 3, 1, np.float64)
    t0 = time.perf_counter(); bt = pn.BallTree.euclidean(pts); build = time.perf_counter() - t0
    dt, (idx, dist) = timed(lambda: bt.query_batch(pts, 10))
    c = bt.counters()
    cpu, ref = cpu_ball(pts, pts, 10, budget=5.0)
    oi, od, _ = ref.query_batch(pts, 10, n_threads=TH)
    return dict(config="C1 BallTree 10k x 3 f64, every point a query, k=10", gpu_qps_e2e=len(pts) / dt, gpu_ms=dt * 1e3,
                device_ms=c["device_ms"], pairs_over_NQ=c["pairs"] / 1e8, gpu_build_s=build,
                parity=bool(np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(bits(dist), bits(od))), **cpu)


def c3():
    n = nq = 1_000_000
    pts = synth.gaussian_mixture(n, 64, 5, n_centers=1024, sigma=0.05, center_seed=4)
    Q = synth.gaussian_mixture(nq, 64, 6, n_centers=1024, sigma=0.05, center_seed=4)
    out = dict(config="C3 VantagePointTree 1M x 64 f32 Gaussian mixture (1024 x sigma 0.05), 1M queries, 1-NN")
    for name, algo in (("simt", pn.PN_ALGO_SIMT), ("tensor", pn.PN_ALGO_TENSOR)):
        t0 = time.perf_counter(); vp = pn.VantagePointTree.euclidean(pts, algo=algo); build = time.perf_counter() - t0
        dt, (vi, vd) = timed(lambda: vp.query_nearest_batch(Q), reps=2)
        c = vp.counters()
        out[name] = dict(gpu_qps_e2e=nq / dt, gpu_ms=dt * 1e3, device_ms=c["device_ms"], scan_ms=c["scan_ms"],
                         pairs_over_NQ=c["pairs"] / (float(n) * nq), gpu_build_s=build)
        s = np.arange(0, nq, nq // 300)[:300]
        oi, od = pyoracle.brute_knn(pts, Q[s], 1)
        out[name]["parity_sample"] = bool(np.array_equal(vi[s], oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd[s]), bits(od[:, 0])))
        del vp
    t0 = time.perf_counter(); ref = pyoracle.VantagePointTree.euclidean(pts); out["cpu_build_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); ref.query_nearest_batch(Q[:64], n_threads=TH); per = (time.perf_counter() - t0) / 64
    m = int(max(64, min(nq, 15.0 / per)))
    t0 = time.perf_counter(); ri, rd, nd = ref.query_nearest_batch(Q[:m], n_threads=TH); out["cpu_qps_all_cores"] = m / (time.perf_counter() - t0)
    m1 = int(max(16, min(m, 4.0 / (per * TH))))
    t0 = time.perf_counter(); ref.query_nearest_batch(Q[:m1], n_threads=1); out["cpu_qps_1_thread"] = m1 / (time.perf_counter() - t0)
    out.update(cpu_cores=TH, cpu_sample=m, cpu_dist_evals_per_query=nd / m)
    return out


def c4():
    n, nq, r = 10_000_000, 1_000_000, np.float32(0.01)
    pts = synth.uniform(n, 3, 7, np.float32)
    Q = synth.uniform(nq, 3, 8, np.float32)
    t0 = time.perf_counter(); bt = pn.BallTree.euclidean(pts); build = time.perf_counter() - t0
    dt, (offs, ind) = timed(lambda: bt.query_radius_batch(Q, r), reps=2)
    c = bt.counters()
    out = dict(config="C4 BallTree::query_radius 10M x 3 f32, 1M queries, r=0.01", gpu_qps_e2e=nq / dt, gpu_ms=dt * 1e3,
               device_ms=c["device_ms"], scan_ms=c["scan_ms"], hits_per_query=float(offs[-1]) / nq, pairs_per_query=c["pairs"] / nq,
               gpu_build_s=build)
    s = np.arange(0, nq, nq // 200)[:200]
    boffs, bind = pyoracle.brute_radius(pts, Q[s], r)
    ok = True
    for t, qi in enumerate(s):
        ok &= np.array_equal(ind[offs[qi]:offs[qi + 1]], bind[boffs[t]:boffs[t + 1]].astype(np.uint64))
    out["parity_sample"] = bool(ok)
    dt, (idx, dist) = timed(lambda: bt.query_batch(Q, 10), reps=2)     # k-NN on the same tree (d = 3: pruned SIMT scan)
    c = bt.counters()
    out["knn_k10"] = dict(gpu_qps_e2e=nq / dt, gpu_ms=dt * 1e3, scan_ms=c["scan_ms"], pairs_per_query=c["pairs"] / nq)
    oi, od = pyoracle.brute_knn(pts, Q[s], 10)
    out["knn_k10"]["parity_sample"] = bool(np.array_equal(idx[s], oi.astype(np.uint64)) and np.array_equal(bits(dist[s]), bits(od)))
    t0 = time.perf_counter(); ref = pyoracle.BallTree.euclidean(pts); out["cpu_build_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); ref.query_radius_batch(Q[:2000], r, n_threads=TH); per = (time.perf_counter() - t0) / 2000
    m = int(max(2000, min(nq, 10.0 / per)))
    t0 = time.perf_counter(); ref.query_radius_batch(Q[:m], r, n_threads=TH); out["cpu_qps_all_cores"] = m / (time.perf_counter() - t0)
    m1 = int(max(500, min(m, 3.0 / (per * TH))))
    t0 = time.perf_counter(); ref.query_radius_batch(Q[:m1], r, n_threads=1); out["cpu_qps_1_thread"] = m1 / (time.perf_counter() - t0)
    out.update(cpu_cores=TH, cpu_sample=m)
    return out


def t128():
    n, nq, d = 10_000_000, 100_000, 128
    pts = synth.uniform(n, d, 2, np.float32)
    Q = synth.uniform(nq, d, 3, np.float32)
    t0 = time.perf_counter(); bt = pn.BallTree.euclidean(pts); build = time.perf_counter() - t0
    dt, (idx, dist) = timed(lambda: bt.query_batch(Q, 10), reps=2)
    c = bt.counters()
    s_ = 4
    b_alg = n * d * s_ + nq * d * s_ + nq * 10 * 12 + c["pairs"] * d * s_ / 128.0
    out = dict(config="T north-star shape: BallTree 10M x 128 f32 uniform, 100k queries, k=10", gpu_qps_e2e=nq / dt, gpu_ms=dt * 1e3,
               device_ms=c["device_ms"], scan_ms=c["scan_ms"], pairs_over_NQ=c["pairs"] / (float(n) * nq), rerank_per_query=c["rerank_pairs"] / nq,
               gpu_build_s=build, hbm_roofline_frac=b_alg / (c["scan_ms"] * 1e-3) / 6539.2e9,
               tf_equiv_tflops=2.0 * d * c["pairs"] / (c["scan_ms"] * 1e-3) / 1e12, info=bt.info())
    s = np.arange(0, nq, nq // 48)[:48]
    oi, od = pyoracle.brute_knn(pts, Q[s], 10)
    out["parity_sample"] = bool(np.array_equal(idx[s], oi.astype(np.uint64)) and np.array_equal(bits(dist[s]), bits(od)))
    del bt
    try:
        cpu, _ = cpu_ball(pts[:1_000_000], Q, 10, budget=15.0)     # reference port at N = 1M (10M needs ~9 GB of centroids and minutes to build)
        out["cpu_at_N_1M"] = cpu
    except Exception as e:  # noqa
        out["cpu_error"] = repr(e)
    return out


for name in sys.argv[1:]:
    t0 = time.perf_counter()
    try:
        res = {"c1": c1, "c3": c3, "c4": c4, "t128": t128}[name]()
    except Exception as e:  # noqa
        import traceback
        res = dict(config=name, error=repr(e), tb=traceback.format_exc()[-1500:])
    res["wall_s"] = time.perf_counter() - t0
    print(json.dumps(res), flush=True)
