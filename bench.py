#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path (BASELINE.json metric: k-NN queries/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|t128|c1]

One "step" = one pass of the hot path over one batch of synthetic queries: batched exact k-NN
(k = 10) on the configuration BASELINE.json quotes the metric on, config 2: BallTree 1M x 16 f32
uniform points, 1M queries per GPU.  N > 1 shards the queries over ranks with the tree replicated
(no data-path collective; weak scaling: every rank answers its own 1M queries).

  value : whole-job queries/s with queries and outputs resident in HBM (pn_tree_query_knn_dev),
          CUDA events on the launching stream, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI call (pn_balltree_query_f32) with pinned
          host buffers -- H2D of the queries and D2H of (idx, dist) inside the timed region.
  roofline / cpu_baseline : see DESIGN.md "Measurement".

`--impl reference` times the reference's own CPU implementation of the path (the oracle's C
restatement of the Rust crate -- no Rust toolchain exists in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "knn_queries_per_sec_k10"
UNIT = "queries/s"

WORKLOADS = {
    # name: (n_points, dim, n_queries per GPU, k, dtype, generator, label)
    "c2": (1_000_000, 16, 1_000_000, 10, np.float32, "uniform",
           "BallTree 1M x 16 f32 uniform, 1M batched queries per GPU, k=10 (BASELINE config 2)"),
    "t128": (10_000_000, 128, 100_000, 10, np.float32, "uniform",
             "BallTree 10M x 128 f32 uniform, 100k batched queries per GPU, k=10 (north-star target shape)"),
    "c1": (10_000, 3, 10_000, 10, np.float64, "self",
           "BallTree 10k x 3 f64 uniform, every point a query, k=10 (BASELINE config 1)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(wl, rank):
    import petal_neighbors_b200  # noqa: F401
    from petal_neighbors_b200 import synth
    n, d, nq, k, dtype, gen, _ = WORKLOADS[wl]
    pts = synth.uniform(n, d, 2, dtype)
    if gen == "self":
        Q = pts.copy()
    else:
        Q = synth.uniform(nq, d, 3, dtype, row0=rank * nq)  # each rank owns a slice of one query stream
    return pts, Q


def algorithmic_bytes(n, d, nq, k, s, pairs):
    """SURVEY.md 8d / BASELINE.md 4: B_alg = B_min + pairs*d*s/128."""
    b_min = n * d * s + nq * d * s + nq * k * (8 + s)
    return b_min + pairs * d * s / 128.0


def cpu_reference(pts, Q, k, budget_s=20.0):
    """The reference's CPU path (oracle port) on a bounded sample of the same workload."""
    from oracle import pyoracle
    pyoracle.build()
    threads = pyoracle.max_threads()
    t0 = time.perf_counter()
    tree = pyoracle.BallTree.euclidean(pts)
    build_s = time.perf_counter() - t0
    # calibrate the sample on a few queries, then time about budget_s of CPU work
    probe = min(64, Q.shape[0])
    t0 = time.perf_counter()
    tree.query_batch(Q[:probe], k, n_threads=threads)
    per_q = (time.perf_counter() - t0) / probe
    sample = int(max(probe, min(Q.shape[0], budget_s / max(per_q, 1e-9))))
    t0 = time.perf_counter()
    _, _, ndist = tree.query_batch(Q[:sample], k, n_threads=threads)
    dt = time.perf_counter() - t0
    single = min(sample, max(16, int(3.0 / max(per_q * threads, 1e-9))))
    t0 = time.perf_counter()
    tree.query_batch(Q[:single], k, n_threads=1)
    dt1 = time.perf_counter() - t0
    return {
        "value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"first {sample} of the {Q.shape[0]} queries, reference tree (1-2 point leaves) built on all "
                  f"{pts.shape[0]} points, OpenMP outer loop over queries",
        "single_thread_value": single / dt1, "single_thread_sample": single,
        "build_seconds": build_s, "distance_evals_per_query": ndist / sample,
    }, tree, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    n, d, nq, k, dtype, gen, label = WORKLOADS[wl]
    pts, Q = make_inputs(wl, 0)
    base, tree, sample = cpu_reference(pts, Q, k, budget_s=8.0)
    threads = base["cores"]
    for _ in range(args.warmup):
        tree.query_batch(Q[:max(16, sample // 8)], k, n_threads=threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = (s * sample) % max(1, Q.shape[0] - sample + 1)
        tree.query_batch(Q[lo:lo + sample], k, n_threads=threads)
    dt = time.perf_counter() - t0
    v = args.steps * sample / dt
    base["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
        "config": {"workload": label, "step": f"{sample} queries per step on the host cores (bounded sample)"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--bucket", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--queries", type=int, default=0, help="override queries per GPU (profiling runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import petal_neighbors_b200 as pn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the B200 arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload
    if args.queries:
        w = list(WORKLOADS[wl]); w[2] = args.queries; w[6] += f" [REDUCED to {args.queries} queries: profiling run, not a bench line]"
        WORKLOADS[wl] = tuple(w)
    n, d, nq, k, dtype, gen, label = WORKLOADS[wl]
    s = np.dtype(dtype).itemsize
    tdtype = torch.float32 if dtype == np.float32 else torch.float64
    pts, Q = make_inputs(wl, rank)
    nq = Q.shape[0]
    algo = {"auto": pn.PN_ALGO_AUTO, "simt": pn.PN_ALGO_SIMT, "tensor": pn.PN_ALGO_TENSOR}[args.algo]
    ncpu = os.cpu_count() or 8
    tree = pn.BallTree.euclidean(pts, device=local, algo=algo, bucket_size=args.bucket,
                                 host_threads=max(1, ncpu // max(world, 1)))
    info = tree.info()

    stream = torch.cuda.Stream()  # explicit (non-default) stream: handle 0 would mean "the tree's own stream"
    torch.cuda.set_stream(stream)
    q_dev = torch.from_numpy(Q).cuda()
    idx_dev = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    dist_dev = torch.empty((nq, k), dtype=tdtype, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    q_pin = torch.from_numpy(Q).pin_memory()
    idx_pin = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    dist_pin = torch.empty((nq, k), dtype=tdtype).pin_memory()

    def step_dev():
        tree.query_knn_dev(q_dev.data_ptr(), nq, d, k, idx_dev.data_ptr(), dist_dev.data_ptr(),
                           stream=stream.cuda_stream, sync=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---------------- kernel-resident timing (value) ----------------
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    t_wall0 = time.time()
    evs = []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_dev()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(ms_steps))
    # one synchronous call to read the engine's work counters for exactly one step
    tree.query_knn_dev(q_dev.data_ptr(), nq, d, k, idx_dev.data_ptr(), dist_dev.data_ptr(),
                       stream=stream.cuda_stream, sync=True)
    ctr = tree.counters()

    # ---------------- end-to-end through the host-buffer C ABI (e2e) ----------------
    import ctypes as C
    from petal_neighbors_b200 import _ffi
    fn = getattr(_ffi.lib(), "pn_balltree_query_f32" if dtype == np.float32 else "pn_balltree_query_f64")

    def step_host():
        rc = fn(tree._h, q_pin.data_ptr(), nq, d, k, idx_pin.data_ptr(), dist_pin.data_ptr())
        if rc != 0:
            raise RuntimeError(_ffi.last_error())

    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ctr_host = tree.counters()
    # sanity: both paths return the same answer
    same = bool(torch.equal(idx_pin, idx_dev.cpu()) and torch.equal(dist_pin, dist_dev.cpu()))

    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        hbm_peak, bf16_peak, which = peaks()
        pairs = ctr["pairs"]
        b_alg = algorithmic_bytes(n, d, nq, k, s, pairs)
        f_alg = 2.0 * d * pairs
        scan_ms = ctr["scan_ms"] if ctr["scan_ms"] > 0 else total_ms / args.steps
        hbm_ach = b_alg / (scan_ms * 1e-3) / 1e9
        tf32_peak = bf16_peak / 2.0
        tc_ach = f_alg / (scan_ms * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                ent = json.load(open(prof)).get(wl if ctr["filter_pairs"] else wl + "_simt")
                traffic = float(ent["bytes_per_launch"]) * nq / float(ent["queries"])  # scaled to this launch size
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": world * nq * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": label, "queries_per_gpu": nq, "k": k, "sharding": f"queries x{world}, tree replicated",
                       "l2": "flushed between timed steps (256 MiB write, untimed)",
                       "bucket_size_max": info["bucket_size_max"], "tree_levels": info["n_levels"],
                       "algo": args.algo, "tree_build_seconds": info["build_seconds"]},
            "e2e": {"value": world * nq * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(ctr_host["h2d_bytes"]), "d2h_bytes_per_step": int(ctr_host["d2h_bytes"]),
                    "timing": "wall clock around the synchronous host-buffer ABI call, pinned buffers, max over ranks",
                    "matches_device_path": same},
            "gpu_launches": int(ctr["kernel_launches"]) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                         "traffic": traffic, "peak_source": f"of {which}",
                         "kernel": ("tc::knn_filter_kernel (tcgen05 FP16 filter + exact rerank)" if ctr["filter_pairs"] else "knn_tile_kernel (exact SIMT scan)") + " + merge, CUDA events inside the engine, per step",
                         "rerank_pairs": int(ctr["rerank_pairs"]),
                         "kernel_ms": scan_ms, "algorithmic_bytes": b_alg, "pairs": int(pairs),
                         "pairs_over_NQ": pairs / (float(n) * nq),
                         "tensor_frac_if_counted": tc_ach / tf32_peak, "tf32_peak_assumed_tflops": tf32_peak},
        }
        if not args.no_cpu_baseline:
            cap = 2_000_000 if d > 32 else n  # the reference tree of 10M x 128 needs ~9 GB of centroids and minutes to build
            base, _, _ = cpu_reference(pts[:cap], Q, k)
            if cap < n:
                base["sample"] += f" [reference tree built on the first {cap} of {n} points]"
            out["cpu_baseline"] = base
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
