"""Generates tests/golden/synth_golden.npz: expected outputs of the CPU oracle (brute force, (distance, index) order;
the reference-tree oracle for the radius and VP cases) on small seeded synthetic inputs, plus a checksum of the inputs
so that the generator itself is pinned.  Run from the repo root:  python tests/golden/make_synth_golden.py
The inputs are regenerated from the seeds at test time; only seeds, checksums and expected outputs are stored."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from petal_neighbors_b200 import synth  # noqa: E402

CASES = [  # name, kind, dtype, n, d, nq, k or r, seeds
    ("c1_f64_d3_k10", "knn", np.float64, 2000, 3, 64, 10, (1, 11)),
    ("c2_f32_d16_k10", "knn", np.float32, 3000, 16, 64, 10, (2, 3)),
    ("t_f32_d128_k10", "knn", np.float32, 1500, 128, 32, 10, (9, 10)),
    ("c3_f32_d64_nn", "vp_nearest", np.float32, 3000, 64, 64, 1, (5, 6)),
    ("c4_f32_d3_r", "radius", np.float32, 20000, 3, 64, 0.05, (7, 8)),
]


def inputs(kind, dtype, n, d, nq, seeds):
    if kind == "vp_nearest":
        return (synth.gaussian_mixture(n, d, seeds[0], n_centers=64, sigma=0.05, center_seed=4, dtype=dtype),
                synth.gaussian_mixture(nq, d, seeds[1], n_centers=64, sigma=0.05, center_seed=4, dtype=dtype))
    return synth.uniform(n, d, seeds[0], dtype), synth.uniform(nq, d, seeds[1], dtype)


def expected(kind, pts, Q, kr):
    if kind == "radius":
        offs, ind = pyoracle.brute_radius(pts, Q, pts.dtype.type(kr))
        return {"offsets": offs.astype(np.uint64), "indices": ind.astype(np.uint64)}
    k = int(kr)
    oi, od = pyoracle.brute_knn(pts, Q, k)
    return {"idx": oi.astype(np.uint64), "dist": od}


def checksum(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    pyoracle.build()
    out = {}
    for name, kind, dtype, n, d, nq, kr, seeds in CASES:
        pts, Q = inputs(kind, dtype, n, d, nq, seeds)
        out[name + "/sha256"] = np.array(checksum(pts, Q))
        for key, val in expected(kind, pts, Q, kr).items():
            out[name + "/" + key] = val
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "synth_golden.npz"), **out)
    print("wrote", len(out), "arrays")
