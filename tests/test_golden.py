"""Committed golden fixtures (tests/golden/):
  reference_kat.json -- the reference's own known-answer vectors for the hot path, with file:line citations;
  synth_golden.npz   -- oracle outputs on seeded synthetic inputs of each BASELINE config's shape (made by
                        tests/golden/make_synth_golden.py), with a checksum of the inputs.
CPU: the oracle must reproduce both (pins the oracle, the brute-force tie order and the input generator).
GPU: the engine, through the C ABI, must reproduce both bit for bit."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))["cases"]
EPS = np.finfo(np.float64).eps

_spec = importlib.util.spec_from_file_location("make_synth_golden", os.path.join(HERE, "golden", "make_synth_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


def run_kat(mod, case):
    pts = np.array(case["points"], dtype=np.float64)
    q = np.array(case["query"], dtype=np.float64)
    tree = (mod.BallTree if case["tree"] == "ball" else mod.VantagePointTree).euclidean(pts)
    if case["op"] == "query":
        idx, dist = tree.query(q, case["k"])
        idx, dist = list(np.asarray(idx).tolist()), np.asarray(dist, dtype=np.float64)
    elif case["op"] == "query_nearest":
        i, d = tree.query_nearest(q)
        idx, dist = [int(i)], np.array([d], dtype=np.float64)
    else:
        idx, dist = sorted(np.asarray(tree.query_radius(q, case["r"])).tolist()), None
    if "idx" in case:
        assert idx == case["idx"], case["cite"]
    if "idx_set" in case:
        assert idx == sorted(case["idx_set"]), case["cite"]
    if "dist" in case:
        want = np.array(case["dist"], dtype=np.float64)
        tol = case.get("abs_tol", EPS * case.get("eps", 1))
        assert dist.shape == want.shape and np.all(np.abs(dist - want) <= tol), case["cite"]


@pytest.mark.parametrize("case", KAT, ids=[f"{i}:{c['op']}" for i, c in enumerate(KAT)])
def test_oracle_reference_kat(oracle, case):
    run_kat(oracle, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", KAT, ids=[f"{i}:{c['op']}" for i, c in enumerate(KAT)])
def test_engine_reference_kat(case):
    import petal_neighbors_b200 as pn
    run_kat(pn, case)


def golden():
    return np.load(os.path.join(HERE, "golden", "synth_golden.npz"))


@pytest.mark.parametrize("spec", gen.CASES, ids=[c[0] for c in gen.CASES])
def test_oracle_matches_synth_golden(oracle, spec):
    name, kind, dtype, n, d, nq, kr, seeds = spec
    g = golden()
    pts, Q = gen.inputs(kind, dtype, n, d, nq, seeds)
    assert gen.checksum(pts, Q) == str(g[name + "/sha256"]), "the synthetic input generator changed"
    for key, val in gen.expected(kind, pts, Q, kr).items():
        want = g[name + "/" + key]
        assert val.dtype == want.dtype and np.array_equal(val.view(np.uint8), want.view(np.uint8)), (name, key)
    if kind == "vp_nearest":  # the reference's VP search must agree with brute force on distinct distances
        t = oracle.VantagePointTree.euclidean(pts)
        for i in range(8):
            j, dd = t.query_nearest(Q[i])
            assert j == int(g[name + "/idx"][i, 0]) and dtype(dd) == g[name + "/dist"][i, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1, 2])
@pytest.mark.parametrize("spec", gen.CASES, ids=[c[0] for c in gen.CASES])
def test_engine_matches_synth_golden(spec, algo):
    import petal_neighbors_b200 as pn
    name, kind, dtype, n, d, nq, kr, seeds = spec
    if algo == 2 and dtype != np.float32:
        pytest.skip("the tensor path is f32 only")
    g = golden()
    pts, Q = gen.inputs(kind, dtype, n, d, nq, seeds)
    if kind == "radius":
        offs, ind = pn.BallTree.euclidean(pts, algo=algo).query_radius_batch(Q, dtype(kr))
        assert np.array_equal(offs, g[name + "/offsets"]) and np.array_equal(ind, g[name + "/indices"])
        return
    tree = (pn.VantagePointTree if kind == "vp_nearest" else pn.BallTree).euclidean(pts, algo=algo)
    if kind == "vp_nearest":
        idx, dist = tree.query_nearest_batch(Q)
        idx, dist = idx.reshape(-1, 1), dist.reshape(-1, 1)
    else:
        idx, dist = tree.query_batch(Q, int(kr))
    want_d = g[name + "/dist"]
    assert np.array_equal(idx.astype(np.uint64), g[name + "/idx"])
    assert np.array_equal(dist.view(np.uint8), want_d.view(np.uint8)), "distances are not bit-identical"
