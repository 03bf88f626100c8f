"""Host-side check of the tensor filter's rounding bound (csrc/tc_filter.cuh).

The filter may only ever ADD candidates: for every (query, point) pair the fp16 contraction value D~^2 must not
exceed the exact squared distance (scaled units) by more than E_q, so that `exact <= kth` implies `D~^2 <= Theta_q`.
This restates the operand construction of build_aaug_kernel / build_baug_kernel in numpy (fp16 round-to-nearest
operands, three-piece norms, fp32 accumulation emulated by a float32 dot product in a worst-ish order) and checks
the inequality over several distributions.  No GPU involved."""
import numpy as np
import pytest


def split3(x):
    h1 = x.astype(np.float16)
    r1 = (x - h1.astype(np.float32)).astype(np.float32)
    h2 = r1.astype(np.float16)
    h3 = (r1 - h2.astype(np.float32)).astype(np.float32).astype(np.float16)
    return h1, h2, h3


def operands(P, Q):
    """Returns (A rows, B rows, E_q, scale) exactly as the device kernels build them."""
    n, d = P.shape
    c = (P.astype(np.float64).mean(axis=0)).astype(np.float32)
    maxabs = np.abs(P - c).max()
    ex = int(np.frexp(np.float32(maxabs))[1]) if maxabs > 0 and np.isfinite(maxabs) else 0
    s = np.float32(np.ldexp(1.0, -ex))
    kp = (d + 6 + 31) // 32 * 32

    def norms(X):
        V = ((X - c) * s).astype(np.float32)
        nrm = np.zeros(len(X), np.float32)
        for j in range(d):                      # sequential fp32 accumulation, as on the device
            nrm = (nrm + V[:, j] * V[:, j]).astype(np.float32)
        return V, nrm

    Vp, np2 = norms(P)
    Vq, nq2 = norms(Q)
    pmax = np.float32(np.sqrt(np2).max() * np.float32(1.000001))
    B = np.zeros((n, kp), np.float16)
    B[:, :d] = (np.float32(-2.0) * Vp).astype(np.float16)
    B[:, d:d + 3] = np.float16(1)               # the six norm slots follow the data dimensions, zero padding last
    B[:, d + 3], B[:, d + 4], B[:, d + 5] = split3(np2)
    A = np.zeros((len(Q), kp), np.float16)
    A[:, :d] = Vq.astype(np.float16)
    A[:, d], A[:, d + 1], A[:, d + 2] = split3(nq2)
    A[:, d + 3:d + 6] = np.float16(1)
    qn = (np.sqrt(nq2) * np.float32(1.000001)).astype(np.float32)
    sn = qn + pmax
    E = (np.float32(1.01 * 0.001953125) * qn * pmax + np.float32(6.2e-05) * np.float32(np.sqrt(d)) * (qn + 2 * pmax)
         + np.float32((kp + 8) * 4.76837158203125e-07) * sn * sn).astype(np.float32)
    assert (qn <= 200).all(), "test data must stay in the fp16 range (out-of-range queries get E = +inf on the device)"
    return A, B, E, s


@pytest.mark.parametrize("case", ["uniform16", "uniform128", "offset_cluster", "tiny_scale", "wide_range", "near_dupes"])
def test_filter_never_drops_a_candidate(case):
    rng = np.random.default_rng(hash(case) % (1 << 31))
    if case == "uniform16":
        P, Q = rng.random((3000, 16), np.float32), rng.random((200, 16), np.float32)
    elif case == "uniform128":
        P, Q = rng.random((1500, 128), np.float32), rng.random((100, 128), np.float32)
    elif case == "offset_cluster":            # large common offset: centring must remove it
        P = (1000 + 0.01 * rng.standard_normal((3000, 32))).astype(np.float32)
        Q = (1000 + 0.01 * rng.standard_normal((200, 32))).astype(np.float32)
    elif case == "tiny_scale":
        P, Q = (1e-6 * rng.random((3000, 20))).astype(np.float32), (1e-6 * rng.random((200, 20))).astype(np.float32)
    elif case == "wide_range":                # coordinates spanning many binades, queries outside the hull
        P = (rng.standard_normal((3000, 24)) * np.logspace(-3, 0, 24)).astype(np.float32)
        Q = (3 * rng.standard_normal((200, 24)) * np.logspace(-3, 0, 24)).astype(np.float32)
    else:                                     # near duplicates: differences far below the fp16 resolution
        base = rng.random((300, 16), np.float32)
        P = np.repeat(base, 10, axis=0) + (1e-5 * rng.standard_normal((3000, 16))).astype(np.float32)
        Q = base[:200] + (1e-5 * rng.standard_normal((200, 16))).astype(np.float32)
    A, B, E, s = operands(P, Q)
    # fp32-accumulated contraction of the fp16 operands (the tensor core accumulates in fp32; numpy's float32 matmul
    # stands in for one admissible accumulation order, float64 for the exact sum of the rounded products)
    D32 = A.astype(np.float32) @ B.astype(np.float32).T
    D64 = A.astype(np.float64) @ B.astype(np.float64).T
    exact = ((Q.astype(np.float64)[:, None, :] - P.astype(np.float64)[None, :, :]) ** 2).sum(axis=2) * float(s) ** 2
    for D in (D32.astype(np.float64), D64):
        excess = D - exact                    # what the filter value adds to the exact squared distance (scaled units)
        ratio = excess / E[:, None].astype(np.float64)
        assert ratio.max() <= 1.0, f"{case}: filter value exceeds exact + E_q (worst ratio {ratio.max():.3f})"
    # and the bound is not absurdly loose on well-conditioned data
    if case == "uniform16":
        assert np.median(E) < 0.05 * np.median(np.sort(exact, axis=1)[:, 9])
