"""petal_neighbors::distance -- the `Metric` trait and the `Euclidean` type (src/distance.rs:9-55).

These are the drop-in *types*: the trees accept `Euclidean()` as their metric.  The tree queries
never call these scalar methods (neither does the reference: only `distance` is used, and the
engine evaluates it on the GPU); they exist so code written against the reference's trait keeps
working, and they follow the same sequential non-FMA fold.  `Cosine` (src/distance.rs:76-122) is
not offered: it is not a metric, the north star names Euclidean only, and there is no CPU
fallback to run a tree with it.
"""
from __future__ import annotations

import numpy as np


class Metric:
    """trait Metric<A> src/distance.rs:9-14."""

    def distance(self, x1, x2):
        raise NotImplementedError

    def rdistance(self, x1, x2):
        raise NotImplementedError

    def rdistance_to_distance(self, d):
        raise NotImplementedError

    def distance_to_rdistance(self, d):
        raise NotImplementedError


class Euclidean(Metric):
    """struct Euclidean src/distance.rs:16-55."""

    def __eq__(self, other):
        return isinstance(other, Euclidean)

    def __hash__(self):
        return hash("Euclidean")

    def __repr__(self):
        return "Euclidean"

    def rdistance(self, x1, x2):
        x1 = np.asarray(x1)
        x2 = np.asarray(x2, dtype=x1.dtype)
        t = x1.dtype.type
        s = t(0)
        for a, b in zip(x1, x2):  # zip truncates to the shorter input, like the reference
            diff = t(a) - t(b)
            s = t(s + t(diff * diff))
        return s

    def distance(self, x1, x2):
        return np.sqrt(self.rdistance(x1, x2))

    def rdistance_to_distance(self, d):
        return np.sqrt(d)

    def distance_to_rdistance(self, d):
        return d * d


def pairwise(x, metric: Metric = None, device: int = -1):
    """distance::pairwise (src/distance.rs:58-74) on the GPU: dense symmetric n x n matrix of exact
    Euclidean distances (bit-identical to the reference fold).  Only `Euclidean` is offered."""
    from . import _ffi, _check  # late import: _check lives in the package root
    if metric is not None and not isinstance(metric, Euclidean):
        raise TypeError("only Euclidean is offered by the B200 engine")
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError("x must be 2-D")
    if x.dtype not in (np.float32, np.float64):
        raise TypeError("A must be f32 or f64")
    if x.shape[0] and x.strides[1] != x.dtype.itemsize and x.shape[1] > 1:
        x = np.ascontiguousarray(x)
    n, d = x.shape
    out = np.zeros((n, n), dtype=x.dtype)
    if n >= 2:
        fn = getattr(_ffi.lib(), "pn_pairwise_f32" if x.dtype == np.float32 else "pn_pairwise_f64")
        rs = x.strides[0] // x.dtype.itemsize
        _check(fn(device, x.ctypes.data, n, d, rs, out.ctypes.data))
    return out
