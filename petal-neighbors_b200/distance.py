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
