"""ctypes front-end of the CPU oracle (oracle/libpetal_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  The class and method
names follow the reference API (BallTree::query / query_nearest / query_radius,
VantagePointTree::query_nearest; src/ball_tree.rs:80-142, src/vantage_point_tree.rs:88-98).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpetal_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (make -C oracle)."""
    src_newer = (not os.path.exists(_SO)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("petal_oracle.c", "petal_oracle_impl.h", "Makefile")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_sz = C.c_size_t
_psz = C.POINTER(C.c_size_t)
_ull = C.c_ulonglong


def _declare(L):
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_max_threads.restype = C.c_int
    for sfx, real in (("f32", C.c_float), ("f64", C.c_double)):
        pr = C.POINTER(real)

        def f(name, restype, argtypes):
            fn = getattr(L, f"{name}_{sfx}")
            fn.restype = restype
            fn.argtypes = argtypes

        f("orc_distance", real, [pr, pr, _sz])
        f("orc_rdistance", real, [pr, pr, _sz])
        f("orc_pairwise", None, [pr, _sz, _sz, _sz, pr])
        f("orc_node_init", None, [pr, _sz, _sz, _psz, _sz, pr, pr])
        f("orc_max_spread_column", _sz, [pr, _sz, _sz, _sz, _psz, _sz])
        f("orc_halve_node_indices", C.c_int, [_psz, _sz, pr, _sz])
        f("orc_balltree_new", C.c_void_p, [pr, _sz, _sz, _sz, _sz, C.POINTER(C.c_int)])
        f("orc_balltree_free", None, [C.c_void_p])
        f("orc_balltree_num_nodes", _sz, [C.c_void_p])
        f("orc_balltree_num_points", _sz, [C.c_void_p])
        f("orc_balltree_idx", _psz, [C.c_void_p])
        f("orc_balltree_node", None, [C.c_void_p, _sz, _psz, _psz, pr, C.POINTER(C.c_int), pr])
        f("orc_balltree_node_distance_lower_bound", real, [C.c_void_p, _sz, _sz])
        f("orc_balltree_nearest_in_subtree", C.c_int, [C.c_void_p, pr, _sz, real, _psz, pr])
        f("orc_balltree_query_nearest", None, [C.c_void_p, pr, _psz, pr])
        f("orc_balltree_query", _sz, [C.c_void_p, pr, _sz, _psz, pr, C.POINTER(_ull)])
        f("orc_balltree_query_radius", _sz, [C.c_void_p, pr, real, C.POINTER(_psz)])
        f("orc_vptree_new", C.c_void_p, [pr, _sz, _sz, _sz, _sz, C.POINTER(C.c_int)])
        f("orc_vptree_free", None, [C.c_void_p])
        f("orc_vptree_query_nearest", None, [C.c_void_p, pr, _psz, pr, C.POINTER(_ull)])
        f("orc_brute_knn", None, [pr, _sz, _sz, _sz, pr, _sz, _sz, _sz, _psz, pr, C.c_int])
        f("orc_brute_radius", None, [pr, _sz, _sz, _sz, pr, _sz, _sz, real, _psz, _psz, _psz, C.c_int])
        f("orc_balltree_query_batch", _ull, [C.c_void_p, pr, _sz, _sz, _sz, _psz, pr, C.c_int])
        f("orc_vptree_query_nearest_batch", _ull, [C.c_void_p, pr, _sz, _sz, _psz, pr, C.c_int])
        f("orc_balltree_query_radius_batch", None,
          [C.c_void_p, pr, _sz, _sz, real, _psz, _psz, _psz, C.c_int])


class ArrayError(Exception):
    """ArrayError::{Empty, NotContiguous}, src/lib.rs:9-16."""


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(f"oracle supports f32/f64, got {dtype}")


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _strides(points):
    """(row_stride, col_stride) in elements; the oracle borrows the buffer like CowArray."""
    it = points.dtype.itemsize
    rs = points.strides[0] // it if points.shape[0] > 1 else points.shape[1]
    cs = points.strides[1] // it if points.shape[1] > 1 else 1
    return rs, cs


def max_threads() -> int:
    """Host threads the baseline may use: the CPUs this process may run on.  (Not
    omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which would silently make the
    'all cores' baseline single-threaded; the batch drivers pass num_threads explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def distance(x1, x2):
    x1 = np.ascontiguousarray(x1)
    x2 = np.ascontiguousarray(x2, dtype=x1.dtype)
    sfx, real = _sfx(x1.dtype)
    d = min(x1.shape[0], x2.shape[0])  # zip truncates, src/distance.rs:27-28
    return x1.dtype.type(getattr(lib(), f"orc_distance_{sfx}")(_ptr(x1, real), _ptr(x2, real), d))


def pairwise(x):
    x = np.ascontiguousarray(x)
    sfx, real = _sfx(x.dtype)
    n, d = x.shape
    out = np.empty((n, n), dtype=x.dtype)
    getattr(lib(), f"orc_pairwise_{sfx}")(_ptr(x, real), n, d, d, _ptr(out, real))
    return out


def node_init(points, idx):
    points = np.ascontiguousarray(points)
    sfx, real = _sfx(points.dtype)
    idx = np.ascontiguousarray(idx, dtype=np.uintp)
    c = np.empty(points.shape[1], dtype=points.dtype)
    r = real()
    getattr(lib(), f"orc_node_init_{sfx}")(_ptr(points, real), points.shape[1], points.shape[1],
                                          _ptr(idx, C.c_size_t), idx.size, _ptr(c, real), C.byref(r))
    return c, points.dtype.type(r.value)


def max_spread_column(points, idx):
    points = np.ascontiguousarray(points)
    sfx, real = _sfx(points.dtype)
    idx = np.ascontiguousarray(idx, dtype=np.uintp)
    nrows, ncols = points.shape if points.ndim == 2 else (0, 0)
    v = getattr(lib(), f"orc_max_spread_column_{sfx}")(_ptr(points, real), nrows, ncols, ncols,
                                                      _ptr(idx, C.c_size_t), idx.size)
    if v == 2**64 - 1:
        raise RuntimeError("empty matrix")
    if v == 2**64 - 2:
        raise RuntimeError("index out of bounds")
    return int(v)


def halve_node_indices(idx, col):
    col = np.ascontiguousarray(col)
    sfx, real = _sfx(col.dtype)
    idx = np.ascontiguousarray(idx, dtype=np.uintp).copy()
    rc = getattr(lib(), f"orc_halve_node_indices_{sfx}")(_ptr(idx, C.c_size_t), idx.size, _ptr(col, real), 1)
    if rc != 0:
        raise OverflowError("attempt to subtract with overflow")
    return idx


class BallTree:
    """Restatement of petal_neighbors::BallTree<A, Euclidean> (src/ball_tree.rs)."""

    def __init__(self, points):
        points = np.asarray(points)
        if points.ndim != 2:
            raise ValueError("points must be 2-D")
        self._sfx, self._real = _sfx(points.dtype)
        self._points = points  # borrowed, like CowArray
        n, d = points.shape
        rs, cs = _strides(points)
        err = C.c_int(0)
        self._h = getattr(lib(), f"orc_balltree_new_{self._sfx}")(
            _ptr(points, self._real) if n else None, n, d, rs, cs, C.byref(err))
        if err.value == 1:
            raise ArrayError("array is empty")
        if err.value == 2:
            raise ArrayError("array is not contiguous in memory")
        self.n, self.d = n, d

    @classmethod
    def euclidean(cls, points):
        return cls(points)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            getattr(lib(), f"orc_balltree_free_{self._sfx}")(h)
            self._h = None

    def _fn(self, name):
        return getattr(lib(), f"{name}_{self._sfx}")

    def _q(self, q):
        q = np.ascontiguousarray(q, dtype=self._points.dtype)
        return q

    @property
    def idx(self):
        p = self._fn("orc_balltree_idx")(self._h)
        return np.ctypeslib.as_array(p, shape=(self.n,)).copy()

    def num_nodes(self):
        return int(self._fn("orc_balltree_num_nodes")(self._h))

    def num_points(self):
        return int(self._fn("orc_balltree_num_points")(self._h))

    def node(self, i):
        lo, hi, r, leaf = C.c_size_t(), C.c_size_t(), self._real(), C.c_int()
        c = np.empty(self.d, dtype=self._points.dtype)
        self._fn("orc_balltree_node")(self._h, i, C.byref(lo), C.byref(hi), C.byref(r), C.byref(leaf),
                                      _ptr(c, self._real))
        return dict(range=(lo.value, hi.value), radius=self._points.dtype.type(r.value),
                    is_leaf=bool(leaf.value), centroid=c)

    def node_distance_lower_bound(self, n1, n2):
        return self._points.dtype.type(self._fn("orc_balltree_node_distance_lower_bound")(self._h, n1, n2))

    def nearest_neighbor_in_subtree(self, q, root, radius):
        q = self._q(q)
        i, dd = C.c_size_t(), self._real()
        some = self._fn("orc_balltree_nearest_in_subtree")(self._h, _ptr(q, self._real), root,
                                                            radius, C.byref(i), C.byref(dd))
        return (int(i.value), self._points.dtype.type(dd.value)) if some else None

    def query_nearest(self, q):
        q = self._q(q)
        i, dd = C.c_size_t(), self._real()
        self._fn("orc_balltree_query_nearest")(self._h, _ptr(q, self._real), C.byref(i), C.byref(dd))
        return int(i.value), self._points.dtype.type(dd.value)

    def query(self, q, k, count=False):
        q = self._q(q)
        cap = max(1, min(k, self.n))
        oi = np.empty(cap, dtype=np.uintp)
        od = np.empty(cap, dtype=self._points.dtype)
        nd = _ull(0)
        n = self._fn("orc_balltree_query")(self._h, _ptr(q, self._real), k, _ptr(oi, C.c_size_t),
                                           _ptr(od, self._real), C.byref(nd))
        if count:
            return oi[:n].copy(), od[:n].copy(), int(nd.value)
        return oi[:n].copy(), od[:n].copy()

    def query_radius(self, q, r):
        q = self._q(q)
        out = _psz()
        n = self._fn("orc_balltree_query_radius")(self._h, _ptr(q, self._real), r, C.byref(out))
        res = np.ctypeslib.as_array(out, shape=(n,)).copy() if n else np.empty(0, dtype=np.uintp)
        lib().orc_free(out)
        return res

    # --- batch drivers (CPU baseline) ---
    def query_batch(self, Q, k, n_threads=1):
        Q = np.ascontiguousarray(Q, dtype=self._points.dtype)
        nq = Q.shape[0]
        oi = np.empty((nq, k), dtype=np.uintp)
        od = np.empty((nq, k), dtype=self._points.dtype)
        nd = self._fn("orc_balltree_query_batch")(self._h, _ptr(Q, self._real), nq, Q.shape[1], k,
                                                   _ptr(oi, C.c_size_t), _ptr(od, self._real), n_threads)
        return oi, od, int(nd)

    def query_radius_batch(self, Q, r, n_threads=1):
        Q = np.ascontiguousarray(Q, dtype=self._points.dtype)
        nq = Q.shape[0]
        counts = np.zeros(nq, dtype=np.uintp)
        f = self._fn("orc_balltree_query_radius_batch")
        f(self._h, _ptr(Q, self._real), nq, Q.shape[1], r, _ptr(counts, C.c_size_t), None, None, n_threads)
        offsets = np.zeros(nq + 1, dtype=np.uintp)
        np.cumsum(counts, out=offsets[1:])
        out = np.empty(int(offsets[-1]), dtype=np.uintp)
        f(self._h, _ptr(Q, self._real), nq, Q.shape[1], r, _ptr(counts, C.c_size_t),
          _ptr(offsets, C.c_size_t), _ptr(out, C.c_size_t), n_threads)
        return offsets, out


class VantagePointTree:
    """Restatement of petal_neighbors::VantagePointTree<A, Euclidean>."""

    def __init__(self, points):
        points = np.asarray(points)
        if points.ndim != 2:
            raise ValueError("points must be 2-D")
        self._sfx, self._real = _sfx(points.dtype)
        self._points = points
        n, d = points.shape
        rs, cs = _strides(points)
        err = C.c_int(0)
        self._h = getattr(lib(), f"orc_vptree_new_{self._sfx}")(
            _ptr(points, self._real) if n else None, n, d, rs, cs, C.byref(err))
        if err.value == 1:
            raise ArrayError("array is empty")
        if err.value == 2:
            raise ArrayError("array is not contiguous in memory")
        self.n, self.d = n, d

    @classmethod
    def euclidean(cls, points):
        return cls(points)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            getattr(lib(), f"orc_vptree_free_{self._sfx}")(h)
            self._h = None

    def query_nearest(self, q, count=False):
        q = np.ascontiguousarray(q, dtype=self._points.dtype)
        i, dd, nd = C.c_size_t(), self._real(), _ull(0)
        getattr(lib(), f"orc_vptree_query_nearest_{self._sfx}")(self._h, _ptr(q, self._real),
                                                               C.byref(i), C.byref(dd), C.byref(nd))
        if count:
            return int(i.value), self._points.dtype.type(dd.value), int(nd.value)
        return int(i.value), self._points.dtype.type(dd.value)

    def query_nearest_batch(self, Q, n_threads=1):
        Q = np.ascontiguousarray(Q, dtype=self._points.dtype)
        nq = Q.shape[0]
        oi = np.empty(nq, dtype=np.uintp)
        od = np.empty(nq, dtype=self._points.dtype)
        nd = getattr(lib(), f"orc_vptree_query_nearest_batch_{self._sfx}")(
            self._h, _ptr(Q, self._real), nq, Q.shape[1], _ptr(oi, C.c_size_t), _ptr(od, self._real), n_threads)
        return oi, od, int(nd)


def brute_knn(points, Q, k, n_threads=0):
    """(distance, index)-lexicographic exact k-NN; rows padded with (2^64-1, +inf) if k > n."""
    points = np.ascontiguousarray(points)
    Q = np.ascontiguousarray(Q, dtype=points.dtype)
    sfx, real = _sfx(points.dtype)
    n, d = points.shape
    nq = Q.shape[0]
    oi = np.empty((nq, k), dtype=np.uintp)
    od = np.empty((nq, k), dtype=points.dtype)
    if k:
        getattr(lib(), f"orc_brute_knn_{sfx}")(_ptr(points, real), n, d, d, _ptr(Q, real), nq, d, k,
                                               _ptr(oi, C.c_size_t), _ptr(od, real), n_threads)
    return oi, od


def brute_radius(points, Q, r, n_threads=0):
    """CSR (offsets, indices) of points with distance < r, ascending by index per query."""
    points = np.ascontiguousarray(points)
    Q = np.ascontiguousarray(Q, dtype=points.dtype)
    sfx, real = _sfx(points.dtype)
    n, d = points.shape
    nq = Q.shape[0]
    counts = np.zeros(nq, dtype=np.uintp)
    f = getattr(lib(), f"orc_brute_radius_{sfx}")
    f(_ptr(points, real), n, d, d, _ptr(Q, real), nq, d, r, _ptr(counts, C.c_size_t), None, None, n_threads)
    offsets = np.zeros(nq + 1, dtype=np.uintp)
    np.cumsum(counts, out=offsets[1:])
    out = np.empty(int(offsets[-1]), dtype=np.uintp)
    f(_ptr(points, real), n, d, d, _ptr(Q, real), nq, d, r, None, _ptr(offsets, C.c_size_t),
      _ptr(out, C.c_size_t), n_threads)
    return offsets, out
