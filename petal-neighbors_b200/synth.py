"""Deterministic synthetic inputs for tests and benches.

Counter-based: element (row, col) of a stream depends only on (seed, row * d + col), so any slice
can be generated independently (per shard, per rank) and identically on any host.
  hash   : splitmix64 finaliser of  counter + (seed + 1) * 0x9E3779B97F4A7C15  (mod 2^64)
  f32    : top 24 bits * 2^-24 ;  f64 : top 53 bits * 2^-53      (uniform in [0, 1))
  normal : Irwin-Hall(12) - 6 built from the twelve 16-bit lanes of three hashes (integer sum,
           so bit-reproducible without libm)
"""
from __future__ import annotations

import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z):
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _hash(seed: int, counters):
    with np.errstate(over="ignore"):
        return _mix(counters + np.uint64((seed + 1) & 0xFFFFFFFFFFFFFFFF) * _G)


def _to_unit(h, dtype):
    if np.dtype(dtype) == np.float32:
        return ((h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def uniform(n: int, d: int, seed: int, dtype=np.float32, row0: int = 0, chunk: int = 1 << 22):
    """U[0,1)^d points, rows row0 .. row0+n."""
    out = np.empty((n, d), dtype=dtype)
    flat = out.reshape(-1)
    total = n * d
    base = np.uint64(row0 * d)
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        c = np.arange(s, e, dtype=np.uint64) + base
        flat[s:e] = _to_unit(_hash(seed, c), dtype)
    return out


def _normal(seed: int, counters):
    """Irwin-Hall(12) - 6 from 12 sixteen-bit uniforms (three hashes per value)."""
    acc = np.zeros(counters.shape, dtype=np.int64)
    for t in range(3):
        with np.errstate(over="ignore"):
            h = _hash(seed, counters * np.uint64(3) + np.uint64(t))
        for lane in range(4):
            acc += ((h >> np.uint64(16 * lane)) & np.uint64(0xFFFF)).astype(np.int64)
    return (acc.astype(np.float64) + 6.0) / 65536.0 - 6.0


def gaussian_mixture(n: int, d: int, seed: int, n_centers: int = 1024, sigma: float = 0.05, center_seed: int = 4,
                     dtype=np.float32, clip: bool = False, row0: int = 0, chunk_rows: int = 1 << 16):
    """Mixture of `n_centers` isotropic Gaussians with centres U[0,1)^d (seed `center_seed`);
    component of a row = hash(row) mod n_centers."""
    centers = uniform(n_centers, d, center_seed, np.float64)
    out = np.empty((n, d), dtype=dtype)
    for s in range(0, n, chunk_rows):
        e = min(n, s + chunk_rows)
        rows = np.arange(row0 + s, row0 + e, dtype=np.uint64)
        comp = (_hash(seed ^ 0x5EED, rows) % np.uint64(n_centers)).astype(np.int64)
        c = (rows[:, None] * np.uint64(d) + np.arange(d, dtype=np.uint64)[None, :])
        z = _normal(seed, c)
        v = centers[comp] + sigma * z
        if clip:
            v = np.clip(v, 0.0, np.nextafter(1.0, 0.0))
        out[s:e] = v.astype(dtype)
    return out
