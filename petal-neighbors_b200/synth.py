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


# ---- the same streams generated with torch (on the GPU: 10M x 128 in well under a second instead of a minute of
# numpy), bit-identical to the numpy generators above (tests/test_host_side.py).  int64 arithmetic wraps mod 2^64. ----
def _s64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr_t(z, s: int):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix_t(z):
    z = (z ^ _lsr_t(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr_t(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr_t(z, 31)


def _hash_t(seed: int, counters):
    return _mix_t(counters + _s64(((seed + 1) & 0xFFFFFFFFFFFFFFFF) * 0x9E3779B97F4A7C15))


def _to_unit_t(h, dtype):
    import torch
    if dtype == torch.float32:
        return _lsr_t(h, 40).to(torch.float32) * (2.0 ** -24)
    return _lsr_t(h, 11).to(torch.float64) * (2.0 ** -53)


def uniform_torch(n: int, d: int, seed: int, dtype=None, row0: int = 0, device="cuda", chunk: int = 1 << 24, out=None):
    """uniform() as a torch tensor on `device`."""
    import torch
    dtype = dtype or torch.float32
    if out is None:
        out = torch.empty((n, d), dtype=dtype, device=device)
    flat = out.view(-1)
    total = n * d
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        c = torch.arange(s + row0 * d, e + row0 * d, dtype=torch.int64, device=out.device)
        flat[s:e] = _to_unit_t(_hash_t(seed, c), dtype)
    return out


def _umod_t(h, m: int):
    """h mod m for h an unsigned 64-bit value held in an int64 tensor."""
    return ((_lsr_t(h, 1) % m) * 2 + (h & 1)) % m


def gaussian_mixture_torch(n: int, d: int, seed: int, n_centers: int = 1024, sigma: float = 0.05, center_seed: int = 4,
                           dtype=None, clip: bool = False, row0: int = 0, device="cuda", chunk_rows: int = 1 << 17):
    """gaussian_mixture() as a torch tensor on `device`."""
    import torch
    dtype = dtype or torch.float32
    centers = uniform_torch(n_centers, d, center_seed, torch.float64, device=device)
    out = torch.empty((n, d), dtype=dtype, device=device)
    cols = torch.arange(d, dtype=torch.int64, device=device)[None, :]
    hi = float(np.nextafter(1.0, 0.0))
    for s in range(0, n, chunk_rows):
        e = min(n, s + chunk_rows)
        rows = torch.arange(row0 + s, row0 + e, dtype=torch.int64, device=device)
        comp = _umod_t(_hash_t(seed ^ 0x5EED, rows), n_centers)
        c = rows[:, None] * d + cols
        acc = torch.zeros(c.shape, dtype=torch.int64, device=device)
        for t in range(3):
            h = _hash_t(seed, c * 3 + t)
            for lane in range(4):
                acc += _lsr_t(h, 16 * lane) & 0xFFFF
        z = (acc.to(torch.float64) + 6.0) / 65536.0 - 6.0
        v = centers[comp] + sigma * z
        if clip:
            v = torch.clamp(v, 0.0, hi)
        out[s:e] = v.to(dtype)
    return out


def fast_uniform(n, d, seed, dtype=np.float32, row0=0):
    """uniform() as a numpy array, generated on the GPU when one is present (identical values either way)."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
            return uniform_torch(n, d, seed, tdt, row0=row0).cpu().numpy()
    except ImportError:
        pass
    return uniform(n, d, seed, dtype, row0=row0)


def fast_gaussian_mixture(n, d, seed, dtype=np.float32, **kw):
    """gaussian_mixture() as a numpy array, generated on the GPU when one is present (identical values either way)."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
            return gaussian_mixture_torch(n, d, seed, dtype=tdt, **kw).cpu().numpy()
    except ImportError:
        pass
    return gaussian_mixture(n, d, seed, dtype=dtype, **kw)
