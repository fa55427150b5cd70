"""TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Gradient fingerprints stored in tests/golden/*.npz (oracle/make_golden.py) and re-derived by the GPU parity tests.
A golden file cannot hold the 110 M gradient values of every case, so each parameter tensor is summarised by its
l2 norm, its sum, its first 8 elements and four projections onto fixed pseudo-random +-1 vectors.
"""
from __future__ import annotations

import zlib

import numpy as np


def projection_vectors(key: str, n: int, count: int = 4) -> np.ndarray:
    """`count` fixed +-1 vectors of length n for parameter `key` (seeded by the name, so a test rebuilds them).  For an
    error vector e the projection <e, r> has standard deviation |e|_2 over the choice of r: comparing the stored
    projections with those of the gradient under test bounds the FULL-tensor relative error |g - g_ref| / |g_ref|."""
    rng = np.random.Generator(np.random.PCG64(zlib.crc32(key.encode())))
    return rng.integers(0, 2, size=(count, n), dtype=np.int8) * np.int8(2) - np.int8(1)      # int8: 4 B per element of g


def project(key: str, grad) -> np.ndarray:
    """The `count` projections of a gradient tensor (float64 accumulation, chunked so the 23 M-element word-embedding
    table needs no float64 copy of the sign matrix)."""
    g = np.asarray(grad, np.float64).reshape(-1)
    r = projection_vectors(key, g.size)
    out = np.zeros(r.shape[0])
    step = 1 << 20
    for i in range(0, g.size, step):
        out += r[:, i:i + step].astype(np.float64) @ g[i:i + step]
    return out


def projection_error(key: str, grad: np.ndarray, stored_proj: np.ndarray, ref_norm: float) -> float:
    """Estimate of |g - g_ref|_2 / |g_ref|_2 from the stored projections: rms over the projections of
    <g, r_i> - <g_ref, r_i> (each has standard deviation |g - g_ref|_2)."""
    d = project(key, grad) - np.asarray(stored_proj, np.float64)
    return float(np.sqrt(np.mean(d * d)) / max(ref_norm, 1e-30))
