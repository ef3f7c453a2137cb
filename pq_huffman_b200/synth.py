"""Seeded synthetic datasets of the shapes BASELINE.json names.

There is no network, so SIFT1M / GIST1M / Deep10M themselves are not available;
these generators reproduce their shape, value range and clustered structure so
that neighbours are meaningful.  Every generator is a pure function of
(n, d, seed) and is chunked so 10M-row sets do not need a second copy in RAM.
"""
import numpy as np

SEED_BASE = 20260101


def _centres(rng, n_centres, d, lo, hi):
    return rng.uniform(lo, hi, size=(n_centres, d)).astype(np.float32)


def _fill_clustered(out, rng, centres, sigma, chunk=1 << 18):
    n = out.shape[0]
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        idx = rng.integers(0, centres.shape[0], size=e - s)
        out[s:e] = centres[idx] + rng.normal(0.0, sigma, size=(e - s, out.shape[1])).astype(
            np.float32)


def sift_like(n, d=128, seed=SEED_BASE + 1, n_centres=1024):
    """uint8-valued integers in [0,255] stored as float32, mixture of Gaussians."""
    rng = np.random.default_rng(seed)
    centres = _centres(rng, n_centres, d, 10.0, 120.0)
    out = np.empty((n, d), dtype=np.float32)
    _fill_clustered(out, rng, centres, 25.0)
    np.rint(out, out=out)
    np.clip(out, 0.0, 255.0, out=out)
    return out


def gist_like(n, d=960, seed=SEED_BASE + 3, n_centres=1024):
    """float32 in [0,1], clustered (GIST descriptors are small non-negative reals)."""
    rng = np.random.default_rng(seed)
    centres = _centres(rng, n_centres, d, 0.02, 0.25)
    out = np.empty((n, d), dtype=np.float32)
    _fill_clustered(out, rng, centres, 0.04)
    np.clip(out, 0.0, 1.0, out=out)
    return out


def deep_like(n, d=96, seed=SEED_BASE + 4, n_centres=2048):
    """clustered Gaussian, L2-normalised rows (Deep1B descriptors are unit norm)."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, 1.0, size=(n_centres, d)).astype(np.float32)
    out = np.empty((n, d), dtype=np.float32)
    _fill_clustered(out, rng, centres, 0.45)
    norms = np.sqrt(np.einsum("ij,ij->i", out, out, dtype=np.float64)).astype(np.float32)
    out /= np.maximum(norms, 1e-20)[:, None]
    return out


def uniform(n, d, seed=SEED_BASE, lo=-1.0, hi=1.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, size=(n, d)).astype(np.float32)


def with_duplicates_and_ties(x, seed=SEED_BASE + 9, n_dups=8, n_ties=8):
    """Plant exact duplicate rows and exactly equidistant pairs (tie / position-0 tests).

    A tie pair is (r + e_j, r - e_j) around an integer-valued row r: both are at
    squared distance exactly 1 from r (the arithmetic is exact in FP32 for
    integer data).
    """
    rng = np.random.default_rng(seed)
    x = x.copy()
    n, d = x.shape
    rows = rng.choice(n, size=2 * n_dups + 3 * n_ties, replace=False)
    for i in range(n_dups):
        x[rows[2 * i + 1]] = x[rows[2 * i]]
    off = 2 * n_dups
    for i in range(n_ties):
        r, a, b = rows[off + 3 * i: off + 3 * i + 3]
        base = np.rint(x[r])
        base = np.clip(base, 1.0, 254.0)
        x[r] = base
        j = int(rng.integers(0, d))
        x[a] = base
        x[a, j] += 1.0
        x[b] = base
        x[b, j] -= 1.0
    return x


def pq_codes_stand_in(x, m=8, seed=SEED_BASE + 5):
    """Deterministic stand-in PQ codes uint8[n x m] (config 5).

    pq_encoder.c needs Yael's k-means and a time(NULL) seed (pq_encoder.c:270-272),
    so it can be neither built nor reproduced; bits/vector parity only needs the
    SAME codes on both sides.  Each sub-vector is projected on a fixed random
    direction and quantised to 256 levels by rank, which keeps neighbouring
    vectors' codes correlated (what the Huffman context model exploits).
    """
    rng = np.random.default_rng(seed)
    n, d = x.shape
    sub = d // m
    codes = np.empty((n, m), dtype=np.uint8)
    for j in range(m):
        w = rng.normal(size=sub).astype(np.float32)
        proj = x[:, j * sub:(j + 1) * sub] @ w
        order = np.argsort(proj, kind="stable")
        ranks = np.empty(n, dtype=np.int64)
        ranks[order] = np.arange(n)
        codes[:, j] = (ranks * 256 // n).astype(np.uint8)
    return codes
