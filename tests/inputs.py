"""Input families for the parity tests.

(1) the reference's own generators, restated (test/compress_testing.hpp:251-284 simple patterns,
    :295-303 squares, :347-370 fuzz with successive halving, :380-386 zeros, :409-423 sparse) and its
    size list (:452-463);
(2) the adversarial cases SURVEY.md section 4 asks for (RLE edges, nbits edges, FIRE arithmetic edges,
    the D4 trigger), which the reference's tests never try.

The reference seeds Eigen's setRandom() with srand(123); libc rand() is not reproducible from numpy, so
"fuzz" here is numpy's PCG64 with a fixed seed -- same distribution family, pinned by our golden files.
"""
from __future__ import annotations

import numpy as np

REF_SIZES = [1, 2, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 66, 71, 72, 73, 127, 128, 129, 135, 136,
             137, 4096, 4096 + 17]


def _dt(esz):
    return np.uint8 if esz == 1 else np.uint16


def simple_patterns(n, esz):
    i = np.arange(n, dtype=np.int64)
    odd = (i % 2) == 1
    fams = {
        "i%64": i % 64,
        "(i+64)%128": (i + 64) % 128,
        "(i+96)%256": (i + 96) % 256,
        "alt0": np.where(odd, (i + 64) % 128, 0),
        "alt72": np.where(odd, (i + 64) % 128, 72),
        "alt62": np.where(odd, (i + 64) % 128, 62 + (i + 1) % 4),
        "alt126": np.where(odd, (i + 64) % 128, 126 + (i + 1) % 4),
        "squares": (i % 16) * (i % 16) + ((i // 16) % 16),
    }
    return {k: v.astype(_dt(esz)) for k, v in fams.items()}


def zeros(n, esz):
    return {"zeros": np.zeros(n, _dt(esz))}


def fuzz(n, esz, seed=123):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << (8 * esz), size=n, dtype=np.uint32).astype(_dt(esz))
    shift = 0 if esz == 1 else 1
    out = {"fuzz256": raw.copy()}
    for name in ("fuzz128", "fuzz64", "fuzz32", "fuzz16"):
        raw = raw // (2 << shift)
        out[name] = raw.copy()
    raw = raw // (8 << (shift * 2))
    out["fuzz2"] = raw.copy()
    return out


def sparse(n, esz, seed=123):
    rng = np.random.default_rng(seed)
    orig = rng.integers(0, 1 << (8 * esz), size=n, dtype=np.uint32)
    sh = 8 * (esz - 1)
    return {
        "sparse56": (orig // (193 << sh)).astype(_dt(esz)),
        "sparse6": (orig // (250 << sh)).astype(_dt(esz)),
        "sparse2": (orig // (254 << sh)).astype(_dt(esz)),
    }


def reference_families(n, esz):
    fams = {}
    fams.update(simple_patterns(n, esz))
    fams.update(zeros(n, esz))
    fams.update(fuzz(n, esz))
    fams.update(sparse(n, esz))
    return fams


# ------------------------------------------------------------------ adversarial / structured

def ar1(nrows, ndims, esz, seed=12345, phi=0.98):
    """The benchmark's synthetic stream: AR(1) per dim, sigma 2 / 200, mid 128 / 32768."""
    rng = np.random.default_rng(seed)
    sigma, mid = (2.0, 128) if esz == 1 else (200.0, 32768)
    e = rng.normal(0.0, sigma, size=(nrows, ndims))
    s = np.empty_like(e)
    acc = np.zeros(ndims)
    for t in range(nrows):
        acc = phi * acc + e[t]
        s[t] = acc
    x = np.rint(mid + s).astype(np.int64) & ((1 << (8 * esz)) - 1)
    return x.astype(_dt(esz)).reshape(-1)


def random_walk_with_runs(nrows, ndims, esz, seed=7, p_const=0.3, step=3):
    """Piecewise: random-walk stretches alternating with constant stretches (exercises run slots)."""
    rng = np.random.default_rng(seed)
    x = np.zeros((nrows, ndims), np.int64)
    cur = rng.integers(0, 1 << (8 * esz), size=ndims)
    t = 0
    while t < nrows:
        seg = int(rng.integers(1, 80))
        const = rng.random() < p_const
        for _ in range(min(seg, nrows - t)):
            if not const:
                cur = cur + rng.integers(-step, step + 1, size=ndims)
            x[t] = cur
            t += 1
    return (x & ((1 << (8 * esz)) - 1)).astype(_dt(esz)).reshape(-1)


def ramps_then_const(nrows, ndims, esz, slope=5000, nramp=64):
    """SURVEY section 8a row D4 trigger: '+slope/step for nramp rows then -1/step'."""
    t = np.arange(nrows, dtype=np.int64)
    col = np.where(t < nramp, t * slope, nramp * slope - (t - nramp))
    x = np.repeat(col[:, None], ndims, axis=1) + np.arange(ndims)[None, :]
    return (x & ((1 << (8 * esz)) - 1)).astype(_dt(esz)).reshape(-1)


def linear_ramp(nrows, ndims, esz, slope=3):
    """Constant slope: FIRE learns coef -> errors become 0 -> FIRE run slots with nonzero prev_delta."""
    t = np.arange(nrows, dtype=np.int64)
    x = t[:, None] * (slope + np.arange(ndims)[None, :] % 3)
    return (x & ((1 << (8 * esz)) - 1)).astype(_dt(esz)).reshape(-1)


def const_then_data(n, esz, nconst, seed=3):
    """nconst constant elements (a long run) followed by noise, then constants again."""
    rng = np.random.default_rng(seed)
    x = np.full(n, 17, np.int64)
    if nconst < n:
        m = max(0, min(n - nconst, (n - nconst) // 2 + 1))
        x[nconst:nconst + m] = rng.integers(0, 1 << (8 * esz), size=m)
    return x.astype(_dt(esz))


def nbits_edge(nrows, ndims, esz, nbits, seed=5):
    """Columns whose zigzagged deltas need exactly `nbits` bits (alternating +/- 2^(nbits-2)-ish)."""
    rng = np.random.default_rng(seed)
    w = 8 * esz
    amp = (1 << max(0, nbits - 1)) - 1 if nbits > 0 else 0   # |delta| <= amp  ->  zigzag < 2^nbits
    x = np.zeros((nrows, ndims), np.int64)
    cur = np.full(ndims, 1 << (w - 1), np.int64)
    for t in range(nrows):
        if amp:
            d = rng.integers(-amp // 2 - 1, amp // 2 + 1, size=ndims)
            d[t % ndims] = amp // 2 if (t & 1) else -(amp // 2) - 1
            cur = cur + d
        x[t] = cur
    return (x & ((1 << w) - 1)).astype(_dt(esz)).reshape(-1)


def extreme_deltas(nrows, ndims, esz, seed=11):
    """Deltas of +-2^(w-1) and full-range jumps: grad_sum / counter wrap paths."""
    rng = np.random.default_rng(seed)
    w = 8 * esz
    choices = np.array([0, 1 << (w - 1), (1 << w) - 1, 1, (1 << (w - 1)) - 1, (1 << (w - 1)) + 1], np.int64)
    x = choices[rng.integers(0, len(choices), size=(nrows, ndims))]
    # long stretches where a column alternates min/max so the FIRE counter saturates/wraps
    x[:, 0] = np.where(np.arange(nrows) % 2 == 0, 0, (1 << w) - 1)
    if ndims > 1:
        x[:, 1] = (np.arange(nrows) * ((1 << (w - 1)) - 1)) & ((1 << w) - 1)
    return x.astype(_dt(esz)).reshape(-1)


def adversarial_families(nrows, ndims, esz):
    w = 8 * esz
    fams = {
        "ar1": ar1(nrows, ndims, esz),
        "walk_runs": random_walk_with_runs(nrows, ndims, esz),
        "d4_trigger": ramps_then_const(nrows, ndims, esz),
        "linear_ramp": linear_ramp(nrows, ndims, esz),
        "extreme": extreme_deltas(nrows, ndims, esz),
    }
    for nb in ({6, 7, 8} if esz == 1 else {7, 8, 9, 14, 15, 16}):
        fams[f"nbits{nb}"] = nbits_edge(nrows, ndims, esz, nb)
    assert all(v.dtype == _dt(esz) and v.size == nrows * ndims for v in fams.values()), w
    return fams
