"""Shared comparison helpers: what "bit-exact with the reference" means for one stream."""
import numpy as np


def compare_stream(got_stream, got_ret, want_stream, want_ret, esz):
    """want_* from the oracle (exact true length) or the reference (ret*esz bytes [+1 dropped byte]).
    Returns None if identical else a short description."""
    if got_ret != want_ret:
        return f"return value {got_ret} != {want_ret}"
    if want_ret < 0:
        return None
    n = want_ret * esz
    g = np.asarray(got_stream[:n]); w = np.asarray(want_stream[:n])
    if g.size != n or w.size != n:
        return f"stream shorter than ret*esz: got {g.size}, want {w.size}, need {n}"
    if not np.array_equal(g, w):
        i = int(np.nonzero(g != w)[0][0])
        return f"first differing byte at {i}/{n}: got {int(g[i]):#04x} want {int(w[i]):#04x}"
    # 16-bit: floor(bytes/2) can drop the last byte from the count; compare it when both have it
    if esz == 2 and len(got_stream) > n and len(want_stream) > n:
        if int(got_stream[n]) != int(want_stream[n]):
            return f"dropped odd byte differs: got {int(got_stream[n]):#04x} want {int(want_stream[n]):#04x}"
    if len(got_stream) not in (n, n + esz - 1):
        return f"true stream length {len(got_stream)} inconsistent with ret {want_ret}"
    return None


def compare_decoded(got, got_ret, x):
    if got_ret != x.size:
        return f"decoded count {got_ret} != {x.size}"
    g = np.asarray(got).reshape(-1).view(x.dtype)[: x.size]
    if not np.array_equal(g, x):
        i = int(np.nonzero(g != x)[0][0])
        return f"first differing element at {i}/{x.size}: got {int(g[i])} want {int(x[i])}"
    return None
