"""Synthetic workload of the benchmark: a row-major multivariate AR(1) stream.

    s[t] = phi * s[t-1] + N(0, sigma),   x[t] = round(mid + s[t]) wrapped to w bits
    phi = 0.98; int8: sigma 2, mid 128; int16: sigma 200, mid 32768

Generated with torch ops on whatever device is asked for (a blocked scan, so it is vectorised), slab by slab
so a 4 GiB stream never needs more than a slab of float64 temporaries.  Workload generation only -- not
part of the codec path.
"""
from __future__ import annotations

import torch

PHI = 0.98
_T = 128          # block length of the scan; phi**-128 = 13.3, safe in float64


def _ar1_slab(e: torch.Tensor, carry: torch.Tensor, phi: float) -> tuple[torch.Tensor, torch.Tensor]:
    """e: (rows, D) float64 noise, rows % _T == 0; carry: (D,) state before the slab. Returns (s, new carry)."""
    rows, D = e.shape
    nb = rows // _T
    eb = e.view(nb, _T, D)
    t = torch.arange(_T, device=e.device, dtype=torch.float64)
    up = (phi ** t).view(1, _T, 1)           # phi^t
    dn = (phi ** (-t)).view(1, _T, 1)        # phi^-t
    local = torch.cumsum(eb * dn, dim=1) * up          # zero-state response inside each block
    ends = local[:, -1, :]                              # (nb, D)
    # state at the end of each block: c[b] = ends[b] + phi^T c[b-1]; phi^T = 0.075 -> 10 terms are exact to 1e-11
    pT = phi ** _T
    c = ends.clone()
    shifted = ends
    coef = 1.0
    for _ in range(10):
        coef *= pT
        shifted = torch.cat([torch.zeros(1, D, device=e.device, dtype=torch.float64), shifted[:-1]], dim=0)
        c = c + coef * shifted
    # contribution of the carry entering the slab
    bidx = torch.arange(1, nb + 1, device=e.device, dtype=torch.float64).view(nb, 1)
    c = c + (pT ** bidx) * carry.view(1, D)
    prev = torch.cat([carry.view(1, D), c[:-1]], dim=0)   # state entering each block
    s = local + (phi * up) * prev.view(nb, 1, D)
    return s.view(rows, D), c[-1].clone()


def ar1_stream(nrows: int, ndims: int, elem_sz: int, seed: int = 12345, device="cuda",
               slab_rows: int = 1 << 20) -> torch.Tensor:
    """-> 1-D tensor of nrows*ndims elements (uint8, or int16 holding the uint16 bit pattern), row-major."""
    sigma, mid = (2.0, 128.0) if elem_sz == 1 else (200.0, 32768.0)
    dt = torch.uint8 if elem_sz == 1 else torch.int16
    out = torch.empty(nrows * ndims, dtype=dt, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    carry = torch.zeros(ndims, dtype=torch.float64, device=device)
    slab_rows = max(_T, (slab_rows // _T) * _T)
    # keep slabs around 16M elements regardless of ndims
    slab_rows = max(_T, min(slab_rows, ((1 << 24) // max(1, ndims)) // _T * _T))
    r0 = 0
    while r0 < nrows:
        rows = min(slab_rows, nrows - r0)
        padded = (rows + _T - 1) // _T * _T
        e = torch.randn(padded, ndims, generator=gen, device=device, dtype=torch.float64) * sigma
        s, carry_full = _ar1_slab(e, carry, PHI)
        if padded != rows:
            carry = s[rows - 1].clone()
        else:
            carry = carry_full
        x = torch.round(mid + s[:rows]).to(torch.int64)
        x = x & ((1 << (8 * elem_sz)) - 1)
        if elem_sz == 2:
            x = torch.where(x >= 32768, x - 65536, x)
        out[r0 * ndims: (r0 + rows) * ndims] = x.to(dt).reshape(-1)
        r0 += rows
    return out
