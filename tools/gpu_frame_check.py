"""Frame path check on the GPU: every chunk vs the oracle on the same slice, then decode == input."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
import parity  # noqa: E402
import sprintz_b200 as spz  # noqa: E402
from oracle import DELTA, XFF, Oracle  # noqa: E402


def check(pred, esz, D, nrows, chunk_rows, O, ragged=0, gen="ar1"):
    x = (inputs.ar1(nrows, D, esz) if gen == "ar1" else inputs.random_walk_with_runs(nrows, D, esz, seed=nrows))
    if ragged:
        x = x[:-ragged]
    xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
    fr = spz.frame_encode(pred, xt, D, chunk_rows)
    torch.cuda.synchronize()
    off, nb, po = fr.directory()
    h = fr.header()
    buf = fr.buf[: fr.nbytes()].cpu().numpy()
    nch = int(h.nchunks)
    clen = chunk_rows * D
    bad = 0
    first = None
    for c in range(nch):
        sl = x[c * clen: (c + 1) * clen]
        ws, wr = O.compress(pred, sl, D)
        gs = buf[po + off[c]: po + off[c] + nb[c]]
        msg = parity.compare_stream(gs, len(gs) // esz, ws, wr, esz)
        if msg:
            bad += 1
            first = first or f"chunk {c}: {msg}"
    out = spz.frame_decode(fr)
    torch.cuda.synchronize()
    dec = out.cpu().numpy().view(x.dtype)
    ok_dec = np.array_equal(dec, x)
    ratio = x.nbytes / max(1, int(h.payload_bytes))
    print(f"frame pred={pred} w={8*esz} D={D} rows={nrows} chunk_rows={chunk_rows} ragged={ragged} {gen}: "
          f"nchunks={nch} bad_chunks={bad} decode_ok={ok_dec} ratio={ratio:.3f} {first or ''}", flush=True)
    return bad == 0 and ok_dec


def main():
    O = Oracle()
    ok = True
    t0 = time.time()
    for esz, D in ((1, 8), (2, 16), (1, 80), (1, 1), (2, 2), (2, 3), (1, 5), (2, 33)):
        for pred in (DELTA, XFF):
            for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (4096, 20000, 0)):
                ok &= check(pred, esz, D, nrows, chunk_rows, O, ragged)
            ok &= check(pred, esz, D, 3000, 512, O, 0, gen="walk")
    print("frame check", "OK" if ok else "FAILED", f"{time.time() - t0:.1f}s")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
