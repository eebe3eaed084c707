"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref/libsprintz_ref.so,
compiled from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each case stores the input, the reference's compressed bytes (first ret*esz bytes + the one byte the
16-bit return value may drop), the reference's return value, and whether the reference's own
decoder round-trips it (it does not for the D4 defect).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
from oracle import DELTA, XFF, Ref  # noqa: E402


def cases():
    for esz in (1, 2):
        for pred in (DELTA, XFF):
            # reference families at a few sizes / dims (both sides of the low/high-dim dispatch)
            dims = [1, 2, 3, 4, 5, 8, 16, 33] if esz == 1 else [1, 2, 3, 5, 8, 16, 33]
            for D in dims:
                for n in (1, 127, 128, 137, 1024 + 17):
                    fams = inputs.reference_families(n, esz)
                    for tag in ("i%64", "alt126", "squares", "zeros", "fuzz256", "fuzz16", "fuzz2", "sparse6"):
                        yield pred, esz, D, f"{tag}/n{n}", fams[tag]
            # structured / adversarial
            for D in (1, 2, 3, 5, 8, 16, 80):
                for tag, x in inputs.adversarial_families(200, D, esz).items():
                    yield pred, esz, D, f"{tag}/r200", x
            # RLE edges around the abort-and-pad logic
            for D in (1, 3, 5, 8):
                B = 8 * D
                for L in (2 * B, 2 * B + 1, 3 * B, 3 * B + B // 2, 4 * B, 5 * B + 1, 18 * B + 3):
                    udt = np.uint8 if esz == 1 else np.uint16
                    yield pred, esz, D, f"zeros/L{L}", np.zeros(L, udt)
                    yield pred, esz, D, f"constdata/L{L}", inputs.const_then_data(L, esz, max(0, L - 2 * B))
            # a run longer than 0x7f blocks (two length bytes)
            yield pred, esz, 5, "const/0x90blocks", np.full(0x90 * 40 + 7, 3, np.uint8 if esz == 1 else np.uint16)


def main():
    R = Ref()
    out = {}
    index = []
    for k, (pred, esz, D, tag, x) in enumerate(cases()):
        stream, ret = R.compress(pred, x, D)
        dec, nret = R.decompress(pred, esz, stream, x.size)
        ref_rt = bool(nret == x.size and np.array_equal(dec, x))
        out[f"x{k}"] = x
        out[f"s{k}"] = stream
        index.append((pred, esz, D, ret, int(ref_rt), tag))
    out["index"] = np.array([f"{p}|{e}|{d}|{r}|{rt}|{t}" for p, e, d, r, rt, t in index])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    nbad = sum(1 for i in index if not i[4])
    print(f"wrote {path}: {len(index)} cases, {os.path.getsize(path) / 1e6:.2f} MB, "
          f"{nbad} cases where the reference decoder does not round-trip")


if __name__ == "__main__":
    main()
