"""Generates tests/golden/golden_rowmajor_v1.npz from the UNMODIFIED reference (oracle/_ref/libsprintz_ref.so):
the non-RLE sibling codecs compress_rowmajor_delta_{8b,16b} (kind 1), compress_rowmajor_{8b,16b} (kind 2),
sprintz_delta.h:27-44, and compress8b_rowmajor_xff (kind 3), sprintz_xff.h:28-31.  Run in the build container only:

    python tests/golden/make_golden_rowmajor.py

Each case stores the input, the reference's compressed bytes (first ret*esz bytes + the byte a 16-bit return value
may drop) and the reference's return value.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
from oracle import RM_DELTA, RM_PLAIN, RM_XFF8, Ref  # noqa: E402


def cases():
    for esz in (1, 2):
        for kind in (RM_DELTA, RM_PLAIN) + ((RM_XFF8,) if esz == 1 else ()):
            for D in (1, 3, 5, 8, 16, 33):
                for n in (1, 127, 128, 137, 1024 + 17):
                    fams = inputs.reference_families(n, esz)
                    for tag in ("i%64", "squares", "zeros", "fuzz256", "fuzz2", "sparse6"):
                        yield kind, esz, D, f"{tag}/n{n}", fams[tag]
            for D in (2, 8, 16, 80):
                for tag, x in inputs.adversarial_families(200, D, esz).items():
                    yield kind, esz, D, f"{tag}/r200", x


def main():
    R = Ref()
    out, index = {}, []
    for k, (kind, esz, D, tag, x) in enumerate(cases()):
        stream, ret = R.compress_rowmajor(kind, x, D)
        dec, nret = R.decompress_rowmajor(kind, esz, stream, x.size)
        assert nret == x.size and np.array_equal(dec, x), (kind, esz, D, tag)
        out[f"x{k}"] = x
        out[f"s{k}"] = stream
        index.append(f"{kind}|{esz}|{D}|{ret}|{tag}")
    out["index"] = np.array(index)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_rowmajor_v1.npz")
    np.savez_compressed(path, **out)
    print(len(index), "cases ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
