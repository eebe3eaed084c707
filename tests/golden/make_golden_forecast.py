"""Generates tests/golden/golden_forecast_v1.npz from the UNMODIFIED reference (oracle/_ref/libsprintz_ref.so): the
stand-alone forecaster transforms encode_{delta,doubledelta}_rowmajor_{8b,16b} (delta.h:17-67; transform 1 / 2) and
encode_xff_rowmajor_{8b,16b} (predict.h:15-30; transform 3).  Run in the build container only:

    python tests/golden/make_golden_forecast.py

Each case stores the input and the reference's output bytes (6-byte metadata included) and return value.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
from oracle import Ref  # noqa: E402


def cases():
    for esz in (1, 2):
        for tr in (1, 2, 3):
            for D in (1, 2, 3, 5, 8, 16, 33, 80):
                for n in (1, 7, 127, 128, 137, 1024 + 17):
                    fams = inputs.reference_families(n, esz)
                    for tag in ("i%64", "squares", "fuzz256", "sparse6"):
                        yield tr, esz, D, f"{tag}/n{n}", fams[tag]
                yield tr, esz, D, "walk/r300", inputs.random_walk_with_runs(300, D, esz, seed=D)[:-1]


def main():
    R = Ref()
    out, index = {}, []
    for k, (tr, esz, D, tag, x) in enumerate(cases()):
        y, ret = R.encode_xff_forecast(x, D) if tr == 3 else R.encode_forecast(tr, x, D)
        out[f"x{k}"] = x
        out[f"y{k}"] = y
        index.append(f"{tr}|{esz}|{D}|{ret}|{tag}")
    out["index"] = np.array(index)
    path = os.path.join(ROOT, "tests", "golden", "golden_forecast_v1.npz")
    np.savez_compressed(path, **out)
    print(len(index), "cases ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
