"""Broad GPU-vs-oracle sweep that does NOT stop at the first failure: prints a table of failures by
(codec, width, ndims) so one gpurun call localises bugs.  Test infrastructure (uses oracle/)."""
import collections
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
import parity  # noqa: E402
import sprintz_b200 as spz  # noqa: E402
from oracle import DELTA, XFF, Oracle  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    O = Oracle()
    fails = collections.OrderedDict()
    nfail = collections.Counter(); ncase = collections.Counter()
    t0 = time.time()

    def run(pred, esz, D, tag, x):
        key = (("xff" if pred else "delta"), 8 * esz, D)
        ncase[key] += 1
        try:
            ws, wr = O.compress(pred, x, D)
            gs, gr = spz.compress(pred, x, D)
            msg = parity.compare_stream(gs, gr, ws, wr, esz)
            if msg is None:
                dec, dr = spz.decompress(pred, esz, ws)
                msg = parity.compare_decoded(dec, dr, x)
                if msg: msg = "DEC " + msg
            else:
                msg = "ENC " + msg
        except Exception as e:  # noqa: BLE001
            msg = f"EXC {type(e).__name__}: {e}"
        if msg:
            nfail[key] += 1
            fails.setdefault(key, [])
            if len(fails[key]) < 4:
                fails[key].append(f"{tag} n={x.size}: {msg}")

    dims = [1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 80, 129] if not quick else [1, 3, 5, 8, 16, 33, 80]
    sizes = inputs.REF_SIZES if not quick else [1, 127, 128, 137, 4113]
    for esz in (1, 2):
        for pred in (DELTA, XFF):
            for D in dims:
                for n in sizes:
                    for tag, x in inputs.reference_families(n, esz).items():
                        run(pred, esz, D, tag, x)
                for nrows in (24, 200):
                    for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                        run(pred, esz, D, tag + f"/r{nrows}", x)
    total_fail = sum(nfail.values())
    print(f"sweep: {sum(ncase.values())} cases, {total_fail} failures, {time.time() - t0:.1f}s")
    for key, msgs in fails.items():
        print(f"  {key}: {nfail[key]}/{ncase[key]} failed")
        for m in msgs:
            print("      " + m)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump({"cases": sum(ncase.values()), "failures": total_fail,
                   "by_key": {str(k): v for k, v in fails.items()}}, f, indent=1)
    return 1 if total_fail else 0


if __name__ == "__main__":
    sys.exit(main())
