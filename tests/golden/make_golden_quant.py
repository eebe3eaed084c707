"""Generates tests/golden/golden_quant_v1.npz from the REFERENCE's own quantiser.

The reference function is python/datasets/compress_bench.py:45-60 (_quantize).  Its module cannot be imported here (it
pulls matplotlib and the dataset loaders), so this script takes the function's source text -- unmodified -- out of the
reference file with `ast` and executes just that definition with numpy.  The function divides by
`max(1, np.max(mat, axis=0, keepdims=True))`, which Python only evaluates for a single column; multi-column fixtures are
therefore produced the way the expression reads (axis = 0): column by column, each column through the reference function.

usage (build container only; /root/reference does not exist on the GPU box): python tests/golden/make_golden_quant.py"""
import ast
import os
import warnings

import numpy as np

REF = "/root/reference/python/datasets/compress_bench.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_quantize():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_quantize")
    code = ast.get_source_segment(src, fn)
    ns = {"np": np}
    exec(compile(code, REF, "exec"), ns)
    return ns["_quantize"]


def main():
    q = reference_quantize()
    rng = np.random.default_rng(20260117)
    cases = {}
    k = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for fdt in (np.float32, np.float64):
            for odt in (np.uint8, np.uint16):
                for rows, D in ((1, 1), (2, 3), (17, 1), (260, 8), (513, 5), (40, 80)):
                    for fam in ("normal", "tiny_range", "constant", "negative_large", "ints", "mixed_scales"):
                        if fam == "normal":
                            x = rng.normal(0, 3, size=(rows, D))
                        elif fam == "tiny_range":                       # column max - min < 1: the divisor is 1
                            x = rng.uniform(5.0, 5.7, size=(rows, D))
                        elif fam == "constant":
                            x = np.full((rows, D), -2.5)
                        elif fam == "negative_large":
                            x = rng.uniform(-1e6, -1e3, size=(rows, D))
                        elif fam == "ints":
                            x = rng.integers(-500, 500, size=(rows, D)).astype(np.float64)
                        else:
                            x = rng.normal(0, 1, size=(rows, D)) * (10.0 ** rng.integers(-3, 6, size=(1, D)))
                        x = np.ascontiguousarray(x.astype(fdt))
                        out = np.empty((rows, D), odt)
                        for d in range(D):
                            out[:, d] = q(x[:, d:d + 1].copy(), odt)[:, 0]
                        cases[f"x{k}"] = x
                        cases[f"y{k}"] = out
                        k += 1
    cases["n"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "golden_quant_v1.npz"), **cases)
    print("wrote", k, "cases")


if __name__ == "__main__":
    main()
