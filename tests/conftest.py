import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libsprintz_ref.so not built (needs /root/reference)")
    return Ref()


class Golden:
    def __init__(self, path):
        z = np.load(path)
        self.cases = []
        for k, line in enumerate(z["index"]):
            pred, esz, D, ret, rt, tag = str(line).split("|", 5)
            self.cases.append(dict(k=k, pred=int(pred), esz=int(esz), D=int(D), ret=int(ret), ref_rt=bool(int(rt)),
                                   tag=tag, x=z[f"x{k}"], stream=z[f"s{k}"]))


@pytest.fixture(scope="session")
def golden():
    return Golden(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
