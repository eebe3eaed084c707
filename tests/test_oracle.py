"""CPU tests: the oracle (oracle/sprintz_oracle.c) against the committed golden vectors (generated from the
compiled reference) and, where oracle/_ref is present, live against the reference itself."""
import numpy as np
import pytest

import inputs
import parity
from oracle import DELTA, XFF


def test_oracle_matches_golden_bytes(oracle, golden):
    bad = []
    for c in golden.cases:
        s, r = oracle.compress(c["pred"], c["x"], c["D"])
        msg = parity.compare_stream(s, r, c["stream"], c["ret"], c["esz"])
        if msg:
            bad.append((c["pred"], c["esz"], c["D"], c["tag"], msg))
    assert not bad, bad[:5]


def test_oracle_decodes_golden_streams(oracle, golden):
    bad = []
    for c in golden.cases:
        # the reference's stream (ret*esz bytes + possibly dropped byte) must decode to the input
        d, n = oracle.decompress(c["pred"], c["esz"], c["stream"], c["x"].size)
        msg = parity.compare_decoded(d, n, c["x"])
        if msg:
            bad.append((c["pred"], c["esz"], c["D"], c["tag"], msg))
    assert not bad, bad[:5]


def test_golden_records_reference_decoder_defect(golden):
    # only 16-bit high-dim FIRE streams fail the reference's own round trip
    bad = [c for c in golden.cases if not c["ref_rt"]]
    assert bad, "fixture should contain the D4 trigger"
    assert all(c["pred"] == XFF and c["esz"] == 2 and c["D"] >= 3 for c in bad)


@pytest.mark.parametrize("esz", [1, 2])
@pytest.mark.parametrize("pred", [DELTA, XFF])
def test_oracle_vs_reference_live(oracle, ref, pred, esz):
    """Reference test matrix (test/compress_testing.hpp:452-486), subset of ndims to keep CPU time small."""
    bad = []
    for D in (1, 2, 3, 4, 5, 7, 8, 16, 17, 33, 80):
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                so, ro = oracle.compress(pred, x, D)
                sr, rr = ref.compress(pred, x, D)
                msg = parity.compare_stream(so, ro, sr, rr, esz)
                if msg:
                    bad.append((D, n, tag, msg))
    assert not bad, bad[:5]


@pytest.mark.parametrize("esz", [1, 2])
@pytest.mark.parametrize("pred", [DELTA, XFF])
def test_oracle_vs_reference_adversarial(oracle, ref, pred, esz):
    bad = []
    for D in (1, 2, 3, 5, 8, 16, 33, 130, 256):
        for nrows in (24, 100, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                so, ro = oracle.compress(pred, x, D)
                sr, rr = ref.compress(pred, x, D)
                msg = parity.compare_stream(so, ro, sr, rr, esz)
                if msg is None:
                    d, n = oracle.decompress(pred, esz, so, x.size)
                    msg = parity.compare_decoded(d, n, x)
                if msg:
                    bad.append((D, nrows, tag, msg))
    assert not bad, bad[:5]


@pytest.mark.parametrize("esz", [1, 2])
@pytest.mark.parametrize("pred", [DELTA, XFF])
def test_oracle_vs_reference_long_runs(oracle, ref, pred, esz):
    """Runs of exactly 0x7f, 0x80, 0x7fff, 0x8000, 2*0x7fff+1 blocks (the reference never tests > 1 Mi elements)."""
    udt = np.uint8 if esz == 1 else np.uint16
    bad = []
    for D in (1, 5):
        for nblocks in (0x7f, 0x80, 0x7fff, 0x8000, 2 * 0x7fff + 1):
            for j in (0, 1, 3):
                n = (nblocks + j) * 8 * D
                x = np.full(n, 9, udt)
                x2 = x.copy()
                x2[-5 * 8 * D:] = np.arange(5 * 8 * D) % 251
                for xx in (x, x2):
                    so, ro = oracle.compress(pred, xx, D)
                    sr, rr = ref.compress(pred, xx, D)
                    msg = parity.compare_stream(so, ro, sr, rr, esz)
                    if msg is None:
                        d, nn = oracle.decompress(pred, esz, so, xx.size)
                        msg = parity.compare_decoded(d, nn, xx)
                    if msg:
                        bad.append((D, nblocks, j, msg))
    assert not bad, bad[:5]


def test_oracle_write_size_false_roundtrip(oracle):
    for pred in (DELTA, XFF):
        for esz in (1, 2):
            for D in (1, 3, 8):
                x = inputs.random_walk_with_runs(300, D, esz)
                full, _ = oracle.compress(pred, x, D, write_size=True)
                bare, _ = oracle.compress(pred, x, D, write_size=False)
                assert np.array_equal(full[8:], bare)
                ngroups = int(full[:4].view(np.uint32)[0]); rem = int(full[4:6].view(np.uint16)[0])
                d, n = oracle.decompress_raw(pred, esz, bare, x.size, D, ngroups, rem)
                assert n == x.size and np.array_equal(d, x)


def test_oracle_rejects_zero_dims(oracle):
    s, r = oracle.compress(DELTA, np.zeros(200, np.uint8), 0)
    assert r == -1
