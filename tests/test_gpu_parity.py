"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(libsprintz_b200.so); the oracle / compiled reference are only the checkers.

Bar: compressed bytes == reference bytes, return value == reference's, decode == input -- bit-exact."""
import numpy as np
import pytest

import inputs
import parity
from oracle import DELTA, XFF

pytestmark = pytest.mark.gpu

CODECS = [(DELTA, 1), (XFF, 1), (DELTA, 2), (XFF, 2)]
IDS = ["delta8", "xff8", "delta16", "xff16"]


def _check_case(spz, oracle, pred, esz, D, x, bad, tag):
    ws, wr = oracle.compress(pred, x, D)
    gs, gr = spz.compress(pred, x, D)
    msg = parity.compare_stream(gs, gr, ws, wr, esz)
    if msg is None and len(gs) != len(ws):
        msg = f"true byte length {len(gs)} != {len(ws)}"
    if msg is None:
        dec, dr = spz.decompress(pred, esz, ws)
        msg = parity.compare_decoded(dec, dr, x)
        if msg:
            msg = "decode: " + msg
    if msg:
        bad.append((D, x.size, tag, msg))


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()      # raises if the CUDA library is missing: no fallback
    return sprintz_b200


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_reference_matrix(spz, oracle, pred, esz):
    """The reference's own test matrix: 24 sizes x 18 input families, ndims across both dispatch sides
    (test/compress_testing.hpp:452-486, test/test_sprintz_8b.cpp:24-64)."""
    bad = []
    for D in (1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16, 17, 31, 32, 33, 63, 64, 65, 80, 129):
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, tag)
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_long_fuzz_and_sparse(spz, oracle, pred, esz):
    """1 Mi + 7 elements, fuzz and sparse families (compress_testing.hpp 'long fuzz' / 'long sparse fuzz')."""
    bad = []
    n = 1024 * 1024 + 7
    fams = {}
    fams.update(inputs.fuzz(n, esz))
    fams.update(inputs.sparse(n, esz))
    for D in (1, 3, 8, 80):
        for tag in ("fuzz256", "fuzz16", "fuzz2", "sparse6", "sparse2"):
            _check_case(spz, oracle, pred, esz, D, fams[tag], bad, tag)
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_adversarial(spz, oracle, pred, esz):
    """nbits edges, FIRE arithmetic edges (counter wrap, +-2^(w-1) deltas), runs inside FIRE (D4 trigger)."""
    bad = []
    for D in (1, 2, 3, 4, 5, 8, 16, 17, 33, 80, 130, 256):
        for nrows in (24, 100, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, f"{tag}/r{nrows}")
            x = inputs.random_walk_with_runs(nrows, D, esz, seed=9)[:-3]     # len not a multiple of D
            _check_case(spz, oracle, pred, esz, D, x, bad, "ragged")
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_rle_edges(spz, oracle, pred, esz):
    """All-zero / constant inputs with L = k*B + j for every j in [0, 2B]: the `<` vs `<=` asymmetry and
    abort-and-pad (sprintz_delta_rle.cpp:226 vs sprintz_xff_rle.cpp:362)."""
    udt = np.uint8 if esz == 1 else np.uint16
    bad = []
    for D in (1, 2, 3, 4, 5, 8):
        B = 8 * D
        for k in (2, 3, 4, 18):
            for j in range(0, 2 * B + 1):
                L = k * B + j
                _check_case(spz, oracle, pred, esz, D, np.zeros(L, udt), bad, "zeros")
                _check_case(spz, oracle, pred, esz, D, inputs.const_then_data(L, esz, (k - 1) * B), bad, "const+data")
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_long_runs(spz, oracle, pred, esz):
    """Runs of exactly 0x7f, 0x80, 0x7fff, 0x8000 and 2*0x7fff+1 blocks; a run that closes in slot 1."""
    udt = np.uint8 if esz == 1 else np.uint16
    bad = []
    for D in (1, 5):
        for nblocks in (0x7f, 0x80, 0x7fff, 0x8000, 2 * 0x7fff + 1):
            for j in (0, 1, 3):
                n = (nblocks + j) * 8 * D
                x = np.full(n, 9, udt)
                x2 = x.copy()
                x2[-5 * 8 * D:] = np.arange(5 * 8 * D) % 251
                _check_case(spz, oracle, pred, esz, D, x, bad, f"const{nblocks}+{j}")
                _check_case(spz, oracle, pred, esz, D, x2, bad, f"const{nblocks}+{j}+data")
    assert not bad, (len(bad), bad[:5])


def test_golden_vectors(spz, golden):
    """The committed fixtures were produced by the compiled reference itself (tests/golden/make_golden.py)."""
    bad = []
    for c in golden.cases:
        gs, gr = spz.compress(c["pred"], c["x"], c["D"])
        msg = parity.compare_stream(gs, gr, c["stream"], c["ret"], c["esz"])
        if msg is None:
            # decode the REFERENCE's bytes; for 16-bit the reference may have dropped the last byte from its count
            dec, dr = spz.decompress(c["pred"], c["esz"], gs)
            msg = parity.compare_decoded(dec, dr, c["x"])
        if msg:
            bad.append((c["pred"], c["esz"], c["D"], c["tag"], msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_live_reference_bytes_and_decode(spz, ref, pred, esz):
    """Against the unmodified reference running on this box's CPU: same bytes, and our decoder agrees with the
    reference's decoder wherever the reference round-trips (it does not on the D4 defect)."""
    bad = []
    for D in (1, 2, 3, 5, 8, 16, 80):
        for seed in range(3):
            x = inputs.random_walk_with_runs(700, D, esz, seed=100 + seed)
            rs, rr = ref.compress(pred, x, D)
            gs, gr = spz.compress(pred, x, D)
            msg = parity.compare_stream(gs, gr, rs, rr, esz)
            if msg is None:
                rd, rn = ref.decompress(pred, esz, gs, x.size)
                gd, gn = spz.decompress(pred, esz, gs)
                ref_ok = rn == x.size and np.array_equal(rd, x)
                if ref_ok and not (gn == rn and np.array_equal(gd, rd)):
                    msg = "decoders disagree on a stream the reference round-trips"
                if gn != x.size or not np.array_equal(gd, x):
                    msg = "our decode != input"
            if msg:
                bad.append((D, seed, msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_write_size_false_and_raw_decoder(spz, oracle, pred, esz):
    """write_size=false omits the 8 metadata bytes; the 5-argument decoders take the triple
    (sprintz_delta.h:52-90, sprintz_xff.h:47-83)."""
    for D in (1, 3, 8, 33):
        for n in (50, 5000):
            x = inputs.random_walk_with_runs(n, D, esz, seed=n)
            full, rf = spz.compress(pred, x, D, True)
            bare, rb = spz.compress(pred, x, D, False)
            wb, wrb = oracle.compress(pred, x, D, write_size=False)
            assert np.array_equal(bare, wb) and rb == wrb
            assert np.array_equal(full[8:], bare)
            ngroups = int(full[:4].view(np.uint32)[0]); rem = int(full[4:6].view(np.uint16)[0])
            d, m = spz.decompress_raw(pred, esz, bare, D, ngroups, rem)
            assert m == x.size and np.array_equal(d, x)


def test_error_conventions(spz):
    # ndims == 0 -> -1 like the reference (sprintz.cpp:71)
    s, r = spz.compress(DELTA, np.zeros(200, np.uint8), 0)
    assert r == -1
    with pytest.raises(spz.SprintzError):
        spz.compress(DELTA, np.zeros(4096 * 2, np.uint8), 2000)       # > SPZ_MAX_NDIMS
    with pytest.raises(spz.SprintzError):
        spz.compress(DELTA, np.zeros(16, np.float32), 2)


def test_reference_named_entry_points(spz, oracle):
    x8 = inputs.ar1(300, 8, 1)
    x16 = inputs.ar1(300, 16, 2)
    for name, pred, x, D in (("delta_8b", DELTA, x8, 8), ("xff_8b", XFF, x8, 8),
                             ("delta_16b", DELTA, x16, 16), ("xff_16b", XFF, x16, 16)):
        comp = getattr(spz, "sprintz_compress_" + name)
        decomp = getattr(spz, "sprintz_decompress_" + name)
        s, r = comp(x, D)
        ws, wr = oracle.compress(pred, x, D)
        assert r == wr and np.array_equal(s, ws)
        d, n = decomp(s)
        assert n == x.size and np.array_equal(d, x)


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_exact_cover_team_shapes(spz, oracle, pred, esz):
    """ndims that the team kernels cover exactly with several CONSECUTIVE dims per lane (24, 40, 48, 64, 80, 96, 128):
    contiguous header fields / payload segments per lane (spz_decode_hd.cu CONTIG path)."""
    bad = []
    for D in (24, 40, 48, 64, 80, 96, 128):
        for nrows in (40, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, f"{tag}/r{nrows}")
        x = inputs.random_walk_with_runs(3000, D, esz, seed=D)
        _check_case(spz, oracle, pred, esz, D, x, bad, "walk+runs")
        _check_case(spz, oracle, pred, esz, D, inputs.ar1(2048, D, esz, seed=D), bad, "ar1")
    assert not bad, (len(bad), bad[:5])
