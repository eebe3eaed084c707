"""Block-size axis of BASELINE config 5 on the CPU: the oracle parameterised on rows per block (oracle_compress_bs).

* block_rows == 8 is byte-identical to the compiled reference over the reference's own matrix (it IS the pinned oracle:
  the 8-row entry points call the same code with 8);
* 16 / 32 / 64 rows have no reference build (the reference's block size is a constant, sprintz_delta_rle.cpp:62-64):
  PARITY UNPINNED BEYOND BLOCK 8.  What can be checked is checked: round trip, the derived constants, the size formula."""
import numpy as np
import pytest

import inputs
from oracle import DELTA, XFF

CODECS = [(DELTA, 1), (XFF, 1), (DELTA, 2), (XFF, 2)]
IDS = ["delta8", "xff8", "delta16", "xff16"]


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_block8_is_the_reference(oracle, ref, pred, esz):
    bad = []
    for D in (1, 2, 3, 5, 8, 16, 17, 80):
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                got, gr = oracle.compress_bs(pred, x, D, 8)
                want, wr = ref.compress(pred, x, D)
                nref = max(wr, 0) * esz
                if gr != wr or not np.array_equal(got[:nref], np.asarray(want[:nref])):
                    bad.append((D, n, tag))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
@pytest.mark.parametrize("R", [16, 32, 64])
def test_larger_blocks_round_trip(oracle, pred, esz, R):
    bad = []
    for D in (1, 2, 3, 5, 8, 16, 33, 80):
        for nrows in (3, 2 * R - 1, 2 * R, 5 * R + 3, 40 * R):
            fams = inputs.adversarial_families(nrows, D, esz)
            fams["zeros"] = np.zeros(nrows * D, fams["ar1"].dtype)
            fams["const+data"] = inputs.const_then_data(nrows * D, esz, (nrows // 2) * D)
            fams["ragged"] = inputs.random_walk_with_runs(nrows, D, esz, seed=R)[: max(1, nrows * D - 3)]
            for tag, x in fams.items():
                s, r = oracle.compress_bs(pred, x, D, R)
                y, n = oracle.decompress_bs(pred, esz, s, x.size, R)
                if n != x.size or not np.array_equal(y, x):
                    bad.append((D, nrows, tag))
    assert not bad, (len(bad), bad[:5])


def test_derived_constants(oracle):
    """min_data_size = 8 * block_sz * group_sz_blocks; a group is two blocks of block_rows rows; one header per group."""
    for R in (8, 16, 32, 64):
        D = 8
        x = inputs.ar1(16 * R // D - 1, D, 1)                # one element row short of min_data_size -> stored raw
        x = x[: 16 * R - 1]
        s, _ = oracle.compress_bs(DELTA, x, D, R)
        assert len(s) == 8 + x.size and int.from_bytes(bytes(s[:4]), "little") == 0
        x = inputs.ar1(6 * R, D, 1, seed=R)                   # exactly three groups, no tail
        s, _ = oracle.compress_bs(DELTA, x, D, R)
        assert int.from_bytes(bytes(s[:4]), "little") == 3 and int.from_bytes(bytes(s[4:6]), "little") == 0
        # all-zero input: one run slot of (rows / R - 2 ..) blocks, exactly like block 8 with larger blocks
        z = np.zeros(64 * R * D, np.uint8)
        s8, _ = oracle.compress_bs(DELTA, z, D, R)
        y, n = oracle.decompress_bs(DELTA, 1, s8, z.size, R)
        assert n == z.size and not y.any() and len(s8) < 8 + 6 + 4 + 2 * R * D + 8
