"""Entropy stage on the CPU (oracle/huff_oracle.c).  PARITY UNPINNED: Huff0 is not in the reference tree
(cpp/Compress/entropy.cpp:46-58 is a stub), so this is the specification of OUR container; the tests check what can be
checked without a reference: prefix-code validity, the 11-bit limit, optimality on small alphabets, round trips, and that
malformed containers are rejected."""
import numpy as np
import pytest


def _families():
    rng = np.random.default_rng(7)
    fib = [1, 1]
    while len(fib) < 32:
        fib.append(fib[-1] + fib[-2])
    deep = np.concatenate([np.full(c, i, np.uint8) for i, c in enumerate(fib[:28])])
    rng.shuffle(deep)
    return {
        "empty": np.zeros(0, np.uint8),
        "one_byte": np.array([200], np.uint8),
        "one_symbol": np.full(5000, 7, np.uint8),
        "two_symbols": rng.integers(0, 2, 10000).astype(np.uint8),
        "uniform": rng.integers(0, 256, 70001).astype(np.uint8),
        "geometric": np.minimum(rng.geometric(0.3, 150000) - 1, 255).astype(np.uint8),
        "fibonacci (length limit)": deep,
        "segment edge 4096": rng.integers(0, 9, 4096).astype(np.uint8),
        "segment edge 4097": rng.integers(0, 9, 4097).astype(np.uint8),
        "ragged 16": rng.integers(0, 40, 8192 + 15).astype(np.uint8),
    }


def test_round_trip_and_container(oracle):
    for name, x in _families().items():
        c = oracle.huff_encode(x)
        y, r = oracle.huff_decode(c, x.size)
        assert r == x.size and np.array_equal(y, x), name
        assert c[:4].tobytes() == b"SPZH" and int.from_bytes(c[16:24].tobytes(), "little") == x.size, name
        assert int.from_bytes(c[48:56].tobytes(), "little") == len(c), name


def test_lengths_are_a_limited_prefix_code(oracle):
    rng = np.random.default_rng(3)
    for trial in range(300):
        k = int(rng.integers(1, 257))
        cnt = np.zeros(256, np.uint64)
        syms = rng.choice(256, size=k, replace=False)
        scale = [4, 1000, 10 ** 9][trial % 3]
        cnt[syms] = rng.integers(1, scale, size=k).astype(np.uint64) ** (1 + trial % 2)
        ln = oracle.huff_lengths(cnt)
        assert ((ln > 0) == (cnt > 0)).all()
        assert ln.max() <= 11
        kraft = sum(2.0 ** -int(v) for v in ln if v)
        assert kraft <= 1.0 + 1e-12 and (k == 1 or kraft > 0.5)


def test_small_alphabets_are_optimal(oracle):
    # counts 1,1,2,4 -> lengths 3,3,2,1; equal counts over 4 symbols -> all 2
    cnt = np.zeros(256, np.uint64); cnt[[10, 20, 30, 40]] = [1, 1, 2, 4]
    assert list(oracle.huff_lengths(cnt)[[10, 20, 30, 40]]) == [3, 3, 2, 1]
    cnt = np.zeros(256, np.uint64); cnt[[1, 2, 3, 4]] = 5
    assert list(oracle.huff_lengths(cnt)[[1, 2, 3, 4]]) == [2, 2, 2, 2]
    cnt = np.full(256, 3, np.uint64)
    assert (oracle.huff_lengths(cnt) == 8).all()


def test_malformed_containers_are_rejected(oracle):
    x = _families()["geometric"]
    c = oracle.huff_encode(x)
    assert oracle.huff_decode(c[:100], x.size)[1] == -1                 # truncated
    bad = c.copy(); bad[0] ^= 1
    assert oracle.huff_decode(bad, x.size)[1] == -1                     # magic
    bad = c.copy(); bad[64:192] = 0x11
    assert oracle.huff_decode(bad, x.size)[1] == -1                     # 256 codes of length 1: not a prefix code
    bad = c.copy(); bad[192 + 8] = 0xff; bad[192 + 14] = 0x7f
    assert oracle.huff_decode(bad, x.size)[1] == -1                     # directory entry outside the container
    assert oracle.huff_decode(c, x.size + 1)[1] == -1                   # wrong raw size
