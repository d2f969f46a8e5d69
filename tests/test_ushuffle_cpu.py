"""The native k-let shuffle (bpb_ushuffle_ohe, host code) against the reference's own
libushuffle.c compiled into oracle/_ref, against libc's random(), and against committed vectors."""
import ctypes
import os

import numpy as np
import pytest

from bpreveal_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libushuffle.so")
GOLD = os.path.join(ROOT, "tests", "golden", "ushuffle_vectors.npz")


def one_hot(seq_idx, alphabet=4):
    return np.eye(alphabet, dtype=np.uint8)[seq_idx]


def ref_shuffle(x, k, n, seed):
    lib = ctypes.CDLL(REF_SO)
    lib.initialize()
    lib.seedRng(ctypes.c_uint(seed))
    x = np.ascontiguousarray(x, dtype=np.int8)
    out = np.empty((n,) + x.shape, dtype=np.int8)
    lib.shuffleOhe(x.ctypes.data_as(ctypes.c_char_p), out.ctypes.data_as(ctypes.c_char_p),
                   ctypes.c_int(x.shape[1]), ctypes.c_int(x.shape[0]), ctypes.c_int(k), ctypes.c_int(n))
    return out.astype(np.uint8)


def klet_counts(x, k):
    codes = (x * (1 << np.arange(x.shape[1]))).sum(axis=1)
    lets = [tuple(codes[i:i + k]) for i in range(len(codes) - k + 1)]
    return sorted(lets)


CASES = [(300, 2, 5, 355687), (3090, 3, 20, 355687), (64, 4, 3, 7), (50, 1, 4, 11), (5, 7, 2, 3),
         (200, 2, 3, 0)]


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("L,k,n,seed", CASES)
def test_bit_identical_to_the_reference_library(L, k, n, seed):
    rng = np.random.default_rng(L + k)
    x = one_hot(rng.integers(0, 4, L))
    if L == 300:
        x[10] = [1, 0, 1, 0]     # multi-hot and all-zero rows are legal input (ushuffle.py:50-58)
        x[20] = 0
    assert np.array_equal(ops.ushuffle_ohe(x, k, n, seed), ref_shuffle(x, k, n, seed))


@pytest.mark.parametrize("L,k,n,seed", CASES)
def test_klets_are_preserved(L, k, n, seed):
    rng = np.random.default_rng(L + k)
    x = one_hot(rng.integers(0, 4, L))
    out = ops.ushuffle_ohe(x, k, n, seed)
    assert out.shape == (n, L, 4)
    for s in out:
        assert klet_counts(s, min(k, L)) == klet_counts(x, min(k, L))
        if 1 < k < L:
            assert np.array_equal(s[:k - 1], x[:k - 1]) and np.array_equal(s[-1], x[-1])
    if L >= 200:
        assert not np.array_equal(out[0], x) and not np.array_equal(out[0], out[1])


def test_generator_matches_libc_random():
    """k = 1 is a Fisher-Yates walk driven by random() % (i + 1): replay it with libc's generator."""
    libc = ctypes.CDLL(None)
    libc.random.restype = ctypes.c_long
    for seed in (1, 355687, 0, 2**31 + 5):
        L = 97
        x = one_hot(np.arange(L) % 4)
        x[:, 0] |= (np.arange(L) % 3 == 0).astype(np.uint8)  # make rows distinguishable enough
        libc.srandom(ctypes.c_uint(seed))
        order = np.arange(L)
        for i in range(L - 1, 0, -1):
            j = libc.random() % (i + 1)
            order[i], order[j] = order[j], order[i]
        assert np.array_equal(ops.ushuffle_ohe(x, 1, 1, seed)[0], x[order])


def test_golden_vectors():
    g = np.load(GOLD)
    for i, (L, k, n, seed) in enumerate(CASES):
        rng = np.random.default_rng(L + k)
        x = one_hot(rng.integers(0, 4, L))
        if L == 300:
            x[10] = [1, 0, 1, 0]
            x[20] = 0
        assert np.array_equal(ops.ushuffle_ohe(x, k, n, seed), g[f"case{i}"])


def test_bad_arguments_are_rejected():
    from bpreveal_b200._lib import BprevealB200Error
    with pytest.raises(BprevealB200Error):
        ops.ushuffle_ohe(np.zeros((10, 9), dtype=np.uint8), 2, 1)
