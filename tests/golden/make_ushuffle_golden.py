"""Writes tests/golden/ushuffle_vectors.npz with the REFERENCE's own libushuffle.c (compiled into
oracle/_ref by `make -C oracle`; needs /root/reference), so tests/test_ushuffle_cpu.py can also
check bpb_ushuffle_ohe where the reference is absent (the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_ushuffle_cpu as T  # noqa: E402

out = {}
for i, (L, k, n, seed) in enumerate(T.CASES):
    rng = np.random.default_rng(L + k)
    x = T.one_hot(rng.integers(0, 4, L))
    if L == 300:
        x[10] = [1, 0, 1, 0]
        x[20] = 0
    out[f"case{i}"] = T.ref_shuffle(x, k, n, seed)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ushuffle_vectors.npz"), **out)
print({k: v.shape for k, v in out.items()})
