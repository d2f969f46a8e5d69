"""Synthetic inputs of the shapes the hot path consumes: uniform-ACGT one-hot
sequences and Poisson profile counts around smooth random bumps with a guaranteed non-zero total
per region (``log(0)`` otherwise, generators.py:139-140 of the reference).  Used by ``bench.py``,
``tools/make_synthetic_data.py`` and the CLI tests -- the product's own generator, so that nothing
outside the tests needs the oracle."""
from __future__ import annotations

import numpy as np


def sequences(n, length, seed=20261019):
    """(n, length, 4) uint8 one-hot of iid uniform bases."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 4, size=(n, length))
    out = np.zeros((n, length, 4), dtype=np.uint8)
    np.put_along_axis(out, idx[..., None], 1, axis=2)
    return out


def profiles(n, length, tasks, seed=20261019, meanReads=200.0):
    """(n, length, tasks) float32 Poisson counts, ~meanReads per region and task."""
    rng = np.random.default_rng(seed + 1)
    pos = np.arange(length)[None, :, None]
    centers = rng.uniform(0, length, size=(n, 1, tasks))
    widths = rng.uniform(length / 40, length / 6, size=(n, 1, tasks))
    rate = np.exp(-0.5 * ((pos - centers) / widths) ** 2) + 0.05
    rate = rate / rate.sum(axis=1, keepdims=True) * meanReads
    out = rng.poisson(rate).astype(np.float32)
    out[:, 0, :] += (out.sum(axis=(1, 2)) == 0)[:, None]
    return out
