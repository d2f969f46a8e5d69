"""Writes a synthetic training file in the layout of prepareTrainingData
(src/prepareTrainingData.py:209-230 of the reference): ``sequence (N, inputLength + 2 J, 4)`` u8
one-hot, ``head_<i> (N, outputLength + 2 J, numTasks)`` f32 counts, gzip, chunks of 128 regions.
Sequences are uniform ACGT, profiles Poisson counts around random bumps."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bpreveal_b200 import h5lite  # noqa: E402


def sequences(n, length, rng):
    idx = rng.integers(0, 4, size=(n, length))
    out = np.zeros((n, length, 4), dtype=np.uint8)
    np.put_along_axis(out, idx[..., None], 1, axis=2)
    return out


def profiles(n, length, tasks, rng, meanReads=200.0):
    pos = np.arange(length)[None, :, None]
    centers = rng.uniform(0, length, size=(n, 1, tasks))
    widths = rng.uniform(length / 40, length / 6, size=(n, 1, tasks))
    rate = np.exp(-0.5 * ((pos - centers) / widths) ** 2) + 0.05
    rate = rate / rate.sum(axis=1, keepdims=True) * meanReads
    out = rng.poisson(rate).astype(np.float32)
    out[:, length // 2, :] += 1.0  # never an empty region: log(0) in generators.py:139-140
    return out


def write(fname, numRegions, inputLength, outputLength, maxJitter, headTasks, seed=20261019,
          meanReads=200.0, compress=True):
    rng = np.random.default_rng(seed)
    w = h5lite.Writer()
    kw = dict(chunks=True, compression="gzip") if compress else {}
    w.create_dataset("sequence", data=sequences(numRegions, inputLength + 2 * maxJitter, rng), **kw)
    for i, t in enumerate(headTasks):
        w.create_dataset(f"head_{i}", data=profiles(numRegions, outputLength + 2 * maxJitter, t,
                                                    rng, meanReads), **kw)
    w.save(fname)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("out")
    ap.add_argument("--regions", type=int, default=1024)
    ap.add_argument("--input-length", type=int, default=3090)
    ap.add_argument("--output-length", type=int, default=1000)
    ap.add_argument("--max-jitter", type=int, default=100)
    ap.add_argument("--tasks", type=int, nargs="+", default=[2])
    ap.add_argument("--seed", type=int, default=20261019)
    a = ap.parse_args()
    write(a.out, a.regions, a.input_length, a.output_length, a.max_jitter, a.tasks, a.seed)
