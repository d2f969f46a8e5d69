"""Writes a synthetic training file in the layout of prepareTrainingData
(src/prepareTrainingData.py:209-230 of the reference): ``sequence (N, inputLength + 2 J, 4)`` u8
one-hot, ``head_<i> (N, outputLength + 2 J, numTasks)`` f32 counts, gzip, chunks of 128 regions.
Sequences are uniform ACGT, profiles Poisson counts around random bumps."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bpreveal_b200 import h5lite, synthetic  # noqa: E402


def write(fname, numRegions, inputLength, outputLength, maxJitter, headTasks, seed=20261019,
          meanReads=200.0, compress=True):
    w = h5lite.Writer()
    kw = dict(chunks=True, compression="gzip") if compress else {}
    w.create_dataset("sequence",
                     data=synthetic.sequences(numRegions, inputLength + 2 * maxJitter, seed), **kw)
    for i, t in enumerate(headTasks):
        w.create_dataset(f"head_{i}",
                         data=synthetic.profiles(numRegions, outputLength + 2 * maxJitter, t,
                                                 seed + 17 * (i + 1), meanReads), **kw)
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
