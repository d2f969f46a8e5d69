"""Device-resident counterpart of ``bpreveal.generators.H5BatchGenerator``
(src/generators.py:17-140 of the reference).

The reference keeps the padded dataset in host memory and, once per epoch, shuffles the region
order and re-crops every region at a random jitter offset with the C helpers ``slide`` /
``slideChar`` (internal/libslide.c:44-91, "takes multiple seconds").  Here the padded arrays live
in HBM, the epoch's ``(regionIndexes, sliceCols)`` come from the same ``default_rng(1234)`` stream
(``parallel.EpochPlan``), and the crop / shuffle / log-count targets are three kernel launches
(``bpb_slide_u8``, ``bpb_slide_f32``, ``bpb_log_row_sums``).  Batches are handed to the model as
device tensors, so an epoch involves no host-device traffic at all.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .parallel import EpochPlan


class DeviceBatchGenerator:
    """``keras.utils.Sequence`` look-alike: ``len(gen)`` batches per epoch,
    ``gen[i] -> (sequences, (profiles_h..., logcounts_h...))``, ``gen.on_epoch_end()``.

    :param headList: the heads section of the configuration (mutated: ``INTERNAL_mean-counts``
        is added like generators.py:57-75).
    :param sequences: (numRegions, inputLength + 2*maxJitter, 4) uint8 one-hot (the ``sequence``
        dataset of the training h5).
    :param profiles: list over heads of (numRegions, outputLength + 2*maxJitter, numTasks) float32
        (the ``head_N`` datasets).
    """

    def __init__(self, headList, sequences, profiles, inputLength, outputLength, maxJitter,
                 batchSize, device=None, seed=1234):
        dev = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        self.headList = headList
        self.inputLength, self.outputLength = inputLength, outputLength
        self.maxJitter, self.batchSize = maxJitter, batchSize
        self.fullSequences = torch.as_tensor(np.ascontiguousarray(sequences)).to(
            device=dev, dtype=torch.uint8).contiguous()
        self.numRegions = self.fullSequences.shape[0]
        assert self.fullSequences.shape[1] == inputLength + 2 * maxJitter, \
            "sequence dataset must be inputLength + 2 * maxJitter wide"
        self.fullData = [torch.as_tensor(np.ascontiguousarray(p)).to(
            device=dev, dtype=torch.float32).contiguous() for p in profiles]
        for p in self.fullData:
            assert p.shape[1] == outputLength + 2 * maxJitter
        self._seq = torch.empty((self.numRegions, inputLength, 4), dtype=torch.uint8, device=dev)
        self._vals = [torch.empty((self.numRegions, outputLength, p.shape[2]), dtype=torch.float32,
                                  device=dev) for p in self.fullData]
        self._counts = [torch.empty((self.numRegions,), dtype=torch.float32, device=dev)
                        for _ in self.fullData]
        self.plan = EpochPlan(self.numRegions, maxJitter, seed=seed)
        self.refreshData()
        self.addMeanCounts()

    def addMeanCounts(self):
        """generators.py:57-75: average counts per region over the un-jittered output window."""
        J = self.maxJitter
        for i, head in enumerate(self.headList):
            win = self.fullData[i][:, J:self.fullData[i].shape[1] - J, :] if J > 0 else self.fullData[i]
            head["INTERNAL_mean-counts"] = float(win.sum().item()) / self.numRegions

    def __len__(self):
        return math.ceil(self.numRegions / self.batchSize)

    def __getitem__(self, idx):
        s = idx * self.batchSize
        e = min((idx + 1) * self.batchSize, self.numRegions)
        vals = [v[s:e] for v in self._vals]
        counts = [c[s:e] for c in self._counts]
        return self._seq[s:e], tuple(vals + counts)

    def refreshData(self):
        """generators.py:108-140: new region order and jitter offsets, then crop everything."""
        regionIndexes, sliceCols = self.plan.next_epoch()
        dev = self._seq.device
        rows = torch.as_tensor(regionIndexes.astype(np.int32), device=dev)
        cols = torch.as_tensor(sliceCols.astype(np.int32), device=dev)
        ops.slide_u8(self.fullSequences, self._seq, rows, cols)
        for i, full in enumerate(self.fullData):
            ops.slide_f32(full, self._vals[i], rows, cols)
            ops.log_row_sums(self._vals[i], self._counts[i])
        self.regionIndexes, self.sliceCols = regionIndexes, sliceCols

    def on_epoch_end(self):
        self.refreshData()


class H5BatchGenerator(DeviceBatchGenerator):
    """``generators.H5BatchGenerator`` with the reference's constructor (generators.py:32-56):
    ``dataH5`` is an opened training file (``h5lite.File``, an ``h5py.File`` or any mapping) with
    ``sequence (N, inputLength + 2 maxJitter, 4)`` u8 and ``head_<i> (N, outputLength + 2
    maxJitter, numTasks)`` f32 as written by prepareTrainingData (prepareTrainingData.py:209-230).
    The whole file goes to HBM once (the reference loads it into host RAM, generators.py:43-53)."""

    def __init__(self, headList, dataH5, inputLength, outputLength, maxJitter, batchSize,
                 device=None, seed=1234):
        seqs = np.asarray(dataH5["sequence"][...])
        profiles = [np.asarray(dataH5[f"head_{i}"][...]) for i in range(len(headList))]
        super().__init__(headList, seqs, profiles, inputLength, outputLength, maxJitter, batchSize,
                         device=device, seed=seed)
