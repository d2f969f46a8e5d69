"""Counterparts of the parts of ``bpreveal.utils`` on the prediction path (src/utils.py of the
reference): ``loadModel`` (24-85), ``oneHotEncode`` / ``oneHotDecode`` (420-520),
``logitsToProfile`` (523-572), ``BatchPredictor`` (781-1192) and ``ThreadedBatchPredictor``
(1195-1460) with the same method names, argument meaning, return shapes and ordering guarantee.

What changes is where the time goes.  The reference's batcher one-hot encodes every sequence with
a Python loop over the alphabet, fills a numpy batch and calls ``model.predict`` once per 16
batches; its "threaded" variant forks N processes that each own a TensorFlow runtime because one
process cannot keep the GPU busy.  Here queued sequences are encoded with one table lookup each
straight into a pinned staging buffer, a flush runs the streaming forward of
``SoloNet.predict_host`` (H2D / kernels / D2H overlapped on three streams), and a single process
saturates the device, so ``ThreadedBatchPredictor`` is the same object (``numThreads`` accepted and
ignored, as ``BatchPredictor`` does with it in the reference).
"""
from __future__ import annotations

import queue
from collections import deque

import numpy as np

from . import logUtils
from .models import loadModel  # noqa: F401  (re-exported: utils.loadModel)

NUM_BASES = 4
ONEHOT_T = np.uint8

_LUT = {}


def _lut(alphabet):
    t = _LUT.get(alphabet)
    if t is None:
        t = np.zeros((256, NUM_BASES), dtype=ONEHOT_T)
        for i, base in enumerate(alphabet):
            t[ord(base.upper()), i] = 1
            t[ord(base.lower()), i] = 1
        _LUT[alphabet] = t
    return t


def oneHotEncode(sequence, allowN=False, alphabet="ACGT"):
    """utils.py:420-486: ``(len, 4)`` uint8, columns in ``alphabet`` order, either case; other
    characters give an all-zero row with ``allowN`` and an AssertionError without."""
    assert len(alphabet) == NUM_BASES, f"Your alphabet {alphabet} has the wrong length. " \
        f"Should be {NUM_BASES}, but got {len(alphabet)}"
    raw = np.frombuffer(sequence.encode("ascii") if isinstance(sequence, str) else sequence,
                        dtype=np.uint8)
    ret = _lut(alphabet)[raw]
    if not allowN:
        assert int(ret.sum()) == len(raw), "Sequence contains unrecognized nucleotides. " \
                                           "Maybe your sequence contains 'N'?"
    return ret


def oneHotDecode(oneHotSequence, alphabet="ACGT"):
    """utils.py:489-520: rows [1,0,0,0] -> A ... all-zero -> N."""
    assert len(alphabet) == NUM_BASES
    a = np.asarray(oneHotSequence).astype(ONEHOT_T)
    ret = np.zeros(a.shape[0], dtype=np.uint8)
    for i, base in enumerate(alphabet):
        ret += a[:, i] * np.uint8(ord(base.upper()))
    ret[ret == 0] = ord("N")
    return ret.tobytes().decode("ascii")


def logitsToProfile(logitsAcrossSingleRegion, logCountsAcrossSingleRegion):
    """utils.py:523-572: softmax over the WHOLE (length x tasks) block of one head, times
    exp(logcounts).  (Batches of regions: ``ops.logits_to_profile`` on the device.)"""
    lg = np.asarray(logitsAcrossSingleRegion, dtype=np.float64)
    e = np.exp(lg - lg.max())
    return (e / e.sum() * np.exp(float(logCountsAcrossSingleRegion))).astype(np.float32)


def tileOutputs(inputLength, outputLength, totalLength):
    """Output windows ``(start, end)`` that ``bedUtils.tileSegments(inputLength, outputLength,
    [0, totalLength), spacing=0)`` yields (bedUtils.py:93-153): consecutive ``outputLength``
    windows inside the region shrunk by the model's padding, plus one flush with its end."""
    padding = (inputLength - outputLength) // 2
    segStart, segEnd = padding, totalLength - padding
    if segEnd - segStart < outputLength:
        return []
    out = []
    startPos = segStart
    endPos = startPos + outputLength
    while endPos < segEnd:
        out.append((startPos, endPos))
        startPos += outputLength
        endPos = startPos + outputLength
    if startPos < segEnd:
        out.append((segEnd - outputLength, segEnd))
    return out


class BatchPredictor:
    """utils.py:781-1192.  Submit sequences, get results back one by one in submission order.

    :param modelFname: a model file (``.keras`` / ``.h5`` / ``.npz``) or an already-built model.
    :param batchSize: sequences per forward pass.
    :param start, numThreads, produceProfiles, quiet: accepted for API compatibility.
    """

    def __init__(self, modelFname, batchSize, start=True, numThreads=0, produceProfiles=False,
                 quiet=False):
        del start, numThreads, quiet
        self._model = loadModel(modelFname) if isinstance(modelFname, str) else modelFname
        self._inputLength = self._model.input.shape[1]
        self._outputLength = self._model.output[0].shape[1]
        outputs = self._model.output
        self._tasksPerHead = [outputs[i].shape[2] for i in range(len(outputs) // 2)]
        self._batchSize = batchSize
        self._produceProfiles = produceProfiles
        self._inQueue = deque()
        self._outQueue = deque()
        self._inWaiting = 0
        self._outWaiting = 0

    def __enter__(self):
        return self

    def __exit__(self, exceptionType, exceptionValue, exceptionTraceback):
        return exceptionType is None

    def start(self):
        """No-op (the reference's threaded batcher forks its workers here)."""

    def stop(self):
        """No-op."""

    def clear(self):
        logUtils.info(f"Clearing batch predictor, purging {self._inWaiting} inputs "
                      f"and {self._outWaiting} outputs.")
        self._inQueue.clear()
        self._outQueue.clear()
        self._inWaiting = 0
        self._outWaiting = 0

    def submitOHE(self, sequence, label):
        """utils.py:923-974.  Inputs longer than the model's input length are tiled
        (``tileOutputs``); collect those with ``getOutputProfile``."""
        sequence = np.asarray(sequence)
        if sequence.shape[0] > self._inputLength:
            tiles = tileOutputs(self._inputLength, self._outputLength, sequence.shape[0])
            for (ts, te) in tiles:
                center = (te + ts) // 2
                s = center - self._inputLength // 2
                self._inQueue.appendleft(
                    (sequence[s:s + self._inputLength], label, len(tiles),
                     ts - (self._inputLength - self._outputLength) // 2))
                self._inWaiting += 1
        else:
            self._inQueue.appendleft((sequence, label, 1, 0))
            self._inWaiting += 1
        if self._inWaiting >= self._batchSize * 16:
            self.runBatch()

    def submitString(self, sequence, label):
        self.submitOHE(oneHotEncode(sequence), label)

    def runBatch(self, maxSamples=None):
        """utils.py:986-1051."""
        if self._inWaiting == 0:
            logUtils.info("runBatch was called even though there was nothing to do.")
            return
        n = self._inWaiting if maxSamples is None else min(self._inWaiting, maxSamples)
        import torch
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.shape[0] < n:
            buf = torch.empty((max(n, self._batchSize * 16), self._inputLength, NUM_BASES),
                              dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            self._pinned = buf
        view = buf.numpy()
        meta = []
        for i in range(n):
            seq, label, numTiles, tileStart = self._inQueue.pop()
            view[i] = seq
            meta.append((label, numTiles, tileStart))
            self._inWaiting -= 1
        preds = self._model.predict(buf[:n], batch_size=self._batchSize, verbose=0)
        numHeads = len(preds) // 2
        for i, (label, numTiles, tileStart) in enumerate(meta):
            cur = [preds[j][i] for j in range(numHeads)] + \
                  [float(preds[j + numHeads][i, 0]) for j in range(numHeads)]
            self._outQueue.appendleft((cur, label, numTiles, tileStart))
            self._outWaiting += 1

    def outputReady(self):
        return self._outWaiting > 0

    def empty(self):
        return self._outWaiting == 0 and self._inWaiting == 0

    def _need_output(self):
        if not self._outWaiting:
            if self._inWaiting:
                self.runBatch()
            else:
                raise queue.Empty("There are no outputs ready, and the input queue is empty.")

    def getOutput(self):
        """-> ``([logits_head0 (Lout, T0), ..., logcounts_head0 (float), ...], label)``."""
        self._need_output()
        preds, label, numTiles, _ = self._outQueue.pop()
        assert numTiles == 1, "Attempted to getOutput but the query input size was not " \
            "equal to the model's input size. Use getOutputProfile() instead."
        self._outWaiting -= 1
        return (preds, label)

    def getOutputProfile(self):
        """utils.py:1127-1192: predicted profiles ``[(width, T_h) ...]``; tiles of a long input
        are placed at their offsets and overlaps averaged."""
        self._need_output()
        rets = [self._outQueue.pop()]
        self._outWaiting -= 1
        for _ in range(rets[0][2] - 1):
            if not self._outWaiting:
                self.runBatch()
            rets.append(self._outQueue.pop())
            self._outWaiting -= 1
        width = rets[-1][3] + self._outputLength
        H = len(self._tasksPerHead)
        prof = [np.zeros((width, t)) for t in self._tasksPerHead]
        cnt = [np.zeros((width, t)) for t in self._tasksPerHead]
        for preds, _, _, tileStart in rets:
            for h in range(H):
                prof[h][tileStart:tileStart + self._outputLength] += logitsToProfile(
                    preds[h], preds[h + H])
                cnt[h][tileStart:tileStart + self._outputLength] += 1
        for h in range(H):
            assert cnt[h].min() >= 1, "Missed placing an output somewhere! Please report this bug."
            prof[h] /= cnt[h]
        return (prof, rets[0][1])


class ThreadedBatchPredictor(BatchPredictor):
    """utils.py:1195-1460.  The reference forks ``numThreads`` processes, each with its own
    TensorFlow runtime, to keep one GPU busy; one process does that here, so this is
    ``BatchPredictor`` (order-preserving as the reference guarantees, utils.py:1221-1223).  Shard
    windows ACROSS GPUs with ``parallel.shard_range`` (one process per GPU)."""

    def __init__(self, modelFname, batchSize, start=False, numThreads=1, produceProfiles=False,
                 quiet=False):
        super().__init__(modelFname, batchSize, start, numThreads, produceProfiles, quiet)


def easyPredict(sequences, modelFname):
    """utils.py:578-667: predicted profiles for a string or a list of strings."""
    single = isinstance(sequences, str)
    seqs = [sequences] if single else list(sequences)
    predictor = BatchPredictor(modelFname, 64)
    for s in seqs:
        predictor.submitString(s, 1)
    out = [predictor.getOutputProfile()[0] for _ in seqs]
    return out[0] if single else out
