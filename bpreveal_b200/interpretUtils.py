"""Streaming attribution: generators, savers and the runner of interpretPisa / interpretFlat.

Counterpart of src/internal/interpretUtils.py of the reference: ``Query`` / ``Result`` (29-114),
the generators ``ListGenerator`` / ``FastaGenerator`` / ``FlatBedGenerator`` /
``PisaBedGenerator`` (767-1011), the savers ``FlatListSaver`` (326), ``FlatH5Saver`` (425) and
``PisaH5Saver`` (605), and ``InterpRunner`` (216-323) -- same class names, constructor arguments
and file layouts.

The reference spreads this over 2 + numThreads processes connected by pickling queues (one
generator, numThreads ``_ShapBatcher`` s with a TensorFlow runtime each, one saver per metric)
and explains ten queries at a time with 10 x (21 + 40 + 40) network passes.  Here the runner is a
loop in one process: queries are gathered into batches, ``interpret.pisa_batch`` /
``interpret.flat_batch`` explain a batch in one CUDA-graph replay (2 n + 1 passes per query), and
results are handed to the savers in query order.  Output files (``h5lite``): PISA --
``shap (N, receptiveField, 4) f16`` and ``sequence`` u8, gzip, chunks of 128; ``input_predictions
(N,)``, ``shuffle_predictions (N, numShuffles)`` f32; ``coords_chrom`` / ``coords_base`` +
``chrom_names`` / ``chrom_sizes`` (bed) or ``descriptions`` (fasta); flat -- ``hyp_scores``,
``input_seqs``, ``input_predictions`` and ``coords_chrom`` / ``coords_start`` / ``coords_end``.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import h5lite, logUtils
from .predictUtils import Genome, addH5Metadata, readBed
from .utils import oneHotEncode

NUM_BASES = 4
H5_CHUNK_SIZE = 128          # internal/constants.py:156
ONEHOT_T = np.uint8          # internal/constants.py:9-40
PRED_T = np.float32
IMPORTANCE_T = np.float16


# ---------------------------------------------------------------------------------------------
# metrics: the reference passes closures over a Keras model (interpretPisa.py:153-166,
# interpretFlat.py:187-250); the three built-in ones are plain descriptions here
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class PisaMetric:
    """``model.outputs[headID][:, 0, taskID]``: the leftmost logit of one task."""
    headID: int
    taskID: int


@dataclasses.dataclass(frozen=True)
class ProfileMetric:
    """interpretFlat.py:187-237: weighted mean-normalised logits of ``taskIDs`` of one head."""
    headID: int
    taskIDs: tuple


@dataclasses.dataclass(frozen=True)
class CountsMetric:
    """interpretFlat.py:240-250: ``model.outputs[numHeads + headID][:, 0]``."""
    numHeads: int
    headID: int


class Query:
    """interpretUtils.py:29-64."""

    def __init__(self, oneHotSequence, passData, index):
        self.sequence, self.passData, self.index = oneHotSequence, passData, index

    def __str__(self):
        return f"seq{self.sequence.shape} pass {self.passData} idx {self.index}"


class Result:
    """interpretUtils.py:67-114."""

    def __init__(self, inputPrediction, shufflePredictions, sequence, shap, passData, index):
        self.inputPrediction, self.shufflePredictions = inputPrediction, shufflePredictions
        self.sequence, self.shap, self.passData, self.index = sequence, shap, passData, index


# ---------------------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------------------
class Generator:
    numRegions: int

    def construct(self):
        pass

    def done(self):
        pass

    def __iter__(self):
        return self


class ListGenerator(Generator):
    """interpretUtils.py:767-816: an iterable of strings (and optional pass-through data)."""

    def __init__(self, sequences, passDataList=None):
        self._sequences = list(sequences)
        self.numSamples = self.numRegions = len(self._sequences)
        self.inputLength = len(self._sequences[0])
        self._passData = list(passDataList) if passDataList is not None else \
            [""] * len(self._sequences)
        self._readHead = 0

    def __next__(self):
        if self._readHead == len(self._sequences):
            raise StopIteration()
        i = self._readHead
        self._readHead += 1
        return Query(oneHotEncode(self._sequences[i]), self._passData[i], i)


class FastaGenerator(Generator):
    """interpretUtils.py:819-876: one query per fasta record, passData = the record's name."""

    def __init__(self, fastaFname):
        self.fastaFname = fastaFname
        n = 0
        with open(fastaFname, "r", encoding="utf-8") as fp:
            for line in fp:
                if line.startswith(">"):
                    n += 1
        self.numRegions = n
        self.index = 0
        logUtils.info(f"Fasta generator initialized with {self.numRegions} regions.")

    def construct(self):
        self.fastaFile = open(self.fastaFname, "r", encoding="utf-8")  # noqa: SIM115
        self.nextSequenceID = self.fastaFile.readline()[1:].strip()
        self.nowStop = self.numRegions == 0

    def done(self):
        self.fastaFile.close()

    def __next__(self):
        if self.nowStop:
            raise StopIteration()
        seq, prevID = [], self.nextSequenceID
        while True:
            line = self.fastaFile.readline()
            if len(line) == 0:
                self.nowStop = True
                break
            if line[0] == ">":
                self.nextSequenceID = line[1:].strip()
                break
            seq.append(line.strip())
        q = Query(oneHotEncode("".join(seq)), prevID, self.index)
        self.index += 1
        return q


class FlatBedGenerator(Generator):
    """interpretUtils.py:879-935: one query per bed region; the window is the region's start
    minus the model's padding, ``inputLength`` long; passData = (chrom, start, end)."""

    def __init__(self, bedFname, genomeFname, inputLength, outputLength):
        self.bedFname, self.genomeFname = bedFname, genomeFname
        self.inputLength, self.outputLength = inputLength, outputLength
        self.numRegions = len(readBed(bedFname))
        logUtils.info(f"Bed generator initialized with {self.numRegions} regions")

    def construct(self):
        self.shapTargets = readBed(self.bedFname)
        self.genome = Genome(self.genomeFname)
        self.readHead = 0

    def done(self):
        self.genome.close()

    def __next__(self):
        if self.readHead >= len(self.shapTargets):
            raise StopIteration()
        r = self.shapTargets[self.readHead]
        padding = (self.inputLength - self.outputLength) // 2
        startPos = r.start - padding
        endPos = startPos + self.inputLength
        seq = self.genome.fetch(r.chrom, startPos, endPos)
        ret = Query(oneHotEncode(seq), (r.chrom, startPos, endPos), self.readHead)
        self.readHead += 1
        return ret


class PisaBedGenerator(Generator):
    """interpretUtils.py:938-1011: one query per BASE of every bed region: the window whose
    leftmost output is that base; passData = (chrom, position)."""

    def __init__(self, bedFname, genomeFname, inputLength, outputLength):
        self.bedFname, self.genomeFname = bedFname, genomeFname
        self.inputLength, self.outputLength = inputLength, outputLength
        self.numRegions = sum(len(r) for r in readBed(bedFname))
        logUtils.info(f"Bed generator initialized with {self.numRegions} regions")

    def construct(self):
        self.shapTargets = [(r.chrom, pos) for r in readBed(self.bedFname)
                            for pos in range(r.start, r.end)]
        self.genome = Genome(self.genomeFname)
        self.readHead = 0

    def done(self):
        self.genome.close()

    def __next__(self):
        if self.readHead >= len(self.shapTargets):
            raise StopIteration()
        curChrom, curStart = self.shapTargets[self.readHead]
        padding = (self.inputLength - self.outputLength) // 2
        startPos = curStart - padding
        seq = self.genome.fetch(curChrom, startPos, startPos + self.inputLength)
        assert len(seq) == self.inputLength, f"Ran off the edge of {curChrom} at {curStart}"
        ret = Query(oneHotEncode(seq), (curChrom, curStart), self.readHead)
        self.readHead += 1
        return ret


# ---------------------------------------------------------------------------------------------
# savers
# ---------------------------------------------------------------------------------------------
class Saver:
    def construct(self):
        pass

    def parentFinish(self):
        pass

    def add(self, result):
        raise NotImplementedError

    def addBatch(self, results):
        for r in results:
            self.add(r)

    def done(self):
        pass


class FlatListSaver(Saver):
    """interpretUtils.py:326-422: keeps everything in arrays (``shap``, ``seq``,
    ``inputPredictions``) for interactive use (utils.easyInterpretFlat)."""

    def __init__(self, numSamples, inputLength):
        self.numSamples, self.inputLength = numSamples, inputLength

    def construct(self):
        n, L = self.numSamples, self.inputLength
        self.shap = np.zeros((n, L, NUM_BASES), dtype=IMPORTANCE_T)
        self.seq = np.zeros((n, L, NUM_BASES), dtype=ONEHOT_T)
        self.inputPredictions = np.zeros((n,), dtype=PRED_T)

    def add(self, result):
        self.shap[result.index] = result.shap
        self.seq[result.index] = result.sequence
        self.inputPredictions[result.index] = result.inputPrediction


class _H5SaverBase(Saver):
    def _genome_tables(self, w, coordNames):
        with Genome(self.genomeFname) as genome:
            names = list(genome.references)
            sizes = [genome.get_reference_length(n) for n in names]
        posDtype = "u8" if any(s > 2 ** 31 for s in sizes) else "u4"
        w.create_dataset("chrom_names", data=[str(n) for n in names], dtype=str)
        w.create_dataset("chrom_sizes", data=np.array(sizes, dtype=posDtype))
        self.chromNameToIdx = {n: i for i, n in enumerate(names)}
        chromDtype = "u1" if len(names) <= 255 else ("u2" if len(names) <= 65535 else "u4")
        self._coords = {"coords_chrom": np.zeros((self.numSamples,), dtype=chromDtype)}
        for nm in coordNames:
            self._coords[nm] = np.zeros((self.numSamples,), dtype=posDtype if nm != "coords_base"
                                        else "u4")


class FlatH5Saver(_H5SaverBase):
    """interpretUtils.py:425-602."""

    def __init__(self, outputFname, numSamples, inputLength, genome=None, useTqdm=False,
                 config=None):
        self.chunkShape = (min(H5_CHUNK_SIZE, numSamples), inputLength, NUM_BASES)
        self._outputFname, self.numSamples = outputFname, numSamples
        self.genomeFname, self.inputLength = genome, inputLength
        self.config = str(config)
        del useTqdm

    def construct(self):
        n, L = self.numSamples, self.inputLength
        self._w = h5lite.Writer()
        addH5Metadata(self._w, config=self.config)
        self._hyp = np.zeros((n, L, NUM_BASES), dtype=IMPORTANCE_T)
        self._seqs = np.zeros((n, L, NUM_BASES), dtype=ONEHOT_T)
        self._preds = np.zeros((n,), dtype=PRED_T)
        self._desc = [""] * n
        if self.genomeFname is not None:
            self._genome_tables(self._w, ("coords_start", "coords_end"))

    def add(self, result):
        i = result.index
        self._hyp[i] = result.shap
        self._seqs[i] = result.sequence
        self._preds[i] = result.inputPrediction
        if self.genomeFname is not None:
            self._coords["coords_chrom"][i] = self.chromNameToIdx[result.passData[0]]
            self._coords["coords_start"][i] = result.passData[1]
            self._coords["coords_end"][i] = result.passData[2]
        else:
            self._desc[i] = str(result.passData)

    def done(self):
        w, c = self._w, self.chunkShape
        w.create_dataset("hyp_scores", data=self._hyp, chunks=c, compression="gzip")
        w.create_dataset("input_predictions", data=self._preds, chunks=(c[0],), compression="gzip")
        w.create_dataset("input_seqs", data=self._seqs, chunks=c, compression="gzip")
        if self.genomeFname is not None:
            for k, v in self._coords.items():
                w.create_dataset(k, data=v)
        else:
            w.create_dataset("descriptions", data=self._desc, dtype=str)
            coords = getattr(self, "coordinates", None)
            if coords is not None:  # interpretFlat.py:303-311: fasta input + separate coordinates
                from .predictUtils import addCoordsInfo
                with Genome(coords[1]) as genome:
                    addCoordsInfo(coords[0], w, genome, stopName="coords_end")
        w.save(self._outputFname)


class PisaH5Saver(_H5SaverBase):
    """interpretUtils.py:605-764: only the first ``receptiveField`` rows of every attribution
    map are stored -- the rest is identically zero."""

    def __init__(self, outputFname, numSamples, numShuffles, receptiveField, genome=None,
                 useTqdm=False, config=None):
        self._outputFname, self.numSamples, self.numShuffles = outputFname, numSamples, numShuffles
        self.genomeFname, self.receptiveField = genome, receptiveField
        self.chunkShape = (min(H5_CHUNK_SIZE, numSamples), receptiveField, NUM_BASES)
        self.config = str(config)
        del useTqdm

    def construct(self):
        n, rf = self.numSamples, self.receptiveField
        self._w = h5lite.Writer()
        addH5Metadata(self._w, config=self.config)
        self._shap = np.zeros((n, rf, NUM_BASES), dtype=IMPORTANCE_T)
        self._seq = np.zeros((n, rf, NUM_BASES), dtype=ONEHOT_T)
        self._inPred = np.zeros((n,), dtype=PRED_T)
        self._shufPred = np.zeros((n, self.numShuffles), dtype=PRED_T)
        self._desc = [""] * n
        if self.genomeFname is not None:
            self._genome_tables(self._w, ("coords_base",))

    def add(self, result):
        i, rf = result.index, self.receptiveField
        self._shap[i] = result.shap[:rf]
        self._seq[i] = result.sequence[:rf]
        self._inPred[i] = result.inputPrediction
        self._shufPred[i] = result.shufflePredictions
        if self.genomeFname is not None:
            self._coords["coords_chrom"][i] = self.chromNameToIdx[result.passData[0]]
            self._coords["coords_base"][i] = result.passData[1]
        else:
            self._desc[i] = str(result.passData)

    def addArrays(self, start, shap, sequence, inputPredictions, shufflePredictions, passData):
        """A whole batch at once: rows ``start .. start + n`` (arrays already on the host)."""
        n, rf = shap.shape[0], self.receptiveField
        self._shap[start:start + n] = shap[:, :rf]
        self._seq[start:start + n] = sequence[:, :rf]
        self._inPred[start:start + n] = inputPredictions
        self._shufPred[start:start + n] = shufflePredictions
        for k, pd in enumerate(passData):
            if self.genomeFname is not None:
                self._coords["coords_chrom"][start + k] = self.chromNameToIdx[pd[0]]
                self._coords["coords_base"][start + k] = pd[1]
            else:
                self._desc[start + k] = str(pd)

    def done(self):
        w, c = self._w, self.chunkShape
        w.create_dataset("input_predictions", data=self._inPred)
        w.create_dataset("shuffle_predictions", data=self._shufPred)
        w.create_dataset("shap", data=self._shap, chunks=c, compression="gzip")
        w.create_dataset("sequence", data=self._seq, chunks=c, compression="gzip")
        if self.genomeFname is not None:
            for k, v in self._coords.items():
                w.create_dataset(k, data=v)
        else:
            w.create_dataset("descriptions", data=self._desc, dtype=str)
        w.save(self._outputFname)


# ---------------------------------------------------------------------------------------------
# runner
# ---------------------------------------------------------------------------------------------
class InterpRunner:
    """interpretUtils.py:216-323: pulls queries from ``generator``, explains them with every
    metric and hands ``Result`` s to the matching saver (``savers[i]`` belongs to ``metrics[i]``).

    :param modelFname: a model file or an already-built solo model.
    :param metrics: ``PisaMetric`` / ``ProfileMetric`` / ``CountsMetric`` descriptions.
    :param batchSize: queries per device batch (the reference uses 10; its batch is 10 x (n + 1)
        sequences, here a batch is one CUDA-graph replay over ``batchSize * (n + 1)`` windows).
    :param numThreads: accepted for compatibility (the reference's way to keep one GPU busy).
    :param backend: "shap" (the ISM backend of the reference is a slow demo path, not built).
    :param shard: ``(rank, world)`` -- explain only this rank's contiguous range of queries
        (``parallel.shard_range``); results keep their global indexes.
    """

    def __init__(self, modelFname, metrics, batchSize, generator, savers, numShuffles, kmerSize,
                 numThreads=1, backend="shap", useHypotheticalContribs=False, shard=None,
                 device=None, precise=True):
        if backend != "shap":
            raise ValueError("only the 'shap' backend is implemented")
        assert len(metrics) == len(savers), "one saver per metric"
        del numThreads
        from .models import SoloModel, loadModel
        model = loadModel(modelFname, device=device, precise=precise) \
            if isinstance(modelFname, str) else modelFname
        self.model = model
        self.solo = isinstance(model, SoloModel)
        # transformation / combined models go through interpret.model_shap_batch (rescale rule
        # on Sigmoid / Relu / Exp / Log, shap.py:782-793), solo models through the graph-captured
        # receptive-field path
        self.net = model.net if self.solo else None
        self.metrics, self.savers, self.generator = list(metrics), list(savers), generator
        self.batchSize, self.numShuffles, self.kmerSize = batchSize, numShuffles, kmerSize
        self.hypothetical = useHypotheticalContribs
        self.shard = shard

    def _explain(self, metric, x):
        from . import interpret
        if isinstance(metric, PisaMetric):
            if self.solo:
                return interpret.pisa_batch(self.net, x, metric.headID, metric.taskID,
                                            self.numShuffles, hypothetical=self.hypothetical,
                                            kmerSize=self.kmerSize)
            return interpret.pisa_model(self.model, x, metric.headID, metric.taskID,
                                        self.numShuffles, kmerSize=self.kmerSize)
        if isinstance(metric, ProfileMetric):
            return interpret.flat_model(self.model, x, metric.headID, self.numShuffles,
                                        kind="profile", taskIDs=list(metric.taskIDs),
                                        kmerSize=self.kmerSize)
        if isinstance(metric, CountsMetric):
            return interpret.flat_model(self.model, x, metric.headID, self.numShuffles,
                                        kind="counts", kmerSize=self.kmerSize)
        raise TypeError(f"unknown metric {metric!r}")

    def run(self):
        import torch
        dev = self.model._device
        for s in self.savers:
            s.construct()
        self.generator.construct()
        lo, hi = 0, self.generator.numRegions
        if self.shard is not None:
            from .parallel import shard_range
            lo, hi = shard_range(self.generator.numRegions, *self.shard)
        batch = []

        def flush():
            if not batch:
                return
            x = torch.as_tensor(np.stack([q.sequence for q in batch])).to(dev)
            for metric, saver in zip(self.metrics, self.savers):
                phi, vx, vr = self._explain(metric, x)
                phi, vx, vr = phi.cpu().numpy(), vx.cpu().numpy(), vr.cpu().numpy()
                for k, q in enumerate(batch):
                    saver.add(Result(vx[k], vr[k], q.sequence, phi[k], q.passData, q.index))
            batch.clear()

        for q in self.generator:
            if not lo <= q.index < hi:
                continue
            batch.append(q)
            if len(batch) == self.batchSize:
                flush()
        flush()
        self.generator.done()
        for s in self.savers:
            s.done()
            s.parentFinish()
