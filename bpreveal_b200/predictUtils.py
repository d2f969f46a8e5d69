"""Readers and the HDF5 writer of ``makePredictions`` (src/internal/predictUtils.py of the
reference: ``FastaReader`` 13-65, ``BedReader`` 68-117, ``H5Writer`` 120-231, ``addGenomeInfo`` /
``addCoordsInfo`` 234-304) -- same classes, attributes and file layout, without pysam,
pybedtools or h5py: fasta / bed parsing is plain Python, the HDF5 file is written by ``h5lite``.

Output layout (predictUtils.py:149-231): ``head_<i>/logits (N, outputLength, numTasks) f32``,
``head_<i>/logcounts (N,) f32``, ``descriptions (N,)`` variable-length UTF-8, the ``metadata``
group of internal/files.py, and for bed input ``chrom_names``, ``chrom_sizes``, ``coords_chrom``,
``coords_start``, ``coords_stop``.  The reference buffers 100 results and commits chunks through
h5py; here the datasets are contiguous and memory-mapped, results are written in place.
"""
from __future__ import annotations

import datetime
import os

import numpy as np

from . import h5lite, logUtils

LOGIT_T = np.float32
LOGCOUNT_T = np.float32
VERSION = "5.1.0-b200"


class Interval:
    """One bed line (the fields of ``pybedtools.Interval`` the hot path reads)."""

    __slots__ = ("chrom", "start", "end", "name", "score", "strand")

    def __init__(self, chrom, start, end, name=".", score=".", strand="."):
        self.chrom, self.start, self.end = chrom, int(start), int(end)
        self.name, self.score, self.strand = name, score, strand

    @property
    def stop(self):
        return self.end

    def __len__(self):
        return self.end - self.start


def readBed(bedFname):
    out = []
    with open(bedFname, "r", encoding="utf-8") as fp:
        for line in fp:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 3:
                f = line.split()
            out.append(Interval(f[0], f[1], f[2], *(f[3:6])))
    return out


class Genome:
    """Random access to a fasta file (what ``pysam.FastaFile`` provides to the reference):
    ``references``, ``nreferences``, ``get_reference_length``, ``fetch(chrom, start, end)``.
    The index (offset, line geometry per record) is built by one scan, or read from ``.fai``."""

    def __init__(self, fname):
        self.fname = fname
        self._index = {}
        self.references = []
        fai = fname + ".fai"
        if os.path.exists(fai) and os.path.getmtime(fai) >= os.path.getmtime(fname):
            with open(fai, "r", encoding="utf-8") as fp:
                for line in fp:
                    name, length, off, lb, lw = line.split("\t")[:5]
                    self._index[name] = (int(length), int(off), int(lb), int(lw))
                    self.references.append(name)
        else:
            self._scan()
        self._fp = open(fname, "rb")  # noqa: SIM115

    def _scan(self):
        with open(self.fname, "rb") as fp:
            name, length, off, lb, lw, pos = None, 0, 0, 0, 0, 0
            for raw in fp:
                if raw.startswith(b">"):
                    if name is not None:
                        self._index[name] = (length, off, lb, lw)
                    name = raw[1:].split()[0].decode("ascii")
                    self.references.append(name)
                    length, off, lb, lw = 0, pos + len(raw), 0, 0
                else:
                    body = raw.rstrip(b"\r\n")
                    if lb == 0 and body:
                        lb, lw = len(body), len(raw)
                    length += len(body)
                pos += len(raw)
            if name is not None:
                self._index[name] = (length, off, lb, lw)

    @property
    def nreferences(self):
        return len(self.references)

    def get_reference_length(self, chrom):
        return self._index[chrom][0]

    def fetch(self, chrom, start, end):
        length, off, lb, lw = self._index[chrom]
        start, end = max(0, int(start)), min(length, int(end))
        if end <= start:
            return ""
        first = off + (start // lb) * lw + start % lb
        last = off + ((end - 1) // lb) * lw + (end - 1) % lb + 1
        self._fp.seek(first)
        return self._fp.read(last - first).replace(b"\n", b"").replace(b"\r", b"").decode("ascii")

    def close(self):
        self._fp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class FastaReader:
    """predictUtils.py:13-65: ``curSequence`` / ``curLabel`` of the current record, ``pop()`` to
    advance, ``numPredictions`` records in total."""

    def __init__(self, fastaFname):
        self.numPredictions = 0
        with open(fastaFname, "r", encoding="utf-8") as fp:
            for line in fp:
                if line.startswith(">"):
                    self.numPredictions += 1
        logUtils.info(f"Found {self.numPredictions} entries in input fasta")
        self._idx = -1
        self._fp = open(fastaFname, "r", encoding="utf-8")  # noqa: SIM115
        self._nextLabel = self._fp.readline().strip()[1:]
        self.curSequence, self.curLabel = "", ""
        self.pop()

    def pop(self):
        self._idx += 1
        if self._idx >= self.numPredictions:
            return
        self.curLabel = self._nextLabel
        parts = []
        line = self._fp.readline()
        while len(line) > 1:
            if line[0] == ">":
                self._nextLabel = line[1:].strip()
                break
            parts.append(line.strip())
            line = self._fp.readline()
        self.curSequence = "".join(parts)


class BedReader:
    """predictUtils.py:68-117: regions of a bed file expanded by ``padding`` on both sides."""

    def __init__(self, bedFname, genomeFname, padding):
        self._bt = readBed(bedFname)
        self.numPredictions = len(self._bt)
        logUtils.info(f"Found {self.numPredictions} entries in input bed file")
        self._idx = 0
        self.genome = Genome(genomeFname)
        self.padding = padding
        self.curLabel = ""
        if self.numPredictions:
            self._fetch()

    def _fetch(self):
        r = self._bt[self._idx]
        self.curSequence = self.genome.fetch(r.chrom, r.start - self.padding,
                                             r.end + self.padding).upper()
        self.curChrom, self.curStart, self.curEnd = r.chrom, r.start, r.end

    def pop(self):
        self._idx += 1
        if self._idx < self.numPredictions:
            self._fetch()


def addH5Metadata(writer, **kwargs):
    """internal/files.py:8-24: the ``metadata`` group with version, date and ``kwargs``."""
    writer.create_group("metadata")
    a = writer.attrs("metadata")
    a["bpreveal_version"] = VERSION
    a["created_date"] = str(datetime.datetime.today())
    for k, v in kwargs.items():
        a[str(k)] = str(v)


def addGenomeInfo(writer, genome):
    """predictUtils.py:234-273: ``chrom_names`` / ``chrom_sizes``; returns the dtypes for
    chromosome indexes and positions and the name -> index map."""
    names = list(genome.references)
    writer.create_dataset("chrom_names", data=[str(n) for n in names], dtype=str)
    sizes = [genome.get_reference_length(n) for n in names]
    writer.create_dataset("chrom_sizes", data=np.array(sizes, dtype=np.uint64))
    chromDtype = np.uint16 if len(names) > 127 else np.uint8
    assert len(names) < 65535, "The genome has more than 2^16 chromosomes"
    posDtype = np.uint64 if any(s > 2 ** 31 - 1 for s in sizes) else np.uint32
    return chromDtype, posDtype, {n: i for i, n in enumerate(names)}


def addCoordsInfo(regions, writer, genome, stopName="coords_stop"):
    """predictUtils.py:276-304."""
    chromDtype, posDtype, idx = addGenomeInfo(writer, genome)
    writer.create_dataset("coords_chrom", data=np.array([idx[r.chrom] for r in regions],
                                                        dtype=chromDtype))
    writer.create_dataset("coords_start", data=np.array([r.start for r in regions], dtype=posDtype))
    writer.create_dataset(stopName, data=np.array([r.end for r in regions], dtype=posDtype))


class H5Writer:
    """predictUtils.py:120-231: ``addEntry((predictions, label))`` per sequence in order, then
    ``close()``.  Whole batches: ``addBatch(logits, logcounts, labels)``.

    Results are collected in arrays sized from the first result (in RAM, or in memory-mapped
    scratch files next to the output beyond ``spillBytes``) and the file is laid out once at
    ``close()``, streamed from those arrays."""

    def __init__(self, fname, numHeads, numPredictions, bedFname=None, genomeFname=None,
                 config=None, spillBytes=2 << 30):
        self._fname = fname
        self._w = h5lite.Writer()
        addH5Metadata(self._w, config=str(config))
        self.numHeads, self.numPredictions = numHeads, numPredictions
        self.writeHead = 0
        self._descriptionList = []
        self._logits = self._logcounts = None
        self._spill, self._scratch = spillBytes, []
        if bedFname is not None:
            assert genomeFname is not None, "Must supply a genome to get coordinate information."
            with Genome(genomeFname) as genome:
                addCoordsInfo(readBed(bedFname), self._w, genome)

    def buildDatasets(self, sampleOutputs):
        """Sizes the buffers from the first result (predictUtils.py:149-182)."""
        self._logits, self._logcounts = [], []
        for h in range(self.numHeads):
            shape = (self.numPredictions,) + tuple(np.shape(sampleOutputs[h]))
            nbytes = int(np.prod(shape)) * 4
            if nbytes > self._spill:
                tmp = f"{self._fname}.head{h}.scratch"
                self._scratch.append(tmp)
                self._logits.append(np.memmap(tmp, dtype=LOGIT_T, mode="w+", shape=shape))
            else:
                self._logits.append(np.empty(shape, dtype=LOGIT_T))
            self._logcounts.append(np.empty((self.numPredictions,), dtype=LOGCOUNT_T))

    def addEntry(self, batcherOut):
        preds, label = batcherOut
        if self._logits is None:
            self.buildDatasets(preds)
        self._descriptionList.append(label)
        for h in range(self.numHeads):
            self._logits[h][self.writeHead] = preds[h]
            self._logcounts[h][self.writeHead] = preds[h + self.numHeads]
        self.writeHead += 1

    def addBatch(self, logitsPerHead, logcountsPerHead, labels):
        """Vectorised ``addEntry``: ``logitsPerHead[h]`` (n, Lout, T_h), ``logcountsPerHead[h]``
        (n,) or (n, 1)."""
        n = len(labels)
        if self._logits is None:
            self.buildDatasets([lg[0] for lg in logitsPerHead])
        s = self.writeHead
        for h in range(self.numHeads):
            self._logits[h][s:s + n] = logitsPerHead[h]
            self._logcounts[h][s:s + n] = np.reshape(logcountsPerHead[h], (n,))
        self._descriptionList.extend(labels)
        self.writeHead += n

    def close(self):
        """Lays the file out and writes it (predictUtils.py:220-231 adds ``descriptions`` at
        this point as well)."""
        if self._logits is not None:
            for h in range(self.numHeads):
                self._w.create_dataset(f"head_{h}/logcounts", data=self._logcounts[h])
                self._w.create_dataset(f"head_{h}/logits", data=self._logits[h])
        self._w.create_dataset("descriptions", data=[str(d) for d in self._descriptionList],
                               dtype=str)
        logUtils.info("Closing h5.")
        self._w.save(self._fname)
        self._logits = None
        for tmp in self._scratch:
            os.remove(tmp)
