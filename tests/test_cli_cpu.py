"""Host side of the config-driven programs: the reference's schema corpus, the readers / writers /
savers and the batch predictor's queueing, tiling and stitching (no GPU: the model is a stub with
the Keras duck-type the predictor touches)."""
import json
import os

import jsonschema
import numpy as np
import pytest

from bpreveal_b200 import h5lite, interpretUtils, predictUtils, schema, utils

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(GOLD, "schema_cases.json"), encoding="utf-8"))["cases"]
HOT = [k for k, c in CASES.items() if c["program"] in schema.schemaMap]


@pytest.mark.parametrize("name", HOT)
def test_schema_corpus_of_the_reference(name):
    """src/schematools/runTest.py:15-57: good configurations validate, bad ones do not, and no
    configuration validates under ANOTHER program's schema."""
    c = CASES[name]
    v = schema.schemaMap[c["program"]]
    if c["good"]:
        v.validate(c["config"])
    else:
        with pytest.raises(jsonschema.ValidationError):
            v.validate(c["config"])
    for other, ov in schema.schemaMap.items():
        if other in (c["program"], "base"):
            continue
        with pytest.raises(jsonschema.ValidationError):
            ov.validate(c["config"])


def test_one_hot_helpers_match_the_reference_docstrings():
    # utils.py:456-464
    assert utils.oneHotEncode("ACGTTT").tolist() == [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0],
                                                     [0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 0, 1]]
    assert utils.oneHotEncode("acgt").tolist() == np.eye(4, dtype=np.uint8).tolist()
    with pytest.raises(AssertionError):
        utils.oneHotEncode("ACGN")
    x = utils.oneHotEncode("ACGNTRY", allowN=True)
    assert x.dtype == np.uint8 and x[3].sum() == 0 and x[5].sum() == 0
    assert utils.oneHotDecode(x) == "ACGNTNN"
    # revcomp = np.flip (utils.py:442-449)
    assert utils.oneHotDecode(np.flip(utils.oneHotEncode("AACG"))) == "CGTT"
    golden = np.load(os.path.join(GOLD, "reference_answers.npz"))
    n = 0
    for k in golden.files:  # outputs of the reference's own utils.oneHotEncode (make_golden.py)
        if k.startswith("ohe_seq"):
            assert np.array_equal(utils.oneHotEncode(str(golden[k]), allowN=True),
                                  golden["ohe_out" + k[7:]])
            n += 1
    assert n >= 3


def test_logits_to_profile_against_the_reference_run():
    golden = np.load(os.path.join(GOLD, "reference_answers.npz"))
    got = utils.logitsToProfile(golden["l2p_logits"], float(golden["l2p_logcounts"]))
    assert np.allclose(got, golden["l2p_profile"], rtol=1e-5)
    lg = np.random.default_rng(0).standard_normal((50, 2))
    p = utils.logitsToProfile(lg, 3.0)
    assert p.shape == (50, 2) and np.isclose(p.sum(), np.exp(3.0), rtol=1e-5)


def test_tiling_of_long_inputs():
    # bedUtils.py:112-121 draws this example with spacing 5; spacing 0 is what the predictor uses
    assert utils.tileOutputs(10, 6, 30) == [(2, 8), (8, 14), (14, 20), (20, 26), (22, 28)]
    assert utils.tileOutputs(10, 6, 10) == [(2, 8)]
    assert utils.tileOutputs(10, 6, 9) == []
    for (li, lo, tot) in [(52, 20, 300), (3090, 1000, 7000), (100, 100, 450)]:
        tiles = utils.tileOutputs(li, lo, tot)
        pad = (li - lo) // 2
        cover = np.zeros(tot, dtype=int)
        for s, e in tiles:
            assert e - s == lo and s >= pad and e <= tot - pad
            cover[s:e] += 1
        assert cover[pad:tot - pad].min() >= 1 and cover[:pad].sum() == 0


class StubModel:
    """predict(x) -> logits that encode the window (sum of the one-hot's A column per output
    position) so that tiling / stitching can be checked exactly."""

    def __init__(self, Lin, Lout, tasks):
        self.Lin, self.Lout, self.tasks = Lin, Lout, tasks
        shape = lambda *s: type("T", (), {"shape": s})()  # noqa: E731
        self.input = shape(None, Lin, 4)
        self.output = [shape(None, Lout, t) for t in tasks] + [shape(None, 1) for _ in tasks]
        self.calls = []

    def predict(self, x, batch_size=64, verbose=0):
        x = np.asarray(x)
        self.calls.append(x.shape[0])
        pad = (self.Lin - self.Lout) // 2
        core = x[:, pad:pad + self.Lout, :].astype(np.float32)
        outs = [np.repeat(core[:, :, :1] * (h + 1), t, axis=2) for h, t in enumerate(self.tasks)]
        outs += [core[:, :, 1].sum(axis=1, keepdims=True) * 0.01 + h for h in range(len(self.tasks))]
        return outs


def test_batch_predictor_orders_batches_and_stitches():
    rng = np.random.default_rng(1)
    m = StubModel(30, 10, [2, 1])
    bp = utils.BatchPredictor(m, batchSize=4)
    seqs = ["".join(rng.choice(list("ACGT"), size=30)) for _ in range(70)]
    got = []
    with bp:
        for i, s in enumerate(seqs):
            bp.submitString(s, i)
            while bp.outputReady():
                got.append(bp.getOutput())
        while not bp.empty():
            got.append(bp.getOutput())
    assert [g[1] for g in got] == list(range(70))              # submission order
    assert m.calls[0] == 64 and sum(m.calls) == 70             # flush at 16 x batchSize
    for (preds, label) in got:
        x = utils.oneHotEncode(seqs[label])
        assert len(preds) == 4 and preds[0].shape == (10, 2) and isinstance(preds[2], float)
        assert np.array_equal(preds[0][:, 0], x[10:20, 0])
        assert preds[3] == pytest.approx(x[10:20, 1].sum() * 0.01 + 1)
    with pytest.raises(Exception):
        bp.getOutput()                                          # queue.Empty
    # a long input is tiled; profiles are stitched with overlaps averaged
    long = "".join(rng.choice(list("ACGT"), size=75))
    bp.submitString(long, "long")
    with pytest.raises(AssertionError, match="getOutputProfile"):
        bp.getOutput()
    bp.clear()
    bp.submitString(long, "long")
    prof, label = bp.getOutputProfile()
    assert label == "long" and prof[0].shape == (55, 2) and prof[1].shape == (55, 1)
    tiles = utils.tileOutputs(30, 10, 75)
    want = np.zeros((55, 2))
    cnt = np.zeros((55, 2))
    xl = utils.oneHotEncode(long)
    for (s, e) in tiles:
        win = xl[s - 10:s - 10 + 30]
        lg = np.repeat(win[10:20, :1].astype(np.float32), 2, axis=1)
        want[s - 10:e - 10] += utils.logitsToProfile(lg, win[10:20, 1].sum() * 0.01)
        cnt[s - 10:e - 10] += 1
    assert np.allclose(prof[0], want / cnt, rtol=1e-5)


def _write_genome(tmp_path, rng):
    seqs = {"chrA": "".join(rng.choice(list("ACGT"), size=1000)),
            "chrB": "".join(rng.choice(list("ACGT"), size=333))}
    fa = tmp_path / "g.fa"
    with open(fa, "w", encoding="utf-8") as fp:
        for k, v in seqs.items():
            fp.write(f">{k} some description\n")
            for i in range(0, len(v), 60):
                fp.write(v[i:i + 60] + "\n")
    return str(fa), seqs


def test_genome_bed_and_fasta_readers(tmp_path):
    rng = np.random.default_rng(2)
    fa, seqs = _write_genome(tmp_path, rng)
    with predictUtils.Genome(fa) as g:
        assert g.references == ["chrA", "chrB"] and g.nreferences == 2
        assert g.get_reference_length("chrB") == 333
        for (c, s, e) in [("chrA", 0, 10), ("chrA", 55, 130), ("chrB", 300, 333),
                          ("chrA", 990, 1200), ("chrB", 59, 61), ("chrA", 119, 121)]:
            assert g.fetch(c, s, e) == seqs[c][s:e]
    bed = tmp_path / "r.bed"
    bed.write_text("chrA\t100\t120\tpeak\t0\t+\nchrB\t50\t70\n")
    r = predictUtils.BedReader(str(bed), fa, 5)
    assert r.numPredictions == 2 and r.curSequence == seqs["chrA"][95:125]
    assert (r.curChrom, r.curStart, r.curEnd) == ("chrA", 100, 120)
    r.pop()
    assert r.curSequence == seqs["chrB"][45:75]
    f = predictUtils.FastaReader(fa)
    assert f.numPredictions == 2 and f.curLabel == "chrA some description"
    assert f.curSequence == seqs["chrA"]
    f.pop()
    assert f.curSequence == seqs["chrB"] and f.curLabel == "chrB some description"
    # PISA: one query per base, window starting `padding` before it
    gen = interpretUtils.PisaBedGenerator(str(bed), fa, inputLength=40, outputLength=10)
    assert gen.numRegions == 40
    gen.construct()
    qs = list(gen)
    gen.done()
    assert len(qs) == 40 and qs[0].passData == ("chrA", 100) and qs[20].passData == ("chrB", 50)
    assert utils.oneHotDecode(qs[3].sequence) == seqs["chrA"][103 - 15:103 - 15 + 40]
    fg = interpretUtils.FlatBedGenerator(str(bed), fa, inputLength=40, outputLength=20)
    fg.construct()
    q0 = next(fg)
    assert q0.passData == ("chrA", 90, 130) and utils.oneHotDecode(q0.sequence) == seqs["chrA"][90:130]
    fq = interpretUtils.FastaGenerator(fa)
    fq.construct()
    recs = list(fq)
    assert [q.passData for q in recs] == ["chrA some description", "chrB some description"]
    assert utils.oneHotDecode(recs[1].sequence) == seqs["chrB"]


def test_prediction_file_layout(tmp_path):
    """predictUtils.py:149-231, 276-304."""
    rng = np.random.default_rng(3)
    fa, _ = _write_genome(tmp_path, rng)
    bed = tmp_path / "r.bed"
    bed.write_text("chrA\t100\t120\nchrB\t50\t70\nchrA\t500\t520\n")
    out = str(tmp_path / "pred.h5")
    w = predictUtils.H5Writer(out, 2, 3, bedFname=str(bed), genomeFname=fa, config={"x": 1},
                              spillBytes=64)  # forces the memory-mapped scratch path
    lg = [rng.standard_normal((3, 20, 2)).astype(np.float32),
          rng.standard_normal((3, 20, 1)).astype(np.float32)]
    lc = rng.standard_normal((3, 2)).astype(np.float32)
    w.addEntry(([lg[0][0], lg[1][0], float(lc[0, 0]), float(lc[0, 1])], "first"))
    w.addBatch([lg[0][1:], lg[1][1:]], [lc[1:, 0], lc[1:, 1:2]], ["second", "third"])
    w.close()
    assert sorted(os.listdir(tmp_path)) == ["g.fa", "pred.h5", "r.bed"]  # scratch removed
    f = h5lite.File(out)
    assert sorted(f.keys()) == ["chrom_names", "chrom_sizes", "coords_chrom", "coords_start",
                                "coords_stop", "descriptions", "head_0", "head_1", "metadata"]
    assert f["head_0/logits"].shape == (3, 20, 2) and f["head_0/logits"].dtype == np.float32
    assert np.array_equal(f["head_0/logits"][...], lg[0])
    assert np.array_equal(f["head_1/logits"][...], lg[1])
    assert np.array_equal(f["head_1/logcounts"][...], lc[:, 1])
    assert list(f["descriptions"][...]) == ["first", "second", "third"]
    assert list(f["chrom_names"][...]) == ["chrA", "chrB"]
    assert list(f["chrom_sizes"][...]) == [1000, 333] and f["chrom_sizes"].dtype == np.uint64
    assert list(f["coords_chrom"][...]) == [0, 1, 0] and f["coords_chrom"].dtype == np.uint8
    assert list(f["coords_start"][...]) == [100, 50, 500]
    assert list(f["coords_stop"][...]) == [120, 70, 520]
    md = f["metadata"].attrs
    assert md["config"] == "{'x': 1}" and "bpreveal_version" in md and "created_date" in md


def test_attribution_file_layouts(tmp_path):
    """interpretUtils.py:425-602 (flat) and 605-764 (PISA)."""
    rng = np.random.default_rng(4)
    fa, _ = _write_genome(tmp_path, rng)
    n, Lin, rf, nshuf = 5, 30, 21, 4
    res = [interpretUtils.Result(np.float32(i), rng.standard_normal(nshuf).astype(np.float32),
                                 rng.integers(0, 2, (Lin, 4)).astype(np.uint8),
                                 rng.standard_normal((Lin, 4)).astype(np.float32),
                                 ("chrB", 10 + i), i) for i in range(n)]
    out = str(tmp_path / "pisa.h5")
    s = interpretUtils.PisaH5Saver(out, n, nshuf, rf, genome=fa, config="cfg")
    s.construct()
    for r in reversed(res):  # the reference gives no ordering guarantee to the saver
        s.add(r)
    s.done()
    f = h5lite.File(out)
    assert f["shap"].shape == (n, rf, 4) and f["shap"].dtype == np.float16
    assert f["sequence"].dtype == np.uint8
    for r in res:
        assert np.array_equal(f["shap"][...][r.index], r.shap[:rf].astype(np.float16))
        assert np.array_equal(f["sequence"][...][r.index], r.sequence[:rf])
        assert np.array_equal(f["shuffle_predictions"][...][r.index], r.shufflePredictions)
    assert list(f["input_predictions"][...]) == [0, 1, 2, 3, 4]
    assert list(f["coords_base"][...]) == [10, 11, 12, 13, 14] and list(f["coords_chrom"][...]) == [1] * 5
    assert list(f["chrom_names"][...]) == ["chrA", "chrB"]
    # fasta-style: descriptions instead of coordinates
    out2 = str(tmp_path / "flat.h5")
    s2 = interpretUtils.FlatH5Saver(out2, n, Lin, genome=None, config="cfg")
    s2.construct()
    for r in res:
        r.passData = f"seq{r.index}"
        s2.add(r)
    s2.done()
    f2 = h5lite.File(out2)
    assert sorted(f2.keys()) == ["descriptions", "hyp_scores", "input_predictions", "input_seqs",
                                 "metadata"]
    assert f2["hyp_scores"].shape == (n, Lin, 4) and f2["hyp_scores"].dtype == np.float16
    assert list(f2["descriptions"][...]) == [f"seq{i}" for i in range(n)]
    assert np.array_equal(f2["input_seqs"][...][2], res[2].sequence)
    # bed-style flat saver
    out3 = str(tmp_path / "flat_bed.h5")
    s3 = interpretUtils.FlatH5Saver(out3, n, Lin, genome=fa, config="cfg")
    s3.construct()
    for r in res:
        r.passData = ("chrA", 100 + r.index, 130 + r.index)
        s3.add(r)
    s3.done()
    f3 = h5lite.File(out3)
    assert list(f3["coords_start"][...]) == [100, 101, 102, 103, 104]
    assert list(f3["coords_end"][...]) == [130, 131, 132, 133, 134]


def test_training_file_round_trip(tmp_path):
    """tools/make_synthetic_data.py writes prepareTrainingData's layout (gzip, chunks of 128)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "msd", os.path.join(os.path.dirname(GOLD), "..", "tools", "make_synthetic_data.py"))
    msd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(msd)
    out = str(tmp_path / "train.h5")
    msd.write(out, 300, 52, 20, 4, [2, 1], seed=5)
    f = h5lite.File(out)
    assert f["sequence"].shape == (300, 60, 4) and f["sequence"].dtype == np.uint8
    assert f["head_0"].shape == (300, 28, 2) and f["head_1"].shape == (300, 28, 1)
    x = f["sequence"][...]
    assert np.array_equal(x.sum(axis=2), np.ones((300, 60)))
    assert (f["head_1"][...].sum(axis=(1, 2)) > 0).all()
