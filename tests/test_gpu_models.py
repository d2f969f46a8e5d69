"""GPU: the bpreveal.models drop-in builders (soloModel / transformationModel / combinedModel)
against the oracle restatement of models.py:53-387 and layers.py:10-75."""
import numpy as np
import pytest
import torch

from oracle import bpnet_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

HEADS = [{"head-name": "a", "num-tasks": 2, "profile-loss-weight": 1.0, "counts-loss-weight": 2.0},
         {"head-name": "b", "num-tasks": 1, "profile-loss-weight": 1.5, "counts-loss-weight": 3.0,
          "use-bias-counts": True}]


def solo_state(numFilters, numLayers, wi, wo, seed):
    return O.init_solo_params(numFilters, numLayers, wi, wo, [2, 1], seed=seed, biasScale=0.1)


def load_solo(model, params):
    model.net.load_state(params)


class ArrayGen:
    """Minimal keras.utils.Sequence look-alike (generators.py:77-90)."""

    def __init__(self, x, counts, tasks, batch):
        self.x, self.counts, self.tasks, self.batch = x, counts, tasks, batch

    def __len__(self):
        return int(np.ceil(self.x.shape[0] / self.batch))

    def __getitem__(self, i):
        s = slice(i * self.batch, min((i + 1) * self.batch, self.x.shape[0]))
        vals, cnts, k = [], [], 0
        for t in self.tasks:
            v = self.counts[s, :, k:k + t]
            vals.append(v)
            cnts.append(np.log(v.sum(axis=(1, 2))))
            k += t
        return self.x[s], tuple(vals + cnts)


def test_solo_model_duck_type_and_predict(dev):
    from bpreveal_b200 import models
    m = models.soloModel(52, 20, 8, 3, 3, 3, HEADS, "solo", precise=True, device=dev)
    assert m.name == "solo_model" and m.inputs[0].shape == (None, 52, 4)
    assert m.input.shape[1] == 52
    assert [o.shape for o in m.outputs] == [(None, 20, 2), (None, 20, 1), (None, 1), (None, 1)]
    assert m.output_names == ["solo_profile_a", "solo_profile_b", "solo_logcounts_a",
                              "solo_logcounts_b"]
    names = [k for k, _ in m.named_weights()]
    assert names[:4] == ["solo_initial_conv/kernel", "solo_initial_conv/bias",
                         "solo_conv_0/kernel", "solo_conv_0/bias"]
    with pytest.raises(ValueError):
        models.soloModel(60, 20, 8, 3, 3, 3, HEADS, "bad", device=dev)  # inconsistent lengths
    params = solo_state(8, 3, 3, 3, seed=5)
    load_solo(m, params)
    x = O.synthetic_sequences(5, 52, seed=1)
    outs = m.predict(x, batch_size=2)
    ref_l, ref_c = Hh.oracle_forward(params, x)
    assert [o.shape for o in outs] == [(5, 20, 2), (5, 20, 1), (5, 1), (5, 1)]
    got_l = np.concatenate(outs[:2], axis=2)
    got_c = np.concatenate(outs[2:], axis=1)
    assert Hh.rel_err(got_l, ref_l) < 1e-4 and Hh.rel_err(got_c, ref_c) < 1e-4
    lines = []
    m.summary(print_fn=lines.append)
    assert any("solo_profile_a/kernel" in ln for ln in lines)
    assert m.count_params() == sum(v.size for v in params.values())


def test_solo_fit_history_and_save_roundtrip(dev, tmp_path):
    from bpreveal_b200 import models
    m = models.soloModel(52, 20, 16, 3, 3, 3, HEADS, "solo", device=dev, seed=7)
    m.compile(optimizer=models.Adam(learning_rate=0.004), loss=None,
              loss_weights=[1.0, 1.5, 1.0, 1.0])
    x = O.synthetic_sequences(24, 52, seed=2)
    c = O.synthetic_profiles(24, 20, 3, seed=2, meanReads=40.0)
    gen, val = ArrayGen(x, c, [2, 1], 8), ArrayGen(x[:8], c[:8], [2, 1], 8)
    seen = []

    class Stop:
        def set_model(self, model):
            self.model = model

        def on_epoch_end(self, epoch, logs):
            seen.append((epoch, logs["loss"], logs["val_loss"]))
            if epoch == 3:
                self.model.stop_training = True

    hist = m.fit(gen, epochs=10, validation_data=val, callbacks=[Stop()], verbose=0)
    assert len(hist.history["loss"]) == 4 and len(seen) == 4  # stopped by the callback
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    for k in ("solo_profile_a_loss", "solo_logcounts_b_loss", "val_solo_profile_b_loss"):
        assert k in hist.history
    path = str(tmp_path / "m.npz")
    m.save(path)
    m2 = models.loadModel(path, device=dev)
    a, b = m.predict(x[:4]), m2.predict(x[:4])
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    # Keras formats, written and read without h5py (keras_io + h5lite)
    for ext in (".keras", ".h5"):
        kp = str(tmp_path / ("m" + ext))
        m.save(kp)
        m3 = models.loadModel(kp, device=dev)
        assert m3.name == "solo_model" and m3.output_names == m.output_names
        for u, v in zip(a, m3.predict(x[:4])):
            assert np.array_equal(u, v)
    # utils.py:67-72: a configuration still naming "<x>.model" finds "<x>.keras"
    m4 = models.loadModel(str(tmp_path / "m.model"), device=dev)
    assert np.array_equal(m4.predict(x[:4])[0], a[0])


def test_transformation_model_forward_and_training(dev):
    from bpreveal_b200 import models
    spec_p = {"name": "simple", "types": ["linear", "sigmoid"]}
    spec_c = {"name": "simple", "types": ["linear"]}
    solo = models.soloModel(52, 20, 8, 3, 3, 3, HEADS, "solo", precise=True, device=dev)
    params = solo_state(8, 3, 3, 3, seed=5)
    load_solo(solo, params)
    tm = models.transformationModel(solo, spec_p, spec_c, HEADS)
    assert solo.trainable is False and tm.name == "transformation_model"
    # non-trivial coefficients through the Keras-named weight interface
    rng = np.random.default_rng(3)
    tp = O.init_transform_params(2, spec_p, spec_c)
    for k in tp:
        tp[k] = (tp[k] + rng.standard_normal(2) * 0.2).astype(np.float32)
    named = dict(tm.named_weights())
    for h, n in enumerate(["a", "b"]):
        named[f"regress_linear_profile_{n}_conv/kernel"] = tp[f"profile{h}_linear"][0].reshape(1, 1, 1)
        named[f"regress_linear_profile_{n}_conv/bias"] = tp[f"profile{h}_linear"][1].reshape(1)
        named[f"sigmoid_in_linear_profile_{n}_conv/kernel"] = tp[f"profile{h}_sigmoid_in"][0].reshape(1, 1, 1)
        named[f"sigmoid_in_linear_profile_{n}_conv/bias"] = tp[f"profile{h}_sigmoid_in"][1].reshape(1)
        named[f"sigmoid_out_linear_profile_{n}_conv/kernel"] = tp[f"profile{h}_sigmoid_out"][0].reshape(1, 1, 1)
        named[f"sigmoid_out_linear_profile_{n}_conv/bias"] = tp[f"profile{h}_sigmoid_out"][1].reshape(1)
        named[f"regress_linear_logcounts_{n}_conv/kernel"] = tp[f"logcounts{h}_linear"][0].reshape(1, 1, 1)
        named[f"regress_linear_logcounts_{n}_conv/bias"] = tp[f"logcounts{h}_linear"][1].reshape(1)
    tm.set_weights(named)
    x = O.synthetic_sequences(4, 52, seed=9)
    p64 = O.to_torch(params, torch.float64)
    tpt = {k: torch.tensor(v, dtype=torch.float64) for k, v in tp.items()}
    po, co = O.transformation_forward(torch.tensor(x, dtype=torch.float64), p64, tpt, spec_p, spec_c)
    outs = tm.predict(x)
    assert Hh.rel_err(np.concatenate(outs[:2], axis=2), torch.cat(po, dim=2).numpy()) < 1e-4
    assert Hh.rel_err(np.concatenate(outs[2:], axis=1), torch.cat(co, dim=1).numpy()) < 1e-4
    # training moves only the regression scalars and lowers the loss
    tm.compile(optimizer=models.Adam(learning_rate=0.05), loss_weights=[1.0, 1.0, 1.0, 1.0])
    c = O.synthetic_profiles(16, 20, 3, seed=4, meanReads=40.0)
    xs = O.synthetic_sequences(16, 52, seed=4)
    before = [v.copy() for v in solo.get_weights()]
    hist = tm.fit(ArrayGen(xs, c, [2, 1], 8), epochs=6, shuffle=False)
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    for u, v in zip(before, solo.get_weights()):
        assert np.array_equal(u, v)


def test_transformation_gradients_match_oracle(dev):
    """d loss / d (m, b) of every regression against torch autograd through the oracle."""
    from bpreveal_b200 import models
    spec_p = {"name": "simple", "types": ["linear", "relu"]}
    spec_c = {"name": "simple", "types": ["sigmoid"]}
    heads = [dict(h) for h in HEADS]
    solo = models.soloModel(52, 20, 8, 3, 3, 3, heads, "solo", precise=True, device=dev)
    params = solo_state(8, 3, 3, 3, seed=11)
    load_solo(solo, params)
    tm = models.transformationModel(solo, spec_p, spec_c, heads)
    tm.compile(optimizer=models.Adam(learning_rate=0.0), loss_weights=[1.0, 1.5, 1.0, 1.0])
    rng = np.random.default_rng(1)
    c0 = tm.coeffs.cpu().numpy() + rng.standard_normal(tm.coeffs.shape).astype(np.float32) * 0.1
    tm.coeffs.copy_(torch.tensor(c0))
    x = O.synthetic_sequences(6, 52, seed=3)
    counts = O.synthetic_profiles(6, 20, 3, seed=3, meanReads=30.0)
    # oracle with the same scalars
    tp = {}
    for h in range(2):
        lo, _ = tm._rows(h, "p")
        tp[f"profile{h}_linear"] = torch.tensor(c0[lo, :2], dtype=torch.float64, requires_grad=True)
        tp[f"profile{h}_relu_out"] = torch.tensor(c0[lo + 1, :2], dtype=torch.float64, requires_grad=True)
        tp[f"profile{h}_relu_in"] = torch.tensor(c0[lo + 1, 2:], dtype=torch.float64, requires_grad=True)
        lo, _ = tm._rows(h, "c")
        tp[f"logcounts{h}_sigmoid_out"] = torch.tensor(c0[lo, :2], dtype=torch.float64, requires_grad=True)
        tp[f"logcounts{h}_sigmoid_in"] = torch.tensor(c0[lo, 2:], dtype=torch.float64, requires_grad=True)
    p64 = O.to_torch(params, torch.float64)
    po, co = O.transformation_forward(torch.tensor(x, dtype=torch.float64), p64, tp, spec_p, spec_c)
    ct = torch.tensor(counts, dtype=torch.float64)
    trueP = [ct[:, :, 0:2], ct[:, :, 2:3]]
    trueC = [torch.log(t.sum(dim=(1, 2))) for t in trueP]
    tot, pl, cl = O.total_loss(po, co, trueP, trueC, [1.0, 1.5], [2.0, 3.0])
    tot.backward()
    # device: run one "training" step with lr 0 and read the Adam first moment = 0.1 * grad
    losses = tm._step(x, counts, train=True)
    got = tm._adam["m"].cpu().numpy() / 0.1
    np.testing.assert_allclose(losses[:2], [v.item() for v in pl], rtol=1e-4)
    for h in range(2):
        lo, _ = tm._rows(h, "p")
        np.testing.assert_allclose(got[lo, :2], tp[f"profile{h}_linear"].grad.numpy(), rtol=2e-3, atol=1e-4)
        np.testing.assert_allclose(got[lo + 1, :2], tp[f"profile{h}_relu_out"].grad.numpy(), rtol=2e-3, atol=1e-4)
        np.testing.assert_allclose(got[lo + 1, 2:], tp[f"profile{h}_relu_in"].grad.numpy(), rtol=2e-3, atol=1e-4)
        lo, _ = tm._rows(h, "c")
        np.testing.assert_allclose(got[lo, :2], tp[f"logcounts{h}_sigmoid_out"].grad.numpy(), rtol=2e-3, atol=1e-5)
        np.testing.assert_allclose(got[lo, 2:], tp[f"logcounts{h}_sigmoid_in"].grad.numpy(), rtol=2e-3, atol=1e-5)


def test_combined_model_matches_oracle_and_trains_residual_only(dev):
    from bpreveal_b200 import models
    vec = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden",
                                            "oracle_vectors.npz"))
    spec_p = {"name": "simple", "types": ["linear", "sigmoid"]}
    spec_c = {"name": "simple", "types": ["linear"]}
    heads = [dict(HEADS[0], **{"use-bias-counts": True}), dict(HEADS[1], **{"use-bias-counts": False})]
    solo = models.soloModel(52, 20, 8, 3, 3, 3, heads, "solo", precise=True, device=dev)
    load_solo(solo, {k[2:]: vec[k] for k in vec.files if k.startswith("p_")})
    tm = models.transformationModel(solo, spec_p, spec_c, heads)
    named = dict(tm.named_weights())
    for h, n in enumerate(["a", "b"]):
        for key, (kern, bias) in {
                f"profile{h}_linear": (f"regress_linear_profile_{n}_conv/kernel", f"regress_linear_profile_{n}_conv/bias"),
                f"profile{h}_sigmoid_in": (f"sigmoid_in_linear_profile_{n}_conv/kernel", f"sigmoid_in_linear_profile_{n}_conv/bias"),
                f"profile{h}_sigmoid_out": (f"sigmoid_out_linear_profile_{n}_conv/kernel", f"sigmoid_out_linear_profile_{n}_conv/bias"),
                f"logcounts{h}_linear": (f"regress_linear_logcounts_{n}_conv/kernel", f"regress_linear_logcounts_{n}_conv/bias")}.items():
            named[kern] = vec[f"t_{key}"][0].reshape(1, 1, 1)
            named[bias] = vec[f"t_{key}"][1].reshape(1)
    tm.set_weights(named)
    xr = vec["xr"]
    comb, resid, biasHeads = models.combinedModel(xr.shape[1], 20, 8, 4, 3, 3, heads, tm,
                                                  precise=True)
    assert biasHeads is tm and tm.trainable is False and comb.name == "combined_model"
    assert comb.output_names[0] == "combined_add_profile_a" and comb.cropFlank == (xr.shape[1] - 52) // 2
    resid.net.load_state({k[2:]: vec[k] for k in vec.files if k.startswith("r_")})
    outs = comb.predict(xr)
    assert Hh.rel_err(np.concatenate(outs[:2], axis=2), vec["comb_logits"]) < 1e-4
    assert Hh.rel_err(np.concatenate(outs[2:], axis=1), vec["comb_logcounts"]) < 1e-4
    # training: bias model untouched, residual moves, loss falls
    comb.compile(optimizer=models.Adam(learning_rate=0.004), loss_weights=[1.0, 1.0, 1.0, 1.0])
    c = O.synthetic_profiles(xr.shape[0], 20, 3, seed=8, meanReads=40.0)
    bias_before = [v.copy() for v in tm.get_weights()]
    resid_before = [v.copy() for v in resid.get_weights()]
    hist = comb.fit(ArrayGen(xr, c, [2, 1], xr.shape[0]), epochs=8, shuffle=False)
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    for u, v in zip(bias_before, tm.get_weights()):
        assert np.array_equal(u, v)
    assert any(not np.array_equal(u, v) for u, v in zip(resid_before, resid.get_weights()))


def test_device_batch_generator_matches_reference_semantics(dev):
    """generators.py:57-140 + libslide.c:44-91: region shuffle, jitter crop, log-count targets,
    mean counts -- against the oracle's restatement (itself pinned to the compiled libslide)."""
    from bpreveal_b200 import models
    from bpreveal_b200.generators import DeviceBatchGenerator
    rng = np.random.default_rng(0)
    N, J, Lin, Lout = 37, 4, 52, 20
    heads = [dict(h) for h in HEADS]
    seq = O.synthetic_sequences(N, Lin + 2 * J, seed=4)
    prof = [rng.poisson(2.0, size=(N, Lout + 2 * J, t)).astype(np.float32) + 0.5 for t in (2, 1)]
    gen = DeviceBatchGenerator(heads, seq, prof, Lin, Lout, J, batchSize=8, device=dev)
    assert len(gen) == 5
    want_mean = [p[:, J:-J, :].sum() / N for p in prof]
    for h, m in zip(heads, want_mean):
        assert abs(h["INTERNAL_mean-counts"] - m) < 1e-3 * m
    # replay the reference's generator arithmetic on the host for two epochs
    plan_rng = np.random.default_rng(seed=1234)
    reg = np.arange(N)
    for epoch in range(2):
        plan_rng.shuffle(reg)
        cols = plan_rng.integers(0, 2 * J, size=(N,), dtype=np.int32)
        want_seq = O.slide(seq, np.zeros((N, Lin, 4), dtype=np.uint8), reg, cols)
        want_vals = [O.slide(p, np.zeros((N, Lout, p.shape[2]), dtype=np.float32), reg, cols)
                     for p in prof]
        got_rows = 0
        for bi in range(len(gen)):
            x, y = gen[bi]
            s = bi * 8
            e = s + x.shape[0]
            assert np.array_equal(x.cpu().numpy(), want_seq[s:e])
            for h in range(2):
                assert np.array_equal(y[h].cpu().numpy(), want_vals[h][s:e])
                np.testing.assert_allclose(y[2 + h].cpu().numpy(),
                                           np.log(want_vals[h][s:e].sum(axis=(1, 2))), rtol=1e-5)
            got_rows += x.shape[0]
        assert got_rows == N and x.shape[0] == N - 4 * 8  # ragged last batch
        gen.on_epoch_end()
    # and it drives fit() without leaving the device
    m = models.soloModel(Lin, Lout, 16, 3, 3, 3, heads, "solo", device=dev, seed=3)
    m.compile(optimizer=models.Adam(learning_rate=0.004), loss_weights=[1.0, 1.0, 1.0, 1.0])
    hist = m.fit(gen, epochs=3)
    assert hist.history["loss"][-1] < hist.history["loss"][0]


def test_train_model_with_the_reference_callback_set(dev, tmp_path):
    """training.trainModel = buildLosses + getCallbacks + fit, as trainSoloModel.py:174-190 drives
    it: adaptive counts-loss weights reach the kernels, the logs carry the metric names the
    reference's callbacks match, the checkpoint and the history file are written."""
    import copy
    import json
    from bpreveal_b200 import models, training
    heads = copy.deepcopy(HEADS)
    heads[0]["counts-loss-frac-target"] = 0.2
    losses, weights = training.buildLosses(heads)
    m = models.soloModel(52, 20, 16, 3, 3, 3, heads, "solo", device=dev, seed=7)
    m.compile(optimizer=models.Adam(learning_rate=0.004), loss=losses, loss_weights=weights)
    x = O.synthetic_sequences(32, 52, seed=3)
    c = O.synthetic_profiles(32, 20, 3, seed=3, meanReads=40.0)
    gen, val = ArrayGen(x, c, [2, 1], 8), ArrayGen(x[:16], c[:16], [2, 1], 8)
    prefix = str(tmp_path / "solo")
    hist = training.trainModel(m, gen, val, epochs=5, earlyStop=10, outputPrefix=prefix,
                               plateauPatience=10, heads=heads,
                               checkpointSuffix=".checkpoint.npz")
    h = hist.history
    for k in ("solo_profile_a_multinomial_nll", "val_solo_logcounts_b_reweightable_mse",
              "learning_rate", "loss", "val_loss"):
        assert len(h[k]) == 5, k
    # FixLossCallback: loss = sum of the metrics, without the profile weight 1.5 of head b
    e = 4
    tot = sum(h[f"solo_profile_{n}_multinomial_nll"][e] + h[f"solo_logcounts_{n}_reweightable_mse"][e]
              for n in "ab")
    assert h["loss"][e] == pytest.approx(tot, rel=1e-12)
    lam = h["counts-loss-weight"]
    assert lam["a"][0] == 2.0 and lam["b"] == [3.0] * 5
    # after epoch 0 the weight jumps to t p / (c_raw (1 - t)) computed from the validation losses
    p0, c0 = h["val_solo_profile_a_multinomial_nll"][0], h["val_solo_logcounts_a_reweightable_mse"][0]
    want = 0.2 * p0 / ((c0 / 2.0) * 0.8)
    want = min(max(want, 2.0 / 10), 2.0 * 10)
    assert lam["a"][1] == pytest.approx(want, rel=1e-6)
    # ... and the kernels used it: the device-side weight is the one of the last epoch
    assert float(m.net.lam[0]) == pytest.approx(lam["a"][4], rel=1e-6)
    m2 = models.loadModel(prefix + ".checkpoint.npz", device=dev)
    assert m2.predict(x[:2])[0].shape == (2, 20, 2)
    name = training.saveHistory(hist, {"settings": {"output-prefix": prefix}, "heads": heads})
    assert json.load(open(name, encoding="utf-8"))["counts-loss-weight"]["a"][0] == 2.0


def test_prediction_side_kernels(dev):
    """bpb_onehot_encode == utils.oneHotEncode (docstring matrix utils.py:456-464, case folding, N ->
    zeros, allowN = False raises) and bpb_logits_to_profile == utils.logitsToProfile per head."""
    from bpreveal_b200 import ops
    seq = "ACGTTTacgtNnRx"
    b = torch.tensor(list(seq.encode("ascii")), dtype=torch.uint8, device=dev)
    want = np.array([[c.upper() == a for a in "ACGT"] for c in seq], dtype=np.uint8)
    got = ops.onehot_encode(b, allowN=True).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(got[:6], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1],
                                    [0, 0, 0, 1], [0, 0, 0, 1]])
    with pytest.raises(AssertionError):
        ops.onehot_encode(b)
    ops.onehot_encode(b[:10])  # pure ACGT passes the strict check
    big = torch.randint(0, 256, (3, 1000), dtype=torch.uint8, device=dev)
    gb = ops.onehot_encode(big, allowN=True).cpu().numpy()
    up = np.char.upper(big.cpu().numpy().astype(np.uint8).view("S1"))
    for i, a in enumerate(b"ACGT"):
        assert np.array_equal(gb[..., i], (up == bytes([a])).astype(np.uint8))
    # logitsToProfile for a batch with two heads of 2 and 1 tasks
    g = torch.Generator().manual_seed(1)
    logits = (torch.randn((5, 37, 3), generator=g) * 3).to(dev)
    lc = torch.randn((5, 2), generator=g).to(dev)
    toff = torch.tensor([0, 2, 3], dtype=torch.int32, device=dev)
    prof = ops.logits_to_profile(toff, logits, lc).cpu().numpy().astype(np.float64)
    lg, lcn = logits.cpu().numpy().astype(np.float64), lc.cpu().numpy().astype(np.float64)
    for bi in range(5):
        for h, (k0, k1) in enumerate(((0, 2), (2, 3))):
            blk = lg[bi, :, k0:k1]
            e = np.exp(blk - blk.max())
            ref = e / e.sum() * np.exp(lcn[bi, h])
            assert np.allclose(prof[bi, :, k0:k1], ref, rtol=2e-5, atol=1e-9)
            assert np.isclose(prof[bi, :, k0:k1].sum(), np.exp(lcn[bi, h]), rtol=1e-5)


def test_streaming_host_prediction_equals_device_prediction(dev):
    """SoloNet.predict_host (H2D / forward / D2H on three streams, two staging slots each way, ragged
    tail) returns exactly what the device-resident path returns, also when called twice."""
    from bpreveal_b200 import models
    m = models.soloModel(52, 20, 16, 3, 3, 3, HEADS, "solo", device=dev, seed=7)
    x = O.synthetic_sequences(8 * 5 + 3, 52, seed=4)
    want_l, want_c = m.net.predict(torch.tensor(x, device=dev), batchSize=8)
    torch.cuda.synchronize()
    for pinned in (True, False):
        xh = torch.tensor(x).pin_memory() if pinned else torch.tensor(x)
        for _ in range(2):
            got_l, got_c = m.net.predict_host(xh, batchSize=8)
            assert got_l.is_pinned() and torch.equal(got_l, want_l.cpu()) and \
                torch.equal(got_c, want_c.cpu())
    outs = m.predict(x, batch_size=8)  # the Keras-style call takes the streaming path for numpy
    assert np.array_equal(outs[0], want_l.cpu().numpy()[:, :, :2])
    assert np.array_equal(outs[3], want_c.cpu().numpy()[:, 1:2])


def test_deepshap_on_transformation_and_combined_models(dev):
    """shap.py:782-793: on transformation / combined graphs the rescale rule also covers Sigmoid,
    Exp and Log (CountsLogSumExp), linear regressions keep ordinary gradients.  PISA, the flat
    counts metric (through the fp64 log-sum-exp on the head with use-bias-counts, identity on the
    other) and the flat profile metric against the oracle's autograd restatement."""
    import os
    from bpreveal_b200 import interpret, models
    vec = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
    spec_p = {"name": "simple", "types": ["linear", "sigmoid"]}
    spec_c = {"name": "simple", "types": ["linear", "relu"]}
    heads = [dict(HEADS[0], **{"use-bias-counts": True}), dict(HEADS[1], **{"use-bias-counts": False})]
    bias_p = {k[2:]: vec[k] for k in vec.files if k.startswith("p_")}
    resid_p = {k[2:]: vec[k] for k in vec.files if k.startswith("r_")}
    solo = models.soloModel(52, 20, 8, 3, 3, 3, heads, "solo", precise=True, device=dev)
    load_solo(solo, bias_p)
    tm = models.transformationModel(solo, spec_p, spec_c, heads)
    rng = np.random.default_rng(11)
    tp = O.init_transform_params(2, spec_p, spec_c)
    for k in tp:
        tp[k] = (tp[k] + rng.standard_normal(2) * 0.3).astype(np.float32)
    named = dict(tm.named_weights())
    for h, n in enumerate(["a", "b"]):
        for key, base in ((f"profile{h}_linear", f"regress_linear_profile_{n}_conv"),
                          (f"profile{h}_sigmoid_in", f"sigmoid_in_linear_profile_{n}_conv"),
                          (f"profile{h}_sigmoid_out", f"sigmoid_out_linear_profile_{n}_conv"),
                          (f"logcounts{h}_linear", f"regress_linear_logcounts_{n}_conv"),
                          (f"logcounts{h}_relu_in", f"relu_in_linear_logcounts_{n}_conv"),
                          (f"logcounts{h}_relu_out", f"relu_out_linear_logcounts_{n}_conv")):
            named[base + "/kernel"] = tp[key][0].reshape(1, 1, 1)
            named[base + "/bias"] = tp[key][1].reshape(1)
    tm.set_weights(named)
    xr = vec["xr"]
    Lr = xr.shape[1]
    comb, resid, _ = models.combinedModel(Lr, 20, 8, 4, 3, 3, heads, tm, precise=True)
    resid.net.load_state(resid_p)
    Q, n = 2, 6
    xs = xr[:Q]
    xt = torch.tensor(xs, device=dev)
    p64b, p64r = O.to_torch(bias_p, torch.float64), O.to_torch(resid_p, torch.float64)
    tpt = {k: torch.tensor(v, dtype=torch.float64) for k, v in tp.items()}
    fwd_c = O.combined_shap_forward(p64r, p64b, tpt, spec_p, spec_c, [True, False], 52)

    def fwd_t(joint):
        return O.transformation_forward(joint, p64b, tpt, spec_p, spec_c, relu=O.rescale_relu,
                                        sigmoid=O.rescale_sigmoid)

    f = (Lr - 52) // 2
    cases = [
        ("combined pisa", comb, xs, fwd_c, lambda: interpret.pisa_model(comb, xt, 0, 1, n),
         lambda pr, cn: O.pisa_metric(pr, cn, 0, 1), False),
        ("combined counts (LSE head)", comb, xs, fwd_c,
         lambda: interpret.flat_model(comb, xt, 0, n, kind="counts"),
         lambda pr, cn: O.counts_metric(pr, cn, 0), True),
        ("combined counts (identity head)", comb, xs, fwd_c,
         lambda: interpret.flat_model(comb, xt, 1, n, kind="counts"),
         lambda pr, cn: O.counts_metric(pr, cn, 1), True),
        ("combined profile", comb, xs, fwd_c,
         lambda: interpret.flat_model(comb, xt, 0, n, kind="profile", taskIDs=[0, 1]),
         lambda pr, cn: O.profile_metric(pr, cn, 0, [0, 1], shap=True), True),
        ("transformation pisa", tm, xs[:, f:Lr - f], fwd_t,
         lambda: interpret.pisa_model(tm, xt[:, f:Lr - f].contiguous(), 1, 0, n),
         lambda pr, cn: O.pisa_metric(pr, cn, 1, 0), False),
        ("transformation counts", tm, xs[:, f:Lr - f], fwd_t,
         lambda: interpret.flat_model(tm, xt[:, f:Lr - f].contiguous(), 0, n, kind="counts"),
         lambda pr, cn: O.counts_metric(pr, cn, 0), True),
    ]
    for name, model, xq, fwd, run, metric, hyp in cases:
        phi, vx, vr = run()
        phi, vx, vr = phi.cpu().numpy(), vx.cpu().numpy(), vr.cpu().numpy()
        for q in range(Q):
            refs = O.generate_shuffles(xq[q], n)
            ref_phi, val, rvals = O.deepshap(xq[q], refs, fwd, metric, hypothetical=hyp)
            scale = np.abs(ref_phi).max()
            err = np.abs(phi[q] - ref_phi).max() / scale
            assert err < 3e-3, f"{name}: attribution rel err {err}"
            assert abs(vx[q] - val) <= 1e-4 * max(1.0, abs(val)), name
            assert np.abs(vr[q] - rvals).max() <= 1e-4 * max(1.0, np.abs(rvals).max()), name
            if not hyp:  # efficiency of the standard combiner
                rhs = vx[q] - vr[q].mean()
                assert abs(phi[q].sum() - rhs) <= 3e-3 * max(abs(rhs), scale), name
