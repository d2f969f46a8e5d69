"""Host-side training logic (callbacks.py / training.py of the reference) on a scripted model: no
GPU, no kernels."""
import json

import numpy as np
import pytest

from bpreveal_b200 import callbacks as cb
from bpreveal_b200 import training


class ScriptedModel:
    """Feeds ``fit``-style logs from a script of (profile, raw counts) losses per epoch."""

    def __init__(self, heads, script):
        self.heads, self.script = heads, script
        self.stop_training = False
        self.optimizer = type("Opt", (), {"learning_rate": 0.004})()
        self.weights = [np.zeros(1)]
        self.saved = []

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, w):
        self.weights = [x.copy() for x in w]

    def save(self, path):
        self.saved.append((path, float(self.weights[0][0])))

    def fit(self, gen, epochs, validation_data, callbacks, verbose=0):
        del gen, validation_data, verbose
        history = {}
        for c in callbacks:
            c.set_model(self)
            c.on_train_begin({})
        for epoch in range(epochs):
            self.weights[0][0] = epoch
            logs = {}
            tot = 0.0
            for head, (p, craw) in zip(self.heads, self.script[epoch]):
                lam = float(head["INTERNAL_λ-variable"].read_value())
                n = head["head-name"]
                for pre in ("", "val_"):
                    logs[f"{pre}solo_profile_{n}_multinomial_nll"] = p
                    logs[f"{pre}solo_logcounts_{n}_reweightable_mse"] = lam * craw
                tot += head["profile-loss-weight"] * p + lam * craw
            logs["loss"] = logs["val_loss"] = tot * 1.01  # Keras' "perverted" total
            for c in callbacks:
                c.on_epoch_end(epoch, logs)
            for k, v in logs.items():
                history.setdefault(k, []).append(v)
            if self.stop_training:
                break
        for c in callbacks:
            c.on_train_end({})
        return type("History", (), {"history": history})()


def heads_cfg():
    return [{"head-name": "a", "num-tasks": 2, "profile-loss-weight": 2.0,
             "counts-loss-frac-target": 0.1, "counts-loss-weight": 10.0},
            {"head-name": "b", "num-tasks": 1, "profile-loss-weight": 1.0,
             "counts-loss-weight": 3.0}]


def test_build_losses_injects_variables():
    heads = heads_cfg()
    losses, weights = training.buildLosses(heads)
    assert losses == ["multinomial_nll"] * 2 + ["reweightable_mse"] * 2
    assert weights == [2.0, 1.0, 1.0, 1.0]
    assert float(heads[0]["INTERNAL_λ-variable"]) == 10.0
    assert float(heads[1]["INTERNAL_λ-variable"]) == 3.0


def test_fix_loss_sums_metrics_without_profile_weights():
    heads = heads_cfg()
    training.buildLosses(heads)
    logs = {"solo_profile_a_multinomial_nll": 5.0, "solo_logcounts_a_reweightable_mse": 1.0,
            "solo_profile_b_multinomial_nll": 7.0, "solo_logcounts_b_reweightable_mse": 2.0,
            "val_solo_profile_a_multinomial_nll": 50.0, "val_solo_logcounts_a_reweightable_mse": 10.0,
            "val_solo_profile_b_multinomial_nll": 70.0, "val_solo_logcounts_b_reweightable_mse": 20.0,
            "loss": -1.0, "val_loss": -1.0}
    cb.FixLossCallback(heads).on_epoch_end(0, logs)
    assert logs["loss"] == 15.0 and logs["val_loss"] == 150.0


def test_adaptive_lambda_follows_the_reference_recurrence():
    heads = heads_cfg()
    training.buildLosses(heads)
    # constant profile loss 100 and raw counts loss 0.5 for head a: lambda' = t p / (c (1 - t))
    script = [[(100.0, 0.5), (50.0, 1.0)]] * 6
    model = ScriptedModel(heads, script)
    hist = training.trainModel(model, None, None, epochs=6, earlyStop=20, outputPrefix="/tmp/x",
                               plateauPatience=20, heads=heads)
    target = 0.1 * 100.0 / (0.5 * 0.9)  # 22.22
    lam = hist.history["counts-loss-weight"]["a"]
    assert lam[0] == 10.0                       # weight in use during epoch 0
    assert lam[1] == pytest.approx(target)      # epochs 0 and 1 jump (within the factor-10 clamp)
    assert lam[2] == pytest.approx(target)
    exp = target
    for e in range(3, 6):                       # exponential damping, aggression 0.3 (fixed point)
        exp = exp * 0.7 + target * 0.3
        assert lam[e] == pytest.approx(exp)
    assert hist.history["counts-loss-weight"]["b"] == [3.0] * 6   # no target: never touched
    # the corrected losses are what the history holds
    assert hist.history["loss"][0] == pytest.approx(100.0 + 10.0 * 0.5 + 50.0 + 3.0 * 1.0)


def test_lambda_change_is_clamped():
    heads = heads_cfg()
    heads[0]["counts-loss-weight"] = 0.001
    training.buildLosses(heads)
    model = ScriptedModel(heads, [[(100.0, 0.5), (50.0, 1.0)]] * 4)
    hist = training.trainModel(model, None, None, 4, 20, "/tmp/x", 20, heads)
    lam = hist.history["counts-loss-weight"]["a"]
    assert lam[1] == pytest.approx(0.01) and lam[2] == pytest.approx(0.1)   # x10 in epochs 0, 1
    assert lam[3] == pytest.approx(0.2)                                      # x2 afterwards


def test_early_stopping_checkpoint_and_plateau():
    heads = [{"head-name": "a", "num-tasks": 1, "profile-loss-weight": 1.0,
              "counts-loss-weight": 1.0}]
    training.buildLosses(heads)
    prof = [10.0, 9.0, 8.0, 8.5, 8.4, 8.3, 8.2, 8.1]
    model = ScriptedModel(heads, [[(p, 0.0)] for p in prof])
    hist = training.trainModel(model, None, None, 8, earlyStop=4, outputPrefix="/tmp/m",
                               plateauPatience=2, heads=heads, checkpointSuffix=".checkpoint.npz")
    # best epoch 2; epochs 3..6 do not improve -> stop after epoch 6, best weights restored
    assert len(hist.history["val_loss"]) == 7
    assert model.weights[0][0] == 2
    assert [e for _, e in model.saved] == [0.0, 1.0, 2.0]
    assert model.saved[0][0] == "/tmp/m.checkpoint.npz"
    # plateau patience 2: halved after epochs 4 and 6
    assert hist.history["learning_rate"] == [0.004] * 5 + [0.002] * 2
    assert model.optimizer.learning_rate == pytest.approx(0.001)


def test_history_json(tmp_path):
    heads = heads_cfg()
    training.buildLosses(heads)
    model = ScriptedModel(heads, [[(100.0, 0.5), (50.0, 1.0)]] * 2)
    hist = training.trainModel(model, None, None, 2, 5, str(tmp_path / "m"), 5, heads)
    cfg = {"settings": {"output-prefix": str(tmp_path / "m")}, "heads": heads}
    name = training.saveHistory(hist, cfg)
    data = json.load(open(name, encoding="utf-8"))
    assert data["counts-loss-weight"]["a"][0] == 10.0
    assert data["config"]["heads"][0]["INTERNAL_λ-variable"] == {"λ_a": pytest.approx(22.2222, rel=1e-4)}
    assert len(data["val_loss"]) == 2


def test_display_callback_speaks_the_progress_protocol():
    """callbacks.py:392-716: lines ``∬row∬1∬window∬text`` -- an epoch table at INFO with the
    reference's row order (Epoch, losses, per-head profile then counts metrics, early-stop /
    plateau state, timers), value columns 11 wide, the previous epoch beside the current one, the
    counts-loss weights at DEBUG, batch lines at most once a second plus first and last batch."""
    import logging
    from bpreveal_b200 import logUtils
    heads = heads_cfg()
    training.buildLosses(heads)
    fix, es, ck, pl, ad, disp = cb.getCallbacks(3, "/tmp/nowhere", 2, heads, [0] * 5, [0] * 2)
    assert isinstance(disp, cb.DisplayCallback) and disp.numBatches == 5
    seen = []

    class Grab(logging.Handler):
        def emit(self, record):
            seen.append((record.levelno, record.getMessage()))

    lg = logging.getLogger("BPReveal")
    h, old = Grab(), lg.level
    lg.addHandler(h)
    logUtils.setVerbosity("DEBUG")
    try:
        model = ScriptedModel(heads, [[(10.0, 2.0), (5.0, 1.0)], [(9.0, 1.5), (4.5, 0.8)]])
        ck.on_epoch_end = lambda *a, **k: None  # no files in this test
        disp.params = {"epochs": 2}
        for c in (fix, es, pl, ad, disp):
            c.set_model(model)
            c.on_train_begin({})
        names = ["loss", "solo_profile_a_multinomial_nll", "solo_logcounts_a_reweightable_mse",
                 "solo_profile_b_multinomial_nll", "solo_logcounts_b_reweightable_mse",
                 "solo_profile_a_loss", "solo_logcounts_a_loss"]
        for epoch in range(2):
            disp.on_epoch_begin(epoch)
            for b in range(5):
                disp.on_train_batch_end(b, {n: 1.0 + b for n in names})
            for b in range(2):
                disp.on_test_batch_end(b, {})
            logs = {}
            for head, (p, craw) in zip(heads, model.script[epoch]):
                lam = float(head["INTERNAL_λ-variable"].read_value())
                for pre in ("", "val_"):
                    logs[f"{pre}solo_profile_{head['head-name']}_multinomial_nll"] = p
                    logs[f"{pre}solo_logcounts_{head['head-name']}_reweightable_mse"] = lam * craw
            logs["loss"] = logs["val_loss"] = 30.0 - epoch
            for c in (fix, es, pl, ad, disp):
                c.on_epoch_end(epoch, logs)
    finally:
        lg.removeHandler(h)
        lg.setLevel(old)
    lines = [m for _, m in seen if m.startswith("∬")]
    parsed = [m.split("∬")[1:] for m in lines]            # [row, "1", window, text]
    assert all(len(p) == 4 and p[1] == "1" for p in parsed)
    wins = [p[2] for p in parsed]
    assert {"E", "λ", "B", "BH", "V"} <= set(wins)
    # batch lines: first (BH) and last batch of each epoch at least, never the ignored *_loss keys
    btext = [p[3] for p in parsed if p[2] in ("B", "BH")]
    assert not any("solo_profile_a_loss" in t for t in btext)
    assert any(t.strip().startswith("batch") and "0 /    4" in t for t in btext)
    # epoch table: row order and the 11-wide value column
    e_rows = {p[3].split()[0] if not p[3].strip().startswith(("Best", "Epochs", "Minutes",
                                                                "Seconds", "Setup", "Training",
                                                                "Validation"))
              else " ".join(p[3].split()[:-1]): int(p[0]) for p in parsed if p[2] == "E"}
    assert e_rows["Epoch"] == 2 and e_rows["loss"] == 4 and e_rows["val_loss"] == 5
    assert e_rows["solo_profile_a_multinomial_nll"] == 7
    assert e_rows["val_solo_profile_a_multinomial_nll"] == 8
    assert e_rows["solo_profile_b_multinomial_nll"] == 10
    assert e_rows["solo_logcounts_a_reweightable_mse"] == 13
    epoch_lines = [p[3] for p in parsed if p[2] == "E" and p[3].strip().startswith("Epoch ")]
    assert epoch_lines[0].endswith("   0 /    2") and "   1 /    2" in epoch_lines[1]
    assert epoch_lines[1].endswith("   0 /    2")            # previous epoch's column
    assert any(lv == logging.INFO and "∬E∬" in m for lv, m in seen)
    assert all(lv == logging.DEBUG for lv, m in seen if "∬λ∬" in m or "∬B∬" in m)
    lam_lines = [p for p in parsed if p[2] == "λ"]
    assert lam_lines[0][3].strip().startswith("a") and int(lam_lines[0][0]) == 2
    assert cb.DisplayCallback.formatStr(3.14159) == "      3.142"
    assert cb.DisplayCallback.formatStr((3, 17)) == "   3 /   17"
    assert cb.DisplayCallback.formatStr("x") == "          x"
