"""Training callbacks for the ``models.Model`` duck-type.

Restates the behaviour of the reference's ``callbacks.py`` -- ``FixLossCallback`` (:12-93),
``ApplyAdaptiveCountsLoss`` (:96-389), ``getCallbacks`` (:718-777) -- and of the three Keras
callbacks it configures (``EarlyStopping(monitor="val_loss", mode="min",
restore_best_weights=True)``, ``ModelCheckpoint(save_best_only=True)``,
``ReduceLROnPlateau(factor=0.5)``; Keras 3 semantics recalled from its source, not in the tree).
``DisplayCallback`` (:392-716) emits the log-line protocol that ``showTrainingProgress`` renders.
Pure host logic: no kernels here.
"""
from __future__ import annotations

import re
import time

import numpy as np


class LambdaVariable:
    """Stand-in for the ``tf.Variable`` the reference stores as ``INTERNAL_λ-variable``
    (training.py:46-52): ``assign`` / ``read_value`` / ``numpy`` / ``float()``."""

    def __init__(self, value, name="counts-loss-weight"):
        self.value = float(value)
        self.name = name

    def assign(self, value):
        self.value = float(value)
        return self

    def read_value(self):
        return self.value

    def numpy(self):
        return np.float32(self.value)

    def __float__(self):
        return self.value

    def __repr__(self):
        return f"LambdaVariable({self.value!r})"


class Callback:
    """Keras ``Callback`` surface used by ``Model.fit``."""

    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    # per-batch hooks: ``Model.fit`` calls them only on callbacks that define them, with a lazy
    # ``logs`` mapping (reading a value synchronises with the device)


def _head_loss_names(logs, headName):
    """(profile, counts, val_profile, val_counts) log keys of one head, by the reference's
    regexes; validation names are matched first because the training patterns match them too
    (callbacks.py:56-75)."""
    pats = (("vp", fr"val.*profile_{re.escape(headName)}_multinomial_nll"),
            ("vc", fr"val.*logcounts_{re.escape(headName)}_reweightable_mse"),
            ("p", fr".*profile_{re.escape(headName)}_multinomial_nll"),
            ("c", fr".*logcounts_{re.escape(headName)}_reweightable_mse"))
    found = {}
    for name in logs:
        for key, pat in pats:
            if re.fullmatch(pat, name):
                found[key] = name
                break
    return found.get("p"), found.get("c"), found.get("vp"), found.get("vc")


class FixLossCallback(Callback):
    """``loss`` / ``val_loss`` := sum of the per-head metrics (WITHOUT the profile loss weights),
    which is what every later callback monitors (callbacks.py:12-93).  Must come first."""

    def __init__(self, heads):
        self.heads = heads

    def correctLosses(self, logs):
        total = totalVal = 0.0
        for head in self.heads:
            p, c, vp, vc = _head_loss_names(logs, head["head-name"])
            total += sum(logs[k] for k in (p, c) if k is not None)
            totalVal += sum(logs[k] for k in (vp, vc) if k is not None)
        if total != 0:
            logs["loss"] = total
        if totalVal != 0:
            logs["val_loss"] = totalVal

    def on_epoch_end(self, epoch, logs=None):
        assert logs is not None, "Cannot work with empty logs!"
        self.correctLosses(logs)


class EarlyStopping(Callback):
    """Keras 3 ``EarlyStopping`` for ``mode="min"``: an epoch improves when
    ``monitor < best - min_delta``; after ``patience`` epochs without improvement training stops
    (never on epoch 0); the best weights are put back at the end of training."""

    def __init__(self, monitor="val_loss", patience=0, verbose=0, mode="min", min_delta=0.0,
                 restore_best_weights=False):
        assert mode == "min", "only mode='min' is used on this path (callbacks.py:748-752)"
        del verbose
        self.monitor, self.patience, self.min_delta = monitor, int(patience), abs(min_delta)
        self.restore_best_weights = restore_best_weights
        self.on_train_begin()

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.stopped_epoch = 0
        self.best = np.inf
        self.best_weights = None
        self.best_epoch = 0

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model.get_weights()
            self.best_epoch = epoch
        self.wait += 1
        if current < self.best - self.min_delta:
            self.best = current
            self.best_epoch = epoch
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
            self.wait = 0
            return
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True

    def on_train_end(self, logs=None):
        if self.restore_best_weights and self.best_weights is not None:
            self.model.set_weights(self.best_weights)


class ModelCheckpoint(Callback):
    """Keras ``ModelCheckpoint(save_best_only=True, mode="min")``: saves the whole model whenever
    the monitored value improves."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=True, mode="min"):
        assert mode == "min"
        del verbose
        self.filepath, self.monitor, self.save_best_only = filepath, monitor, save_best_only
        self.best = np.inf
        self.saved_epochs = []

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if self.save_best_only:
            if current is None or not current < self.best:
                return
            self.best = current
        self.model.save(self.filepath)
        self.saved_epochs.append(epoch)


class ReduceLROnPlateau(Callback):
    """Keras ``ReduceLROnPlateau`` (mode min): ``lr *= factor`` after ``patience`` epochs without
    ``monitor < best - min_delta``; logs ``learning_rate`` every epoch."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, min_delta=1e-4,
                 cooldown=0, min_lr=0.0):
        assert factor < 1.0
        del verbose
        self.monitor, self.factor, self.patience = monitor, float(factor), int(patience)
        self.min_delta, self.cooldown, self.min_lr = float(min_delta), int(cooldown), float(min_lr)
        self.on_train_begin()

    def on_train_begin(self, logs=None):
        self.best = np.inf
        self.wait = 0
        self.cooldown_counter = 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        opt = self.model.optimizer
        logs["learning_rate"] = float(opt.learning_rate)
        current = logs.get(self.monitor)
        if current is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if current < self.best - self.min_delta:
            self.best = current
            self.wait = 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = float(opt.learning_rate)
                if old > self.min_lr:
                    opt.learning_rate = max(old * self.factor, self.min_lr)
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class ApplyAdaptiveCountsLoss(Callback):
    """The adaptive counts-loss algorithm (callbacks.py:96-389).  After every epoch, for each head
    with a ``counts-loss-frac-target`` t, the counts weight moves towards
    ``λ' = t·p / (c_raw·(1 − t))`` (p = profile loss, c_raw = counts loss / current λ): a jump in
    epochs 0-1 (validation losses on epoch 0), exponential damping with ``aggression`` afterwards,
    change clamped to a factor 10 (epochs 0-1) or 2; then the ``best`` of the other callbacks is
    re-expressed with the new weights so a larger λ is not mistaken for a worse model."""

    def __init__(self, heads, aggression, lrPlateauCallback, earlyStopCallback,
                 checkpointCallback):
        self.heads = heads
        self.aggression = aggression
        self.lrPlateauCallback = lrPlateauCallback
        self.earlyStopCallback = earlyStopCallback
        self.checkpointCallback = checkpointCallback
        self.logsHistory = []
        self.λHistory = {head["head-name"]: [] for head in heads}

    def on_train_begin(self, logs=None):
        for head in self.heads:
            if "counts-loss-frac-target" in head:
                if "counts-loss-weight" in head:
                    lam = head["counts-loss-weight"]
                else:
                    # BPNet heuristic with the reference's empirical constant (callbacks.py:176-181)
                    lam = head["INTERNAL_mean-counts"] * head["counts-loss-frac-target"] * 37.0
                head["INTERNAL_λ-variable"].assign(lam)

    def getLosses(self, epoch, headName):
        logs = self.logsHistory[epoch]
        names = _head_loss_names(logs, headName)
        if None in names:
            raise ValueError(f"The loss didn't match any regex: head {headName}, "
                             f"log names {list(logs)}")
        return tuple(logs[n] for n in names)

    def whatWouldValLossBe(self, epoch):
        total = 0.0
        for head in self.heads:
            _, _, vpl, vcl = self.getLosses(epoch, head["head-name"])
            total += vpl * float(head["profile-loss-weight"])
            cur = float(head["INTERNAL_λ-variable"].read_value())
            total += vcl * cur / self.λHistory[head["head-name"]][epoch]
        return total

    def resetCallbacks(self):
        corrected = self.whatWouldValLossBe(self.earlyStopCallback.best_epoch)
        self.lrPlateauCallback.best = corrected
        self.earlyStopCallback.best = corrected
        self.checkpointCallback.best = corrected

    def on_epoch_end(self, epoch, logs=None):
        assert logs is not None, "Cannot work with empty logs!"
        self.logsHistory.append(logs)
        for head in self.heads:
            name = head["head-name"]
            if epoch > 0:
                profileLoss, countsLoss, _, _ = self.getLosses(epoch, name)
            else:  # the first batches of epoch 0 dominate its training losses
                _, _, profileLoss, countsLoss = self.getLosses(epoch, name)
            var = head["INTERNAL_λ-variable"]
            cur = float(var.read_value())
            self.λHistory[name].append(cur)
            if "counts-loss-frac-target" not in head:
                continue
            target = head["counts-loss-frac-target"]
            corrected = target * profileLoss / ((countsLoss / cur) * (1 - target))
            new = cur * (1 - self.aggression) + corrected * self.aggression if epoch > 1 \
                else corrected
            threshold = 2 if epoch > 1 else 10
            if max(new, cur) / min(new, cur) > threshold:
                new = cur * threshold if new > cur else cur / threshold
            var.assign(new)
        self.resetCallbacks()


class DisplayCallback(Callback):
    """The training log in the line protocol of the reference's display program
    (callbacks.py:392-716; rendered by showTrainingProgress): every line is

        ``∬<row>∬1∬<window>∬<text>``

    with window ``E`` (epoch table, INFO level: name, this epoch's value, the previous epoch's),
    ``λ`` (counts-loss weights, DEBUG), ``B`` / ``BH`` (running batch losses, DEBUG, at most once a
    second plus the first and last batch) and ``V`` (validation batch counter, DEBUG); ``<text>``
    is the row's fields laid out at their columns, each value 11 characters wide.  Row order:
    Epoch, total and validation loss, per head the profile then the counts losses (train, val),
    then early-stopping / plateau state, learning rate and the timers."""

    SPACER = "EPOCH_SPACER"

    def __init__(self, plateauCallback, earlyStopCallback, adaptiveLossCallback, trainBatchGen,
                 valBatchGen):
        self.plateauCallback = plateauCallback
        self.earlyStopCallback = earlyStopCallback
        self.adaptiveLossCallback = adaptiveLossCallback
        self.numBatches = len(trainBatchGen) if trainBatchGen is not None else 0
        self.numValBatches = len(valBatchGen) if valBatchGen is not None else 0
        self.params = {}
        self.numEpochs = 0
        self.epochNumber = self.batchNumber = 0
        self.rowsEpoch, self.rowsBatch = {}, {}
        self.ignoreMetrics = []
        self.maxLen = 0
        self.prevEpochLogs = None
        self.trainBeginTime = self.lastEpochEndTime = self.curEpochWaitTime = None
        self.lastBatchEndTime = self.lastValBatchEndTime = 0.0
        self.firstBatchTime = self.lastBatchTime = 0.0
        self.firstValBatchTime = self.lastValBatchTime = 0.0

    # -- layout ---------------------------------------------------------------------------
    def _calcPositions(self, names):
        """Rows of both tables from the first batch's metric names (callbacks.py:455-530)."""
        prof, cnt, unknown = [], [], []
        for name in names:
            found = False
            for head in self.adaptiveLossCallback.heads:
                hn = re.escape(head["head-name"])
                if re.fullmatch(fr".*profile_{hn}_.*multinomial_nll", name):
                    prof += [name, "val_" + name, self.SPACER]
                elif re.fullmatch(fr".*logcounts_{hn}_.*reweightable_mse", name):
                    cnt += [name, "val_" + name, self.SPACER]
                elif re.fullmatch(fr".*profile_{hn}_.*loss", name) or \
                        re.fullmatch(fr".*logcounts_{hn}_.*loss", name):
                    # the same number under its output's name: shown once, descriptively
                    self.ignoreMetrics += [name, "val_" + name]
                else:
                    continue
                found = True
                break
            if not found and name not in ("lr", "learning_rate", "epoch", "batch", "loss"):
                unknown.append(name)
        if unknown:
            raise ValueError(f"{unknown} includes unknown loss component, not {prof} / {cnt}")
        epochOrder = ["Epoch", self.SPACER, "loss", "val_loss", self.SPACER] + prof + cnt + [
            "Best epoch", "Best loss", "Epochs until earlystop", "lr", "Epochs until LR plateau",
            self.SPACER, "Setup time", "Training time", "Validation time", "Seconds / epoch",
            "Minutes to earlystop", "Minutes elapsed"]
        batchOrder = [n for n in ["batch", "loss"] + prof + cnt if n != self.SPACER]
        self.rowsEpoch = {n: i + 2 for i, n in enumerate(epochOrder)}  # spacers take a row
        self.rowsBatch = {n: i + 2 for i, n in enumerate(batchOrder)}
        self.maxLen = max(len(n) for n in epochOrder)

    @staticmethod
    def formatStr(val):
        """Eleven characters: right-aligned text, integer, %.3f, or ``a / b`` for a pair."""
        if isinstance(val, str):
            return f"{val:>11s}"
        if isinstance(val, (bool, np.bool_)):
            return f"{str(val):>11s}"
        if isinstance(val, (int, np.integer)):
            return f"{int(val):>11d}"
        if isinstance(val, (float, np.floating)):
            return f"{float(val):>11.3f}"
        if isinstance(val, tuple) and len(val) == 2:
            if all(isinstance(v, (int, np.integer)) for v in val):
                return f"{int(val[0]):>4d} / {int(val[1]):>4d}"
            if val[1] is None:
                return f"{str(val):>11s}"
        return "     FMT_ERR"

    def _emit(self, rows, writer, window):
        """rows: [(row, [(column, text), ...]), ...] -> one protocol line per row."""
        for row, fields in rows:
            width = max(col + len(text) for col, text in fields)
            chars = [" "] * width
            for col, text in fields:
                chars[col:col + len(text)] = text
            writer(f"∬{row}∬1∬{window}∬{''.join(chars)}")

    # -- Keras hooks ----------------------------------------------------------------------
    def on_train_begin(self, logs=None):
        self.trainBeginTime = time.perf_counter()
        total = self.params.get("epochs", self.params.get("nb_epochs"))
        if total is not None:
            self.numEpochs = total

    def on_epoch_begin(self, epoch, logs=None):
        now = time.perf_counter()
        self.epochNumber, self.batchNumber = epoch, 0
        self.firstBatchTime = self.firstValBatchTime = now + 1e10
        if self.lastEpochEndTime is not None:
            self.curEpochWaitTime = now - self.lastEpochEndTime

    def on_train_batch_end(self, batch, logs=None):
        logs = logs if logs is not None else {}
        now = time.perf_counter()
        self.firstBatchTime = min(now, self.firstBatchTime)
        self.lastBatchTime = now
        if not self.rowsEpoch:
            self._calcPositions(list(logs.keys()))
        if now - self.lastBatchEndTime > 1.0 or batch in (0, self.numBatches - 1):
            from . import logUtils
            self.lastBatchEndTime = now
            self.batchNumber = batch
            vals = {k: logs[k] for k in logs.keys()}  # (reads the running means back)
            vals["batch"] = (batch, self.numBatches - 1)
            rows = [(self.rowsBatch[k], [(1, k), (self.maxLen + 2, self.formatStr(v))])
                    for k, v in vals.items() if k not in self.ignoreMetrics and k in self.rowsBatch]
            self._emit(rows, logUtils.debug, "BH" if batch == 0 else "B")

    def on_test_batch_end(self, batch, logs=None):
        now = time.perf_counter()
        self.firstValBatchTime = min(now, self.firstValBatchTime)
        self.lastValBatchTime = now
        if now - self.lastValBatchEndTime > 1.0 or batch in (0, self.numValBatches - 1):
            from . import logUtils
            self.lastValBatchEndTime = now
            row = max(self.rowsBatch.values()) if self.rowsBatch else 2
            self._emit([(row, [(1, "Eval batch"),
                               (20, self.formatStr((batch, self.numValBatches - 1)))])],
                       logUtils.debug, "V")

    def on_epoch_end(self, epoch, logs=None):
        from . import logUtils
        if logs is None:
            logUtils.warning("Received empty logs at epoch end.")
            logs = {}
        logs = {k: (float(v) if isinstance(v, (np.floating, np.integer)) else v)
                for k, v in logs.items()}
        if not self.rowsEpoch:
            self._calcPositions([k for k in logs if not k.startswith("val_")])
        es, pl = self.earlyStopCallback, self.plateauCallback
        now = time.perf_counter()
        logs["Best epoch"] = int(getattr(es, "best_epoch", 0))
        logs["Best loss"] = float(es.best)
        logs["Epochs until earlystop"] = (int(es.wait), int(es.patience))
        logs["Epochs until LR plateau"] = (int(pl.wait), int(pl.patience))
        if self.lastEpochEndTime is not None:
            per = now - self.lastEpochEndTime
            logs["Seconds / epoch"] = per
            logs["Minutes to earlystop"] = per * (es.patience - es.wait) / 60
        if self.curEpochWaitTime is not None:
            logs["Setup time"] = self.curEpochWaitTime
        logs["Training time"] = self.lastBatchTime - self.firstBatchTime
        logs["Validation time"] = self.lastValBatchTime - self.firstValBatchTime
        self.lastEpochEndTime = now
        logs["Minutes elapsed"] = (now - (self.trainBeginTime or now)) / 60
        logs["Epoch"] = (self.epochNumber, self.numEpochs)
        lr = getattr(getattr(self.model, "optimizer", None), "learning_rate", None)
        logs["lr"] = f"{float(lr):10.7f}" if lr is not None else "n/a"
        rows = []
        for k, v in logs.items():
            if k in self.ignoreMetrics or k == "learning_rate" or k not in self.rowsEpoch:
                continue
            fields = [(1, k), (self.maxLen + 2, self.formatStr(v))]
            if self.prevEpochLogs is not None:
                fields.append((self.maxLen + 16, self.formatStr(self.prevEpochLogs.get(k, ""))))
            rows.append((self.rowsEpoch[k], fields))
        self._emit(rows, logUtils.info, "E")
        lam = self.adaptiveLossCallback.λHistory
        self._emit([(i + 2, [(1, name), (20, f"{lam[name][-1]:8.3f}")])
                    for i, name in enumerate(lam.keys()) if lam[name]], logUtils.debug, "λ")
        self.prevEpochLogs = logs


def getCallbacks(earlyStop, outputPrefix, plateauPatience, heads, trainBatchGen=None,
                 valBatchGen=None, checkpointSuffix=".checkpoint.keras"):
    """The reference's callback set in its order (callbacks.py:718-777): fix-loss, early stop,
    checkpoint, LR plateau, adaptive counts loss, display."""
    fixLoss = FixLossCallback(heads)
    earlyStopCb = EarlyStopping(monitor="val_loss", patience=earlyStop, mode="min",
                                restore_best_weights=True)
    checkpoint = ModelCheckpoint(f"{outputPrefix}{checkpointSuffix}", monitor="val_loss",
                                 save_best_only=True, mode="min")
    plateau = ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=plateauPatience)
    adaptive = ApplyAdaptiveCountsLoss(heads, 0.3, plateau, earlyStopCb, checkpoint)
    display = DisplayCallback(plateau, earlyStopCb, adaptive, trainBatchGen, valBatchGen)
    return fixLoss, earlyStopCb, checkpoint, plateau, adaptive, display
