"""The training driver of the reference (training.py:22-176) for the ``models.Model`` duck-type:
``buildLosses`` (injects the adaptive counts-loss variable into every head and returns the Keras
``loss`` / ``loss_weights`` lists), ``trainModel`` (callbacks + fit + λ history) and the history
file of ``trainWithGenerators`` (``<output-prefix>.history.json``).  Host logic only.
"""
from __future__ import annotations

import json

import numpy as np

from .callbacks import LambdaVariable, getCallbacks


def buildLosses(heads):
    """Returns ``(losses, loss_weights)`` ordered profile heads then counts heads
    (training.py:22-65): profile weights from the configuration, ones for the counts heads (their
    weight λ lives inside the loss).  The loss entries are names -- the kernels implement exactly
    these two losses -- so ``model.compile(loss=losses, loss_weights=weights)`` reads like the
    reference's call (trainSoloModel.py:183-185)."""
    profileLosses, countsLosses, profileWeights, countsWeights = [], [], [], []
    for head in heads:
        profileLosses.append("multinomial_nll")
        profileWeights.append(head["profile-loss-weight"])
        lamInit = head["counts-loss-weight"] if "counts-loss-weight" in head else 1
        head["INTERNAL_λ-variable"] = LambdaVariable(lamInit, f"λ_{head['head-name']}")
        countsWeights.append(1.0)
        countsLosses.append("reweightable_mse")
    return profileLosses + countsLosses, profileWeights + countsWeights


def _makeJsonSerializable(cfg):
    """History / configuration objects -> something ``json`` accepts (training.py:68-102)."""
    if isinstance(cfg, dict):
        return {str(k): _makeJsonSerializable(v) for k, v in cfg.items()}
    if isinstance(cfg, (bool, int, float, str)) or cfg is None:
        return cfg
    if isinstance(cfg, (list, tuple, np.ndarray)):
        return [_makeJsonSerializable(x) for x in cfg]
    if isinstance(cfg, LambdaVariable):
        return {cfg.name: float(cfg.read_value())}
    try:
        return float(cfg)
    except (TypeError, ValueError):
        return str(cfg)


def trainModel(model, trainBatchGen, valBatchGen, epochs, earlyStop, outputPrefix,
               plateauPatience, heads, checkpointSuffix=".checkpoint.keras"):
    """Callbacks, the training loop, and the counts-loss-weight history (training.py:143-176)."""
    callbacks = getCallbacks(earlyStop, outputPrefix, plateauPatience, heads, trainBatchGen,
                             valBatchGen, checkpointSuffix=checkpointSuffix)
    history = model.fit(trainBatchGen, epochs=epochs, validation_data=valBatchGen,
                        callbacks=callbacks, verbose=0)
    history.history["counts-loss-weight"] = callbacks[4].λHistory  # ApplyAdaptiveCountsLoss
    return history


def trainWithGenerators(model, config, inputLength, outputLength, device=None):
    """training.py:104-141: open ``train-data`` / ``val-data``, build the batch generators,
    train, write ``<output-prefix>.history.json``.  ``config["heads"]`` must already carry the
    λ variables of ``buildLosses``."""
    from . import h5lite, logUtils
    from .generators import H5BatchGenerator
    model.summary(print_fn=logUtils.debug)
    logUtils.info("Loading data into generators.")
    st = config["settings"]
    with h5lite.File(config["train-data"]) as trainH5, h5lite.File(config["val-data"]) as valH5:
        trainGenerator = H5BatchGenerator(config["heads"], trainH5, inputLength, outputLength,
                                          st["max-jitter"], st["batch-size"], device=device)
        valGenerator = H5BatchGenerator(config["heads"], valH5, inputLength, outputLength,
                                        st["max-jitter"], st["batch-size"], device=device)
    logUtils.info("Generators initialized. Training.")
    history = trainModel(model, trainGenerator, valGenerator, st["epochs"],
                         st["early-stopping-patience"], st["output-prefix"],
                         st["learning-rate-plateau-patience"], config["heads"])
    logUtils.info("Saving history.")
    saveHistory(history, config)
    return history


def saveHistory(history, config, historyName=None):
    """``<output-prefix>.history.json`` as trainWithGenerators writes it (training.py:134-140)."""
    historyName = historyName or f"{config['settings']['output-prefix']}.history.json"
    hist = dict(history.history)
    hist["config"] = config
    with open(historyName, "w", encoding="utf-8") as fp:
        json.dump(_makeJsonSerializable(hist), fp, ensure_ascii=False, indent=4)
    return historyName
