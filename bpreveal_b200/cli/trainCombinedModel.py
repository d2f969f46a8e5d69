"""trainCombinedModel (src/trainCombinedModel.py:60-108): frozen transformation (bias) model +
trainable residual BPNet; writes ``<output-prefix>_combined.keras`` and
``<output-prefix>_residual.keras``."""
from bpreveal_b200 import logUtils, models, training


def main(config, device=None):
    logUtils.setVerbosity(config["verbosity"])
    st = config["settings"]
    arch = st["architecture"]
    inputLength, outputLength = arch["input-length"], arch["output-length"]
    regressionModel = models.loadModel(st["transformation-model"]["transformation-model-file"],
                                       device=device)
    regressionModel.trainable = False
    logUtils.debug("Loaded regression model.")
    combinedModel, residualModel, _ = models.combinedModel(
        inputLength, outputLength, arch["filters"], arch["layers"], arch["input-filter-width"],
        arch["output-filter-width"], config["heads"], regressionModel)
    losses, lossWeights = training.buildLosses(config["heads"])
    for m in (residualModel, combinedModel):
        m.compile(optimizer=models.Adam(learning_rate=st["learning-rate"]), loss=losses,
                  loss_weights=lossWeights, metrics=losses)
    training.trainWithGenerators(combinedModel, config, inputLength, outputLength, device=device)
    combinedModel.save(st["output-prefix"] + "_combined.keras")
    residualModel.save(st["output-prefix"] + "_residual.keras")
    logUtils.info("Training job completed successfully.")
    return combinedModel, residualModel


if __name__ == "__main__":
    from bpreveal_b200.cli import run
    run("trainCombinedModel", main)
