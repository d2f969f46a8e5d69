"""trainSoloModel (src/trainSoloModel.py:167-196): build a solo BPNet from the configuration,
train it on the train / val HDF5 files, write ``<output-prefix>.keras``,
``<output-prefix>.checkpoint.keras`` and ``<output-prefix>.history.json``."""
from bpreveal_b200 import logUtils, models, training


def main(config, device=None):
    logUtils.setVerbosity(config["verbosity"])
    arch = config["settings"]["architecture"]
    inputLength, outputLength = arch["input-length"], arch["output-length"]
    model = models.soloModel(inputLength, outputLength, arch["filters"], arch["layers"],
                             arch["input-filter-width"], arch["output-filter-width"],
                             config["heads"], "solo", device=device)
    logUtils.debug("Model built.")
    losses, lossWeights = training.buildLosses(config["heads"])
    model.compile(optimizer=models.Adam(learning_rate=config["settings"]["learning-rate"]),
                  loss=losses, loss_weights=lossWeights, metrics=losses)
    training.trainWithGenerators(model, config, inputLength, outputLength, device=device)
    model.save(config["settings"]["output-prefix"] + ".keras")
    logUtils.info("Training job completed successfully.")
    return model


if __name__ == "__main__":
    from bpreveal_b200.cli import run
    run("trainSoloModel", main)
