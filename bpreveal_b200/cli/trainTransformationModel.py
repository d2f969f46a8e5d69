"""trainTransformationModel (src/trainTransformationModel.py): regress the frozen solo (bias)
model's outputs onto the experimental data; writes ``<output-prefix>.keras``."""
from bpreveal_b200 import logUtils, models, training


def main(config, device=None):
    logUtils.setVerbosity(config["verbosity"])
    st = config["settings"]
    inputLength, outputLength = st["input-length"], st["output-length"]
    soloModel = models.loadModel(st["solo-model-file"], device=device)
    model = models.transformationModel(soloModel, st["profile-architecture"],
                                       st["counts-architecture"], config["heads"])
    losses, lossWeights = training.buildLosses(config["heads"])
    model.compile(optimizer=models.Adam(learning_rate=st["learning-rate"]), loss=losses,
                  loss_weights=lossWeights, metrics=losses)
    training.trainWithGenerators(model, config, inputLength, outputLength, device=device)
    model.save(st["output-prefix"] + ".keras")
    logUtils.info("Training job completed successfully.")
    return model


if __name__ == "__main__":
    from bpreveal_b200.cli import run
    run("trainTransformationModel", main)
