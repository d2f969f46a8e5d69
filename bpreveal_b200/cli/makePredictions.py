"""makePredictions (src/makePredictions.py:128-250): predict every region of a bed file (with a
genome fasta) or every record of a fasta file; write the prediction HDF5 of
``predictUtils.H5Writer``.  ``num-threads`` is accepted (one process saturates a B200)."""
import json

from bpreveal_b200 import logUtils, predictUtils, utils


def getReader(config):
    """makePredictions.py:128-152."""
    arch = config["settings"]["architecture"]
    if "fasta-file" in config:
        return predictUtils.FastaReader(config["fasta-file"])
    padding = (arch["input-length"] - arch["output-length"]) // 2
    return predictUtils.BedReader(config["bed-file"], config["genome"], padding)


def getWriter(config, numPredictions):
    """makePredictions.py:155-184."""
    st = config["settings"]
    if "fasta-file" in config:
        if "coordinates" in config:
            return predictUtils.H5Writer(st["output-h5"], st["heads"], numPredictions,
                                         bedFname=config["coordinates"]["bed-file"],
                                         genomeFname=config["coordinates"]["genome"],
                                         config=str(config))
        return predictUtils.H5Writer(st["output-h5"], st["heads"], numPredictions,
                                     config=str(config))
    return predictUtils.H5Writer(st["output-h5"], st["heads"], numPredictions,
                                 bedFname=config["bed-file"], genomeFname=config["genome"],
                                 config=str(config))


def main(config):
    logUtils.setVerbosity(config["verbosity"])
    if "genome" in config["settings"]:  # makePredictions.py:193-201: old-format bed json
        config["genome"] = config["settings"].pop("genome")
        logUtils.error("You are using an old-format bed prediction json. The genome argument "
                       "should be moved from \"settings\" to the root of the JSON object. "
                       "Here is a corrected version:")
        logUtils.error(json.dumps(config, indent=4))
    batchSize = config["settings"]["batch-size"]
    modelFname = config["settings"]["architecture"]["model-file"]
    logUtils.info(f"Beginning to make predictions for model {modelFname}")
    reader = getReader(config)
    writer = getWriter(config, reader.numPredictions)
    batcher = utils.BatchPredictor(modelFname, batchSize)
    with batcher:
        for _ in range(reader.numPredictions):
            batcher.submitString(reader.curSequence, reader.curLabel)
            reader.pop()
            while batcher.outputReady():
                writer.addEntry(batcher.getOutput())
        while not batcher.empty():
            writer.addEntry(batcher.getOutput())
    writer.close()
    logUtils.info("Done making predictions, exiting.")


if __name__ == "__main__":
    from bpreveal_b200.cli import run
    run("makePredictions", main)
