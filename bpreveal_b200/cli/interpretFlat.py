"""interpretFlat (src/interpretFlat.py:187-312): hypothetical contribution scores of the profile
and counts metrics of one head for every region of a bed file or record of a fasta file."""
from bpreveal_b200 import interpretUtils, logUtils, predictUtils


def profileMetric(headID, taskIDs):
    """interpretFlat.py:187-237."""
    return interpretUtils.ProfileMetric(headID, tuple(taskIDs))


def countsMetric(numHeads, headID):
    """interpretFlat.py:240-250."""
    return interpretUtils.CountsMetric(numHeads, headID)


def main(config):
    logUtils.setVerbosity(config["verbosity"])
    genomeFname = None
    kmerSize = config.get("kmer-size", 1)
    if "bed-file" in config:
        genomeFname = config["genome"]
        generator = interpretUtils.FlatBedGenerator(
            bedFname=config["bed-file"], genomeFname=genomeFname,
            inputLength=config["input-length"], outputLength=config["output-length"])
    else:
        generator = interpretUtils.FastaGenerator(config["fasta-file"])
    savers = [interpretUtils.FlatH5Saver(outputFname=config[k], numSamples=generator.numRegions,
                                         inputLength=config["input-length"], genome=genomeFname,
                                         config=str(config))
              for k in ("profile-h5", "counts-h5")]
    if "coordinates" in config:  # interpretFlat.py:303-311 (added to the writers before saving)
        regions = predictUtils.readBed(config["coordinates"]["bed-file"])
        for s in savers:
            s.coordinates = (regions, config["coordinates"]["genome"])
    metrics = [profileMetric(config["head-id"], config["profile-task-ids"]),
               countsMetric(config["heads"], config["head-id"])]
    batcher = interpretUtils.InterpRunner(
        modelFname=config["model-file"], metrics=metrics, batchSize=10, generator=generator,
        savers=savers, numShuffles=config["num-shuffles"], kmerSize=kmerSize, numThreads=2,
        backend="shap", useHypotheticalContribs=True)
    batcher.run()
    logUtils.info("interpretFlat complete, exiting.")


if __name__ == "__main__":
    from bpreveal_b200.cli import run
    run("interpretFlat", main)
