"""interpretPisa (src/interpretPisa.py:153-264): DeepSHAP attribution of the leftmost output
logit for every base of a bed file or every record of a fasta file."""
from bpreveal_b200 import interpretUtils, logUtils


def pisaMetric(headID, taskID):
    """interpretPisa.py:153-166."""
    return interpretUtils.PisaMetric(headID, taskID)


def main(config, shard=None):
    logUtils.setVerbosity(config["verbosity"])
    receptiveField = config["input-length"] - config["output-length"]
    if config.get("correct-receptive-field", True):  # interpretPisa.py:175-191
        receptiveField += 1
    kmerSize = config.get("kmer-size", 1)
    numThreads = config.get("num-threads", 1)
    if "fasta-file" in config or "sequence-fasta" in config:
        if "sequence-fasta" in config:
            logUtils.error("DEPRECATION: You are referring to the fasta file in your PISA JSON as "
                           "sequence-fasta. This is deprecated, please change the parameter name "
                           "to fasta-file.")
            config["fasta-file"] = config["sequence-fasta"]
        generator = interpretUtils.FastaGenerator(config["fasta-file"])
        genome = None
    else:
        generator = interpretUtils.PisaBedGenerator(
            bedFname=config["bed-file"], genomeFname=config["genome"],
            inputLength=config["input-length"], outputLength=config["output-length"])
        genome = config["genome"]
    writer = interpretUtils.PisaH5Saver(
        outputFname=config["output-h5"], numSamples=generator.numRegions,
        numShuffles=config["num-shuffles"], receptiveField=receptiveField, genome=genome,
        config=str(config))
    metric = pisaMetric(config["head-id"], config["task-id"])
    batcher = interpretUtils.InterpRunner(
        modelFname=config["model-file"], metrics=[metric], batchSize=10, generator=generator,
        savers=[writer], numShuffles=config["num-shuffles"], kmerSize=kmerSize,
        numThreads=numThreads, backend="shap", useHypotheticalContribs=False, shard=shard)
    batcher.run()
    logUtils.info("PISA interpretation complete, exiting.")


if __name__ == "__main__":
    from bpreveal_b200.cli import run
    run("interpretPisa", main)
