"""Config-driven entry points with the reference's program names and JSON formats:
``python -m bpreveal_b200.cli.<program> config.json`` for trainSoloModel,
trainTransformationModel, trainCombinedModel, makePredictions, interpretPisa, interpretFlat
(src/<program>.py of the reference).  Each module has ``main(config)`` like its counterpart."""
import sys


def loadConfig(fname):
    """The reference reads configurations with its expression interpreter
    (internal/interpreter.py:419-425; interpretPisa.py:260, interpretFlat.py:317), of which plain
    JSON is a special case -- ``bpreveal_b200.interpreter`` is that language."""
    from bpreveal_b200 import interpreter
    return interpreter.evalFile(fname)


def run(program, main):
    """``if __name__ == "__main__"`` body shared by the programs: read, validate, run."""
    from bpreveal_b200 import schema
    if len(sys.argv) != 2:
        sys.exit(f"usage: {program} <config.json>")
    config = loadConfig(sys.argv[1])
    assert isinstance(config, dict)
    schema.schemaMap[program].validate(config)
    main(config)
