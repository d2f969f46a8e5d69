"""Imports the reference's configuration CONTRACTS for the hot-path programs as data.

The JSON-schema files under ``/root/reference/src/schematools`` are the machine-readable form of
the configuration format (doc/bnf/*.bnf is the human-readable one); north_star requires the JSON
configs of trainSoloModel / trainTransformationModel / trainCombinedModel / makePredictions /
interpretPisa / interpretFlat to stay drop-in, so the schemas are taken verbatim -- as data, into
ONE bundle -- rather than restated (a restatement could only drift).  The reference's good / bad
example configurations (schematools/testcases) become test fixtures the same way.

Run here (the reference tree is not on the GPU box):  python tools/import_contracts.py
"""
import glob
import json
import os

REF = "/root/reference/src/schematools"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROGRAMS = ["base", "trainSoloModel", "trainTransformationModel", "trainCombinedModel",
            "makePredictions", "interpretPisa", "interpretFlat", "prepareTrainingData"]


def main():
    schemas = {}
    for name in PROGRAMS:
        with open(os.path.join(REF, name + ".schema"), encoding="utf-8") as fp:
            schemas[name] = json.load(fp)
    out = os.path.join(ROOT, "bpreveal_b200", "contracts", "schemas.json")
    with open(out, "w", encoding="utf-8") as fp:
        json.dump({"source": "mmtrebuchet/bpreveal v5.1.0 src/schematools/*.schema",
                   "schemas": schemas}, fp, indent=1, sort_keys=True)
    cases = {}
    for path in sorted(glob.glob(os.path.join(REF, "testcases", "*.json"))):
        stem = os.path.basename(path)[:-5]
        prog, verdict, _ = stem.rsplit("_", 2)
        with open(path, encoding="utf-8") as fp:
            cases[stem] = {"program": prog, "good": verdict == "good", "config": json.load(fp)}
    out2 = os.path.join(ROOT, "tests", "golden", "schema_cases.json")
    with open(out2, "w", encoding="utf-8") as fp:
        json.dump({"source": "mmtrebuchet/bpreveal v5.1.0 src/schematools/testcases",
                   "cases": cases}, fp, indent=1, sort_keys=True)
    print(f"{len(schemas)} schemas -> {out}\n{len(cases)} cases -> {out2}")


if __name__ == "__main__":
    main()
