"""Validators for the configuration files of the hot-path programs.

Counterpart of the reference's generated ``bpreveal.schema`` (src/schematools/build.py:50-105):
one Draft-7 validator per program, all sharing a registry in which ``/schema/base`` resolves, with
the reference's extra type names.  The schemas themselves are the reference's, imported as data
by ``tools/import_contracts.py`` into ``contracts/schemas.json``.

    from bpreveal_b200 import schema
    schema.trainSoloModel.validate(config)       # raises jsonschema.ValidationError
"""
from __future__ import annotations

import json
import os

from jsonschema import Draft7Validator, validators
from referencing import Registry
from referencing.jsonschema import DRAFT7


def _isNumpyArray(checker, instance):
    import numpy as np
    return isinstance(instance, np.ndarray)


def _isArray(checker, instance):
    return isinstance(instance, (list, tuple))


def _isColorMap(checker, instance):
    return False  # plotting is out of scope here; no colormap object can be passed


_typeChecker = Draft7Validator.TYPE_CHECKER.redefine_many(
    {"ndarray": _isNumpyArray, "colormap": _isColorMap, "array": _isArray})
CustomValidator = validators.extend(Draft7Validator, type_checker=_typeChecker)

with open(os.path.join(os.path.dirname(__file__), "contracts", "schemas.json"),
          encoding="utf-8") as _fp:
    _schemas = json.load(_fp)["schemas"]
for _name, _s in _schemas.items():
    _s["$id"] = "https://example.com/schema/" + _name  # build.py:62-63
_registry = Registry().with_resources(
    [(s["$id"], DRAFT7.create_resource(s)) for s in _schemas.values()])

schemaMap = {name: CustomValidator(s, registry=_registry) for name, s in _schemas.items()}
"""Program name -> validator (``schemaMap["makePredictions"].validate(config)``)."""

base = schemaMap["base"]
trainSoloModel = schemaMap["trainSoloModel"]
trainTransformationModel = schemaMap["trainTransformationModel"]
trainCombinedModel = schemaMap["trainCombinedModel"]
makePredictions = schemaMap["makePredictions"]
interpretPisa = schemaMap["interpretPisa"]
interpretFlat = schemaMap["interpretFlat"]
prepareTrainingData = schemaMap["prepareTrainingData"]
