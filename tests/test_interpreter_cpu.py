"""bpreveal_b200/interpreter.py against vectors produced by the reference's own evaluator
(tests/golden/make_interpreter_golden.py imports /root/reference/src/internal/interpreter.py):
same values, same exception types, and plain JSON configurations evaluate to what json.load gives."""
import json
import os

import pytest

from bpreveal_b200 import interpreter

HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = json.load(open(os.path.join(HERE, "golden", "interpreter_vectors.json"), encoding="utf-8"))


def enc(v):
    if isinstance(v, range):
        return list(v)
    if isinstance(v, dict):
        return {"__dict__": [[enc(k), enc(x)] for k, x in v.items()]}
    if isinstance(v, list):
        return [enc(x) for x in v]
    if callable(v):
        return "__callable__"
    return v


@pytest.mark.parametrize("vec", VECTORS, ids=[str(i) for i in range(len(VECTORS))])
def test_matches_the_reference_evaluator(vec):
    if "error" in vec:
        with pytest.raises(BaseException) as info:
            interpreter.evalString(vec["expr"], dict(vec["env"]))
        assert type(info.value).__name__ == vec["error"], vec["expr"]
    else:
        got = enc(interpreter.evalString(vec["expr"], dict(vec["env"])))
        assert got == vec["value"] and json.dumps(got) == json.dumps(vec["value"]), vec["expr"]


def test_plain_json_configurations_are_a_special_case(tmp_path):
    """Every good configuration of the reference's schema corpus (tests/golden/schema_cases.json),
    written out as JSON text, evaluates to what json.loads gives."""
    cases = json.load(open(os.path.join(HERE, "golden", "schema_cases.json"),
                           encoding="utf-8"))["cases"]
    good = [c["config"] for c in cases.values() if c["good"]]
    assert len(good) >= 10
    for i, cfg in enumerate(good):
        p = tmp_path / f"c{i}.json"
        p.write_text(json.dumps(cfg, indent=2))
        assert interpreter.evalFile(str(p)) == cfg


def test_expression_configuration_for_a_program(tmp_path):
    """A configuration written with the language's extras (comprehension, arithmetic, JSON
    constants, a lambda) reaches the program as the plain dictionary it denotes."""
    src = """{"heads": [{"num-tasks": 2, "head-name": "h{0:d}".format(i),
                         "profile-loss-weight": 1, "counts-loss-weight": 10 ** (i + 1)}
                        for i in range(2)],
             "settings": {"epochs": (lambda n: n * 5)(4), "verbose": false,
                          "architecture": {"filters": 16 * 4, "layers": null}}}"""
    p = tmp_path / "cfg.json"
    p.write_text(src)
    cfg = interpreter.evalFile(str(p))
    assert [h["head-name"] for h in cfg["heads"]] == ["h0", "h1"]
    assert cfg["heads"][1]["counts-loss-weight"] == 100
    assert cfg["settings"] == {"epochs": 20, "verbose": False,
                               "architecture": {"filters": 64, "layers": None}}
    assert interpreter.evalString("x + 1", {"x": 2}, addFunctions=False) == 3
    with pytest.raises(KeyError):
        interpreter.evalString("sqrt(4)", {}, addFunctions=False)
