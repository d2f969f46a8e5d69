"""Golden vectors for bpreveal_b200/interpreter.py, produced by the REFERENCE's own evaluator
(/root/reference/src/internal/interpreter.py, pure Python: ast + math) in the build container:

    python tests/golden/make_interpreter_golden.py   ->  tests/golden/interpreter_vectors.json

Each vector: expression, environment, and either the value the reference returns (JSON-encoded;
ranges as lists) or the name of the exception it raises."""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    "ref_interpreter", "/root/reference/src/internal/interpreter.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [
    # the expressions of the reference's self-test (interpreter.py:428-462) ...
    ("2 + 2", {}), ("x - 1", {"x": 3}), ("1 if 3 > 2 else 0", {}), ("(lambda x: x + 1)(3)", {}),
    ("(lambda x, y: y(x - 1))(3, lambda a: a*3)", {}),
    ("(lambda x, y: y(x, y)) (5, lambda a, b: 1 if a <= 1 else a * b(a-1, b))", {}),
    ("((lambda x: lambda x, y: y(x, y))(0)) (5, lambda a, b: 1 if a <= 1 else a * b(a-1, b))", {}),
    ("sqrt(4)", {}), ("len(s * 3)", {"s": "aoeu"}), ("(lambda x, y=3: x + y)(2)", {}),
    ("(lambda x, y=3: x + y)(2, 9)", {}), ("(lambda x=5, y=(lambda q: x): x + y(0))()", {}),
    ("(lambda x=5, y=(lambda n: 1 if n <= 1 else n * y(n - 1)): y(x))()", {}),
    ("(lambda x, y=(lambda n: 1 if n <= 1 else n * y(n - 1)): y(x))(5)", {}),
    ("(lambda n, isOdd=(lambda n: True if n == 1 else not isEven(n - 1)), "
     "isEven=(lambda n: True if n == 0 else not isOdd(n - 1)): isOdd(n))(5)", {}),
    ("(lambda n, isOdd=(lambda n: False if n == 0 else not isEven(n - 1)), "
     "isEven=(lambda n: True if n == 0 else not isOdd(n - 1)): isOdd(n))(100)", {}),
    ("(lambda a=5, b=(lambda: a): b)()()", {}), ("(lambda a=5, b=(lambda: a): b)(6)()", {}),
    ("[1, 2, 3]", {}), ("range(5)", {}), ("[x for x in range(5)]", {}),
    ("[x + y for x in range(2) for y in range(3)]", {}),
    ("[y for x in [[1, 2], [3, 4]] for y in x]", {}),
    ("{y: x for x in [[1, 2], [3, 4]] for y in x if y > 2}", {}), ("[1, 2, 3][1]", {}),
    ("(lambda x: x)\n(1)", {}), ('"{0:s}h{0:s}".format("aoeu")', {}),
    ("(lambda x: [x[0.2], x['aoeu']])({0.2: 5, 'aoeu': 'htns'})", {}), ("{}", {}),
    # ... and more of the documented surface (interpreter.py:12-33)
    ("0.2 < x < 0.5", {"x": 0.3}), ("0.2 < x < 0.5", {"x": 0.7}), ("3 in [1, 2, 3] and not false", {}),
    ("null", {}), ("[true, false, null, True, False, None]", {}), ("7 // 2 + 7 % 2 + 2 ** 10", {}),
    ("-x + abs(-3) + max(1, 5) + min([4, 2]) + sum([1, 2, 3])", {"x": 2}),
    ("any([False, x > 1]) or all([])", {"x": 2}), ("log(e) + exp(0) + pi - pi", {}),
    ("1 and 2", {}), ("0 or 3", {}), ("'a' + 'b' * 2", {}), ("[1] + [2]", {}),
    ("x.upper()", {"x": "a"}), ("[i for i, j in [[1, 2]]]", {}), ("1; 2", {}), ("~1", {}),
    ("1 is 1", {}), ("f'{1}'", {}), ("{'a': {'b': [1, 2.5, 'c']}, 'n': -3}", {}),
    ('{"heads": [{"num-tasks": n, "head-name": "h{0:d}".format(i)} for i in range(2)], '
     '"settings": {"epochs": 10 * k, "max-jitter": 100 if k > 1 else 0}}', {"n": 2, "k": 3}),
    ("undefined_name", {}), ("(lambda x: x)(1, 2)", {}), ("sqrt", {}),
]


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


out = []
for expr, env in CASES:
    try:
        out.append({"expr": expr, "env": env, "value": enc(ref.evalString(expr, dict(env)))})
    except BaseException as ex:  # noqa: B036  (assertions, KeyError, SyntaxError ... all count)
        out.append({"expr": expr, "env": env, "error": type(ex).__name__})
with open(os.path.join(HERE, "interpreter_vectors.json"), "w", encoding="utf-8") as fp:
    json.dump(out, fp, indent=1)
print(len(out), "vectors;", sum("error" in v for v in out), "of them errors")
