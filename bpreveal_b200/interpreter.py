"""The configuration-file expression language of the reference (``internal/interpreter.py:1-425``;
``interpretPisa.py:260``, ``interpretFlat.py:317`` read their configuration with ``evalFile``): a
subset of Python *expressions*, of which plain JSON is a special case, evaluated against an
environment.

Supported: numbers, strings (with ``.format``), ``True / False / None`` and their JSON spellings
``true / false / null``, lists, dictionaries, list and dict comprehensions, subscripts, unary
``-`` and ``not``, ``+ - * / % ** //``, ``and / or``, (chained) comparisons including ``in``,
conditional expressions, calls, the builtins ``exp log sqrt abs any all min max sum len range`` and
the constants ``e pi``, and ``lambda`` with the reference's letrec extension: default values that
are lambdas see every default of the same parameter list (so they may be mutually recursive), and
such defaults cannot be overridden at the call site.  Anything else is a ``SyntaxError``.

Own implementation: one node-type -> method table over the tree that ``ast.parse`` returns.
"""
from __future__ import annotations

import ast
import math
import operator


class Closure:
    """A lambda and the environment it was created in (interpreter.py:67-106)."""

    def __init__(self, evaluator, node, env):
        self.ev, self.body = evaluator, node.body
        self.names = [a.arg for a in node.args.args]
        self.env = dict(env)
        first = len(self.names) - len(node.args.defaults)
        self.defaults = {first + i: evaluator.eval(d, env)
                         for i, d in enumerate(node.args.defaults)}
        # letrec: a default that is itself a lambda sees all defaults of this parameter list
        lambdas = [v for v in self.defaults.values() if isinstance(v, Closure)]
        self.letrec = bool(lambdas)
        for fn in lambdas:
            for idx, val in self.defaults.items():
                fn.env[self.names[idx]] = val

    def __call__(self, argv):
        if self.letrec and len(argv) + len(self.defaults) > len(self.names):
            raise SyntaxError("Cannot override default params when one is a lambda.")
        env = dict(self.env)
        for i, name in enumerate(self.names):
            env[name] = argv[i] if i < len(argv) else self.defaults[i]
        return self.ev.eval(self.body, env)


def _star(f):
    return lambda argv: f(*argv)


BUILTINS = {name: _star(fn) for name, fn in dict(
    exp=math.exp, log=math.log, sqrt=math.sqrt, abs=abs, any=any, all=all, min=min, max=max,
    sum=sum, len=len, range=range).items()}
BUILTINS.update(pi=math.pi, e=math.e, true=True, false=False, null=None)

_BINARY = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul,
           ast.Div: operator.truediv, ast.Mod: operator.mod, ast.Pow: operator.pow,
           ast.FloorDiv: operator.floordiv}
_COMPARE = {ast.Lt: operator.lt, ast.Gt: operator.gt, ast.LtE: operator.le, ast.GtE: operator.ge,
            ast.Eq: operator.eq, ast.NotEq: operator.ne,
            ast.In: lambda a, b: a in b, ast.NotIn: lambda a, b: a not in b}


class Evaluator:
    def eval(self, node, env):
        method = getattr(self, "on_" + type(node).__name__, None)
        if method is None:
            raise SyntaxError(f"Unable to interpret {ast.dump(node)} ({ast.unparse(node)})")
        return method(node, env)

    # ---- structure
    def on_Module(self, n, env):
        if len(n.body) != 1:
            raise SyntaxError("ast contains more than one expression.")
        return self.eval(n.body[0], env)

    def on_Expr(self, n, env):
        return self.eval(n.value, env)

    def on_Constant(self, n, env):
        return n.value

    def on_Name(self, n, env):
        return env[n.id]

    # ---- containers
    def on_List(self, n, env):
        return [self.eval(e, env) for e in n.elts]

    def on_Dict(self, n, env):
        return {self.eval(k, env): self.eval(v, env) for k, v in zip(n.keys, n.values)}

    def on_Subscript(self, n, env):
        return self.eval(n.value, env)[self.eval(n.slice, env)]

    def _iterate(self, generators, env, emit):
        """Nested ``for ... if ...`` clauses of a comprehension; ``emit(env)`` per combination."""
        if not generators:
            emit(env)
            return
        gen = generators[0]
        assert isinstance(gen.target, ast.Name), "comprehension targets must be plain names"
        env = dict(env)
        for val in self.eval(gen.iter, env):
            env[gen.target.id] = val
            if all(self.eval(c, env) for c in gen.ifs):
                self._iterate(generators[1:], env, emit)

    def on_ListComp(self, n, env):
        out = []
        self._iterate(n.generators, env, lambda e: out.append(self.eval(n.elt, e)))
        return out

    def on_DictComp(self, n, env):
        out = {}

        def emit(e):
            out[self.eval(n.key, e)] = self.eval(n.value, e)
        self._iterate(n.generators, env, emit)
        return out

    # ---- operators
    def on_BinOp(self, n, env):
        fn = _BINARY.get(type(n.op))
        if fn is None:
            raise SyntaxError(f"Unsupported binary operator in expression: {n.op}")
        lhs, rhs = self.eval(n.left, env), self.eval(n.right, env)
        for v in (lhs, rhs):  # the reference restricts arithmetic to scalars and strings
            assert isinstance(v, (int, float, bool, str)), f"cannot do arithmetic on {v!r}"
        return fn(lhs, rhs)

    def on_UnaryOp(self, n, env):
        if isinstance(n.op, ast.USub):
            v = self.eval(n.operand, env)
            assert isinstance(v, (int, float, bool))
            return -v
        if isinstance(n.op, ast.Not):
            return not self.eval(n.operand, env)
        raise SyntaxError(f"Unsupported unary operator in expression: {n.op}")

    def on_BoolOp(self, n, env):
        # like the reference, the result is a bool (not the deciding operand as in Python)
        if isinstance(n.op, ast.And):
            return all(self.eval(v, env) for v in n.values)
        if isinstance(n.op, ast.Or):
            return any(self.eval(v, env) for v in n.values)
        raise SyntaxError(f"Unsupported boolean operator in expression: {n.op}")

    def on_Compare(self, n, env):
        prev = self.eval(n.left, env)
        for op, rhs in zip(n.ops, n.comparators):
            fn = _COMPARE.get(type(op))
            if fn is None:
                raise SyntaxError(f"Unsupported comparison operator in expression: {op}")
            cur = self.eval(rhs, env)
            if not fn(prev, cur):
                return False
            prev = cur
        return True

    def on_IfExp(self, n, env):
        return self.eval(n.body if self.eval(n.test, env) else n.orelse, env)

    # ---- functions
    def on_Lambda(self, n, env):
        return Closure(self, n, env)

    def on_Call(self, n, env):
        fn = self.eval(n.func, env)
        argv = [self.eval(a, env) for a in n.args]
        if not callable(fn):
            raise ValueError(f"Function {fn} is not a closure or callable.")
        return fn(argv)  # builtins and closures both take the argument LIST

    def on_Attribute(self, n, env):
        val = self.eval(n.value, env)
        if n.attr != "format":
            raise SyntaxError(f"Attribute {n.attr} is not supported.")
        assert isinstance(val, str)
        return lambda argv: val.format(*argv)


_EVALUATOR = Evaluator()


def evalAst(tree, env=None, addFunctions=True):
    """interpreter.py:392-408: evaluate a parsed expression; names resolve in ``env``, with the
    builtins underneath unless ``addFunctions`` is false."""
    env = dict(env or {})
    if addFunctions:
        env = {**BUILTINS, **env}
    return _EVALUATOR.eval(tree, env)


def evalString(expr, env=None, addFunctions=True):
    """interpreter.py:411-416."""
    return evalAst(ast.parse(expr), env, addFunctions)


def evalFile(fname, env=None, addFunctions=True):
    """interpreter.py:419-425: how the reference's programs read their configuration files."""
    with open(fname, "r", encoding="utf-8") as fp:
        return evalString(fp.read(), env, addFunctions)
