"""A tiny list-manipulation DSL: interpreter, random program sampler and (input, output → program) dataset synthesis.

Reference counterpart: examples/experiments/grounded_program_synthesis/lang.py (same primitives: take / drop / minimum /
maximum / reverse / sort_asc / sort_des / add_n / sub_n / mul_n / div_n / expand_copy).  Written independently: programs are
parsed by a small recursive-descent parser into an AST and evaluated against a primitive table — no `eval` of model output.
"""
from __future__ import annotations

import random
import re
from typing import Callable, Dict, List, Tuple, Union

Value = Union[int, List[int]]

PRIMITIVES: Dict[str, Tuple[Callable[..., Value], Tuple[str, ...]]] = {
    "take": (lambda xs, n: xs[:n], ("list", "int")),
    "drop": (lambda xs, n: xs[n:], ("list", "int")),
    "minimum": (lambda xs: min(xs), ("list",)),
    "maximum": (lambda xs: max(xs), ("list",)),
    "reverse": (lambda xs: xs[::-1], ("list",)),
    "sort_asc": (lambda xs: sorted(xs), ("list",)),
    "sort_des": (lambda xs: sorted(xs, reverse=True), ("list",)),
    "add_n": (lambda xs, n: [x + n for x in xs], ("list", "int")),
    "sub_n": (lambda xs, n: [x - n for x in xs], ("list", "int")),
    "mul_n": (lambda xs, n: [x * n for x in xs], ("list", "int")),
    "div_n": (lambda xs, n: [x / n for x in xs], ("list", "int")),  # true division, as in the reference (results are floats)
    "expand_copy": (lambda xs: xs + xs, ("list",)),
}
_TOKEN = re.compile(r"\s*(?:(-?\d+)|([A-Za-z_]\w*)|(.))")


class ParseError(ValueError):
    pass


def _tokens(src: str):
    for num, name, sym in _TOKEN.findall(src):
        if num:
            yield ("int", int(num))
        elif name:
            yield ("name", name)
        elif sym.strip():
            yield ("sym", sym)


class Interpreter:
    """`Interpreter()("div_n(reverse([-2, -5, -4]),1)") → [-4, -5, -2]`; returns the string "ERROR" for anything that
    does not parse, type-check or evaluate."""

    def __call__(self, program: str) -> Union[Value, str]:
        try:
            toks = list(_tokens(program))
            value, rest = self._expr(toks)
            if rest:
                raise ParseError("trailing input")
            return value
        except Exception:
            return "ERROR"

    def _expr(self, toks):
        if not toks:
            raise ParseError("unexpected end")
        kind, val = toks[0]
        if kind == "int":
            return val, toks[1:]
        if kind == "sym" and val == "[":
            items, toks = [], toks[1:]
            while toks and toks[0] != ("sym", "]"):
                item, toks = self._expr(toks)
                if not isinstance(item, int):
                    raise ParseError("lists hold integers")
                items.append(item)
                if toks and toks[0] == ("sym", ","):
                    toks = toks[1:]
            if not toks:
                raise ParseError("unclosed list")
            return items, toks[1:]
        if kind == "name":
            if val not in PRIMITIVES or len(toks) < 2 or toks[1] != ("sym", "("):
                raise ParseError(f"unknown call {val}")
            fn, sig = PRIMITIVES[val]
            args, toks = [], toks[2:]
            while toks and toks[0] != ("sym", ")"):
                arg, toks = self._expr(toks)
                args.append(arg)
                if toks and toks[0] == ("sym", ","):
                    toks = toks[1:]
            if not toks or len(args) != len(sig):
                raise ParseError("bad arity")
            for a, t in zip(args, sig):
                if (t == "list") != isinstance(a, list):
                    raise ParseError("type error")
            return fn(*args), toks[1:]
        raise ParseError(f"unexpected {val!r}")


def random_list(rng: random.Random, max_len: int = 5, span: int = 5) -> List[int]:
    return [rng.randint(-span, span) for _ in range(rng.randint(1, max_len))]


def sample_program(rng: random.Random, depth: int = 2) -> str:
    """A random well-typed program of nesting depth ≤ `depth` that evaluates to a list."""
    if depth == 0:
        return str(random_list(rng))
    name = rng.choice([n for n, (_, sig) in PRIMITIVES.items() if n not in ("minimum", "maximum")])
    _, sig = PRIMITIVES[name]
    inner = sample_program(rng, depth - 1)
    if len(sig) == 2:
        n = rng.randint(1, 4)
        return f"{name}({inner},{n})"
    return f"{name}({inner})"


def create_synthetic_dataset(size: int, seed: int = 0, depth: int = 2) -> List[Dict[str, str]]:
    """Records `{"input": "Input: <k> Output: <value> Function:", "output": "<program>"}` — the model sees the I/O example
    and must write a program that reproduces the output."""
    rng, interp, data = random.Random(seed), Interpreter(), []
    while len(data) < size:
        prog = sample_program(rng, rng.randint(1, depth))
        out = interp(prog)
        if out == "ERROR":
            continue
        data.append({"input": f"Input: {rng.randint(1, 4)} Output: {out} Function:", "output": " " + prog})
    return data
