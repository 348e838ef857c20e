"""Tiny Fortran-90 assignment-statement reader used by the model generators.

It understands exactly what KPP-style generated model code needs:
straight-line assignments ``NAME(i[,j]) = <arithmetic expression>`` with
``+ - * /``, parentheses, array references, integer and real literals and
free-form ``&`` continuations.  It keeps Fortran's typing rule for literals
(SURVEY.md fact 3): an un-suffixed / ``e``-exponent real literal is REAL(4)
and is widened to double *after* rounding to single; a ``d``-exponent literal
is DOUBLE PRECISION; an integer literal in a real expression is exact.

This is build tooling (reads /root/reference at generation time only); the
generated headers are committed so nothing here runs on the GPU box.
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass
from typing import List, Tuple, Union

# --------------------------------------------------------------------------
# source handling


def strip_comment(line: str) -> str:
    # no string literals occur in the numerical sections we read
    i = line.find("!")
    return line if i < 0 else line[:i]


def logical_lines(text: str) -> List[str]:
    """Join free-form continuation lines, drop comments and blanks."""
    out: List[str] = []
    cur = ""
    for raw in text.splitlines():
        s = strip_comment(raw).strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1].rstrip()
            continue
        cur += s
        out.append(cur)
        cur = ""
    if cur:
        out.append(cur)
    return out


def routine_body(text: str, name: str) -> List[str]:
    """Logical lines between ``SUBROUTINE name`` and ``END SUBROUTINE``."""
    lines = text.splitlines()
    start = end = None
    pat = re.compile(r"^\s*SUBROUTINE\s+" + re.escape(name) + r"\b", re.I)
    for i, l in enumerate(lines):
        if start is None and pat.match(l):
            start = i + 1
        elif start is not None and re.match(r"^\s*END\s+SUBROUTINE", l, re.I):
            end = i
            break
    if start is None or end is None:
        raise KeyError(name)
    return logical_lines("\n".join(lines[start:end]))


# --------------------------------------------------------------------------
# AST


@dataclass(frozen=True)
class Num:
    value: float          # value AFTER Fortran's kind rounding, as a double
    text: str             # original spelling
    kind: str             # 'int' | 'r4' | 'r8'


@dataclass(frozen=True)
class Ref:
    name: str
    idx: Tuple[int, ...]  # 1-based constant indices (empty = scalar)


@dataclass(frozen=True)
class Neg:
    a: "Expr"


@dataclass(frozen=True)
class Bin:
    op: str
    a: "Expr"
    b: "Expr"


Expr = Union[Num, Ref, Neg, Bin]

_TOK = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?)"
    r"|(?P<name>[A-Za-z_][A-Za-z_0-9]*)"
    r"|(?P<op>\*\*|[-+*/(),]))"
)


def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


def parse_number(text: str) -> Num:
    t = text.lower()
    if re.fullmatch(r"\d+", t):
        return Num(float(int(t)), text, "int")
    if "d" in t:
        return Num(float(t.replace("d", "e")), text, "r8")
    # REAL(4): round decimal -> nearest single (incl. subnormals), then widen
    import numpy as np

    return Num(float(np.float32(t)), text, "r4")


class _Parser:
    def __init__(self, s: str):
        self.toks: List[Tuple[str, str]] = []
        pos = 0
        while pos < len(s):
            m = _TOK.match(s, pos)
            if not m:
                if s[pos:].strip() == "":
                    break
                raise SyntaxError(f"bad token at {s[pos:pos+20]!r} in {s!r}")
            pos = m.end()
            for k in ("num", "name", "op"):
                if m.group(k) is not None:
                    self.toks.append((k, m.group(k)))
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def expect(self, v):
        k, t = self.next()
        if t != v:
            raise SyntaxError(f"expected {v!r} got {t!r}")

    # expr := ['-'|'+'] term {('+'|'-') term}      (Fortran: unary minus binds
    #                                               looser than '*')
    def expr(self) -> Expr:
        k, t = self.peek()
        if t == "-":
            self.next()
            e: Expr = Neg(self.term())
        elif t == "+":
            self.next()
            e = self.term()
        else:
            e = self.term()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            e = Bin(op, e, self.term())
        return e

    def term(self) -> Expr:
        e = self.factor()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            e = Bin(op, e, self.factor())
        return e

    def factor(self) -> Expr:
        k, t = self.next()
        if k == "num":
            return parse_number(t)
        if k == "name":
            if self.peek()[1] == "(":
                self.next()
                idx = []
                while True:
                    kk, tt = self.next()
                    if kk != "num" or not tt.isdigit():
                        raise SyntaxError("non-constant index")
                    idx.append(int(tt))
                    kk, tt = self.next()
                    if tt == ")":
                        break
                    if tt != ",":
                        raise SyntaxError("bad index list")
                return Ref(t.upper(), tuple(idx))
            return Ref(t.upper(), ())
        if t == "(":
            e = self.expr()
            self.expect(")")
            return e
        if t == "-":  # e.g. a*(-b) handled by '(' branch; bare is unusual
            return Neg(self.factor())
        raise SyntaxError(f"unexpected {t!r}")


def parse_expr(s: str) -> Expr:
    p = _Parser(s)
    e = p.expr()
    if p.peek()[0] is not None:
        raise SyntaxError(f"trailing tokens in {s!r}")
    return e


@dataclass(frozen=True)
class Assign:
    lhs: Ref
    rhs: Expr


_ASSIGN = re.compile(r"^([A-Za-z_][A-Za-z_0-9]*\s*\([0-9,\s]+\))\s*=\s*(.+)$")


def parse_assignments(lines: List[str]) -> List[Assign]:
    """Every ``X(const idx) = expr`` line; anything else is ignored."""
    out = []
    for l in lines:
        m = _ASSIGN.match(l)
        if not m:
            continue
        lhs = parse_expr(m.group(1))
        assert isinstance(lhs, Ref)
        out.append(Assign(lhs, parse_expr(m.group(2))))
    return out


# --------------------------------------------------------------------------
# C emission


def c_double(x: float) -> str:
    if x == int(x) and abs(x) < 1e15:
        return f"{int(x)}.0"
    return repr(x)  # shortest repr round-trips the double exactly


def emit_c(e: Expr, ref) -> str:
    """C text with Fortran's evaluation order made explicit by parentheses.

    ``ref(name, idx)`` renders an array/scalar reference.
    """
    if isinstance(e, Num):
        v = c_double(e.value)
        return f"({v})" if e.value < 0 else v
    if isinstance(e, Ref):
        return ref(e.name, e.idx)
    if isinstance(e, Neg):
        return f"(-{emit_c(e.a, ref)})"
    if isinstance(e, Bin):
        return f"({emit_c(e.a, ref)} {e.op} {emit_c(e.b, ref)})"
    raise TypeError(e)


def count_flops(e: Expr) -> int:
    if isinstance(e, Bin):
        return 1 + count_flops(e.a) + count_flops(e.b)
    if isinstance(e, Neg):
        return count_flops(e.a)
    return 0
