#!/usr/bin/env python3
"""Generate the shallow-water Jacobian rows of the reference's JAC callback as straight-line C++ (host + device):
    oracle/gen/swe_jac_gen.h   and   fatode_b200/csrc/gen/swe_jac_gen.h   (identical text)

Source: compute_JacF of EXAMPLES/swe/SWE_ERK_ADJ/swe2D_upwind.F90 (:483-2023), a formula export (1 540 lines, six-digit REAL(4)
literals such as 0.0833333) that fills the dense 4800 x 4800 arrays JF and JG row by row and returns JF = -JF - JG; the JAC
callback of the ERK_ADJ / ERK_TLM drivers hands that matrix to the integrators (swe2D_erk_adj_dr.F90:2-13).  The model is the
user's, not the integrator's (the shipped example models are compiled as device functions), so, like the
KPP mechanism of CBM-IV, it is read at generation time and restated statement by statement with Fortran's evaluation order and
literal kinds; nothing here runs on the GPU box.

Structure of the routine: three loops over the rows p of the h-, u- and v-equations; per row the centre values C1..C3 of the
three fields, their neighbours in the first grid index (NN, N, S, SS: JF part) and in the second (WW, W, E, EE: JG part), and per
eigenvalue sign (feigvalues / geigvalues >= 0) a list of `J?(p,ind) = J?(p,ind) +- expr` with ind = mapgr(p, dk, di, dj), the
periodic neighbour (di, dj) of field (row field + dk).  A row therefore has at most 3 x 5 JF and 3 x 5 JG entries.

Generated interface (namespace swe_jac): rowK(V, fe, ge, dx, dy, g, JF, JG) for K = 0, 1, 2 with
    V[f][0..4] = field f at first-index offsets -2..2, V[f][5..8] = second-index offsets -2, -1, +1, +2
    fe[m], ge[m] = eigenvalue m of the x / y flux Jacobian >= 0 at the row's grid point
    JF[f][di + 2], JG[f][dj + 2] = the row's entries towards field f (raw; the Jacobian is -JF - JG)
Usage: python tools/gen_swe_jac.py [reference_root]
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90expr as fx  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "EXAMPLES/swe/SWE_ERK_ADJ/swe2D_upwind.F90")
OUTS = [os.path.join(HERE, "..", "oracle", "gen", "swe_jac_gen.h"), os.path.join(HERE, "..", "fatode_b200", "csrc", "gen", "swe_jac_gen.h")]

NAMES = {"NN": 0, "N": 1, "C": 2, "S": 3, "SS": 4, "WW": 5, "W": 6, "E": 7, "EE": 8}
OFFS = {"NN": (-2, 0), "N": (-1, 0), "C": (0, 0), "S": (1, 0), "SS": (2, 0), "WW": (0, -2), "W": (0, -1), "E": (0, 1), "EE": (0, 2)}


class Call:
    def __init__(self, name, arg):
        self.name, self.arg = name, arg


class Parser(fx._Parser):
    """f90expr's grammar plus `**` (binds tighter than unary minus and `*`, right associative) and intrinsic calls"""

    def factor(self):
        base = self.primary()
        if self.peek()[1] == "**":
            self.next()
            if self.peek()[1] == "-":
                self.next()
                ex = fx.Neg(self.factor())
            else:
                ex = self.factor()
            return fx.Bin("**", base, ex)
        return base

    def primary(self):
        k, t = self.next()
        if k == "num":
            return fx.parse_number(t)
        if k == "name":
            if self.peek()[1] == "(":
                self.next()
                arg = self.expr()
                self.expect(")")
                return Call(t.upper(), arg)
            return fx.Ref(t.upper(), ())
        if t == "(":
            e = self.expr()
            self.expect(")")
            return e
        raise SyntaxError(f"unexpected {t!r}")


def parse(s):
    p = Parser(s)
    e = p.expr()
    assert p.peek()[0] is None, s
    return e


def kind(e):
    """'d' when the value is DOUBLE PRECISION, 'r4' / 'int' for literal-only subexpressions (must not be combined: they
    would be evaluated in single precision by Fortran)"""
    if isinstance(e, fx.Num):
        return {"int": "int", "r4": "r4", "r8": "d"}[e.kind]
    if isinstance(e, fx.Ref):
        return "d"
    if isinstance(e, fx.Neg):
        return kind(e.a)
    if isinstance(e, Call):
        return kind(e.arg)
    ka, kb = kind(e.a), kind(e.b)
    if e.op == "**":
        assert ka == "d", "power of a literal"
        return "d"
    if ka != "d" and kb != "d":
        raise ValueError("literal-only arithmetic (single precision in Fortran) is not handled")
    return "d"


def emit(e, acc):
    if isinstance(e, fx.Num):
        v = fx.c_double(e.value)
        return f"({v})" if e.value < 0 else v
    if isinstance(e, fx.Ref):
        if e.name == "ACC":
            return acc
        return {"G": "g", "DX": "dx", "DY": "dy", "SQRTC1G": "SqrtC1g"}.get(e.name, e.name)
    if isinstance(e, fx.Neg):
        return f"(-{emit(e.a, acc)})"
    if isinstance(e, Call):
        assert e.name == "SQRT"
        return f"sqrt({emit(e.arg, acc)})"
    if e.op == "**":
        b = emit(e.a, acc)
        x = e.b
        if isinstance(x, fx.Num) and x.kind == "int" and x.value in (2.0, 3.0):   # gfortran expands x**2, x**3 into products
            return f"({b} * {b})" if x.value == 2.0 else f"(({b} * {b}) * {b})"
        if isinstance(x, fx.Num) and x.value == 1.5:                                # REAL(4) 1.5 -> pow(x, 1.5d0)
            return f"SWE_JAC_POW15({b})"
        raise ValueError(f"power {x}")
    return f"({emit(e.a, acc)} {e.op} {emit(e.b, acc)})"


def main():
    lines = open(SRC).read().splitlines()
    start = [i for i, l in enumerate(lines) if re.match(r"\s*subroutine compute_JacF", l, re.I)][0]
    body = fx.logical_lines("\n".join(lines[start + 1:]))
    rows = []          # per row loop: list of items
    cur = None
    field = None       # 1-based field of the current row loop
    stack = []         # open if-blocks: ('eig', which, m) / ('ind',) / ('verbose',)
    pending = None     # (cf, di, dj) of the last mapgr
    nstmt = 0
    for l in body:
        s = l.strip()
        low = s.lower()
        if low.startswith("do p"):
            field = len(rows) + 1
            cur = []
            rows.append(cur)
            continue
        if cur is None:
            continue
        if low == "end do":
            assert not stack
            cur = None
            continue
        m = re.match(r"(\w+)\s*=\s*U\(k([+-]\d+)?,i([+-]\d+)?,j([+-]\d+)?\)$", s)
        if m:
            nm, a, b, c = m.group(1), int(m.group(2) or 0), int(m.group(3) or 0), int(m.group(4) or 0)
            base, f = nm[:-1], int(nm[-1])
            assert f == field + a and OFFS[base] == (b, c), s      # the suffix is the field, the prefix the neighbour
            continue
        if low.startswith("sqrtc1g") or low.startswith("call mapgrinv") or re.match(r"j[fg]\(p,:\)=0d0", low) or low.endswith("continue"):
            continue
        m = re.match(r"if \(([fg])eigvalues\((\d),i,j\) \.GE\. 0d0\) then", s)
        if m:
            cur.append(("if", m.group(1), int(m.group(2)) - 1))
            stack.append("eig")
            continue
        if low == "else":
            assert stack[-1] == "eig"
            cur.append(("else",))
            continue
        if low.startswith("if (verbose"):
            continue                                                # one-line IF ... print
        m = re.match(r"call mapgr\(ind,p,(-?\d+),(-?\d+),(-?\d+)\)", s)
        if m:
            pending = (field - 1 + int(m.group(1)), int(m.group(2)), int(m.group(3)))
            assert 0 <= pending[0] <= 2
            continue
        if low.startswith("if ((ind"):
            stack.append("ind")
            continue
        if low.startswith("end if"):
            if stack.pop() == "eig":
                cur.append(("endif",))
            continue
        m = re.match(r"(J[FG])\(p,ind\)\s*=\s*(.+)$", s)
        if m:
            which, rhs = m.group(1), m.group(2)
            rhs = re.sub(r"J[FG]\(p,ind\)", "ACC", rhs)
            e = parse(rhs)
            kind(e)
            cf, di, dj = pending
            assert (dj == 0) if which == "JF" else (di == 0), s
            cur.append(("stmt", which, cf, di if which == "JF" else dj, e))
            nstmt += 1
            continue
        raise SystemExit(f"unparsed line: {s[:100]}")
    assert len(rows) == 3 and nstmt == 236, (len(rows), nstmt)

    out = []
    w = out.append
    w("// GENERATED by tools/gen_swe_jac.py from compute_JacF of EXAMPLES/swe/SWE_ERK_ADJ/swe2D_upwind.F90 (:483-2023) -- do not edit.")
    w("// Rows of the shallow-water Jacobian as the reference's JAC callback builds them: JF / JG entries of one row (the Jacobian")
    w("// is -JF - JG), statement by statement, Fortran evaluation order, REAL(4) literals rounded to single first.")
    w("#pragma once")
    w("#ifndef SWE_JAC_FN")
    w("#define SWE_JAC_FN static inline")
    w("#endif")
    w("#ifndef SWE_JAC_POW15")
    w("#define SWE_JAC_POW15(x) pow((x), 1.5)")
    w("#endif")
    w("namespace swe_jac {")
    w(f"constexpr int NSTMT = {nstmt};")
    for r, items in enumerate(rows):
        w(f"// row of the {'huv'[r]}-equation")
        w(f"SWE_JAC_FN void row{r}(const double (&V)[3][9], const bool (&fe)[3], const bool (&ge)[3], const double dx, const double dy,")
        w("                       const double g, double (&JF)[3][5], double (&JG)[3][5]) {")
        for f in range(3):
            for nm, ix in NAMES.items():
                w(f"  const double {nm}{f + 1} = V[{f}][{ix}];")
        w("  const double SqrtC1g = sqrt(C1 * g);")
        w("  for (int f = 0; f < 3; ++f) for (int d = 0; d < 5; ++d) { JF[f][d] = 0.0; JG[f][d] = 0.0; }")
        ind = "  "
        for it in items:
            if it[0] == "if":
                w(f"{ind}if ({'fe' if it[1] == 'f' else 'ge'}[{it[2]}]) {{")
                ind += "  "
            elif it[0] == "else":
                w(f"{ind[:-2]}}} else {{")
            elif it[0] == "endif":
                ind = ind[:-2]
                w(f"{ind}}}")
            else:
                _, which, cf, d, e = it
                acc = f"{which}[{cf}][{d + 2}]"
                w(f"{ind}{acc} = {emit(e, acc)};")
        w("  (void)NN1; (void)N1; (void)S1; (void)SS1; (void)WW1; (void)W1; (void)E1; (void)EE1;")
        w("  (void)NN2; (void)N2; (void)S2; (void)SS2; (void)WW2; (void)W2; (void)E2; (void)EE2;")
        w("  (void)NN3; (void)N3; (void)S3; (void)SS3; (void)WW3; (void)W3; (void)E3; (void)EE3; (void)C2; (void)C3; (void)SqrtC1g;")
        w("}")
    w("}  // namespace swe_jac")
    text = "\n".join(out) + "\n"
    for o in OUTS:
        os.makedirs(os.path.dirname(o), exist_ok=True)
        open(o, "w").write(text)
    print(f"{nstmt} statements, {len(text.splitlines())} lines -> {', '.join(os.path.relpath(o) for o in OUTS)}")


if __name__ == "__main__":
    main()
