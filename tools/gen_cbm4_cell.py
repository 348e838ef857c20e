#!/usr/bin/env python3
"""Generate fatode_b200/csrc/gen/cbm4_cell_gen.h — CBM-IV as straight-line per-thread code
("one cell per thread" forward sweep; PRODUCT side, no oracle code involved).

The reference's KPP module (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90) is restated statement by
statement with Fortran's evaluation order and REAL(4) literal rounding, so FUN and
JAC are bit-identical to the reference arithmetic when compiled without FMA contraction.

Linear algebra: the reference factorises E = 1/(h*gamma) I - J with netlib DGETF2 (partial pivoting).
Along CBM-IV trajectories the pivot sequence is almost static: over 6.7e5 factorisations of perturbed
cells only four interchanges ever occur (column 12 with row 19, 18 with 27, 21 with 26, 26 with 27;
0-based).  The generated `lu_factor` is DGETF2 restricted to exactly those candidate interchanges:
  * the symbolic pattern is the union over all 2^4 interchange outcomes (in "position" space, rows are
    physically exchanged like DSWAP does), 367 of 1024 entries, 884 multiply-adds;
  * at every column the IDAMAX decision is re-derived from the data; an unforeseen pivot, a pivot below
    DLAMCH('S') or a zero pivot returns non-zero and the caller hands the cell to the general
    (warp-per-cell, full pivoting) kernel;
  * operations on structural zeros of the actual variant add exact zeros, so the factors and the
    DGETRS('N') substitutions are bit-identical to netlib's (DGER/DTRSM skip exact zeros).
`lu_solve` is DGETRS('N') on that pattern (true divisions by the diagonal, like DTRSM).

Adjoint-side helpers (values the backward sweep needs per stage; not bit-critical): jac_values,
jac_time, hess_values, build_M, param_terms.

The header compiles for the device (nvcc -fmad=false) and for the host (g++ -ffp-contract=off; used by
tests/test_cell_codegen.py to check the generated LU/solve against the oracle's DGETF2/DGETRS bit for bit).
Usage: python tools/gen_cbm4_cell.py [reference_root]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90expr as fx  # noqa: E402
import gen_cbm4_device as gd  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90")
N = 32
SWAPS = [(12, 19), (18, 27), (21, 26), (26, 27)]   # allowed DGETF2 interchanges (column, partner row), 0-based


def symbolic(jpat):
    """Union-pattern symbolic DGETF2 with the candidate interchanges.  Returns the final pattern and the
    per-column work lists (cand, swap columns, rows, cols) in elimination order."""
    P = [[False] * N for _ in range(N)]
    for (r, c) in jpat:
        P[r][c] = True
    for i in range(N):
        P[i][i] = True
    steps = []
    for k in range(N):
        cand = [i for i in range(k, N) if P[i][k]]
        sw = [p for (kk, p) in SWAPS if kk == k]
        assert len(sw) <= 1
        swcols = None
        if sw:
            p = sw[0]
            assert p in cand, (k, p, cand)
            swcols = [c for c in range(N) if P[k][c] or P[p][c]]
            for c in swcols:
                P[k][c] = P[p][c] = True
        rows = [i for i in range(k + 1, N) if P[i][k]]
        cols = [c for c in range(k + 1, N) if P[k][c]]
        for i in rows:
            for c in cols:
                P[i][c] = True
        steps.append(dict(k=k, cand=cand, p=sw[0] if sw else None, swcols=swcols, rows=rows, cols=cols))
    return P, steps


def emit_paired(w, name, exprs):
    """out[e] = expr_e, stored four at a time with one 32-byte store (a full DRAM sector: the outputs go straight to the
    tape in HBM, partial-sector writes would cost a read-modify-write there).  The list is zero-padded to a multiple
    of four; CBM4_CELL_ST4(ptr, a, b, c, d) is provided by the includer."""
    ex = list(exprs)
    while len(ex) % 4:
        ex.append("0.0")
    for e in range(0, len(ex), 4):
        w(f"  {{ const double a = {ex[e]}; const double b = {ex[e + 1]}; const double c = {ex[e + 2]}; const double d = {ex[e + 3]};")
        w(f"    CBM4_CELL_ST4({name} + {e}, a, b, c, d); }}")


def main():
    text = open(SRC).read()
    mech = gd.parse_mechanism(text)
    body = lambda nm: fx.parse_assignments(fx.routine_body(text, nm))
    photo_idx = sorted(mech.photo)          # reaction indices (0-based) with RCONST = c*SUN
    ph_of = {r: k for k, r in enumerate(photo_idx)}

    def ref(name, idx):
        if name == "RCT":
            r = idx[0] - 1
            return f"PH[{ph_of[r]}]" if r in ph_of else f"CBM4_CELL_RC({r})"
        if name == "F":
            assert idx == (1,)
            return "F"
        if name in ("V", "VDOT", "HESS"):              # thread-private vectors: stride SV (shared-memory slots)
            return f"{name}[{idx[0]-1} * SV]"
        if name in ("A", "B", "D2A"):                   # routine-local temporaries (registers)
            return f"{name}[{idx[0]-1}]"
        raise ValueError((name, idx))

    jent = mech.jent                       # (row, col) of the 276 Jacobian entries, by column then row
    jpos = {rc: e for e, rc in enumerate(jent)}
    P, steps = symbolic(jent)
    idx = {}
    for i in range(N):
        for c in range(N):
            if P[i][c]:
                idx[(i, c)] = len(idx)
    NLU = len(idx)
    rowptr = []
    for i in range(N):
        rowptr.append(min(v for (r, c), v in idx.items() if r == i))
    rowptr.append(NLU)
    diag = [idx[(i, i)] for i in range(N)]
    nupd = sum(len(s["rows"]) * len(s["cols"]) for s in steps)

    o = []
    w = o.append
    w("// GENERATED by tools/gen_cbm4_cell.py from the reference's CBM-IV mechanism")
    w("// (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90) — do not edit.  PRODUCT side:")
    w("// straight-line per-thread code (one cell per thread) + DGETF2/DGETRS on the static union pattern.")
    w("// The includer defines CBM4_CELL_FN / CBM4_CELL_INL (function qualifiers) and CBM4_CELL_RC(r) (time-independent rate")
    w("// coefficient r), CBM4_CELL_HESS(h) (the constant Hessian values) and CBM4_CELL_ST4(p, a, b, c, d) (store four")
    w("// consecutive doubles at the 32-byte aligned p); compile without FMA contraction.")
    w("#pragma once")
    w("#include <math.h>")
    w("#include <float.h>")
    w("namespace cbm4_cell {")
    w(f"constexpr int NVAR = {N}, NLU = {NLU}, NLU_UPDATES = {nupd}, NJ = {len(jent)}, NM = {mech.NM}, NHESS = 258, "
      f"NJT = {len(mech.jtime)}, NPHOTO = {len(photo_idx)}, NPAR = {gd.NP};")
    w(f"// allowed interchanges: bit b of `swaps` = column {[k for k, _ in SWAPS]} exchanged with row {[p for _, p in SWAPS]}")
    w(f"constexpr int NSWAP = {len(SWAPS)};")
    w(f"constexpr int SWAP_COL[{len(SWAPS)}] = {{{', '.join(str(k) for k, _ in SWAPS)}}};")
    w(f"constexpr int SWAP_ROW[{len(SWAPS)}] = {{{', '.join(str(p) for _, p in SWAPS)}}};")
    w(f"// row-major storage of the union pattern: entries of row i are LU_ROWPTR[i]..LU_ROWPTR[i+1]-1, columns LU_COL[]")
    w(f"constexpr int LU_ROWPTR[{N + 1}] = {{{', '.join(map(str, rowptr))}}};")
    cols_flat = [c for (i, c), v in sorted(idx.items(), key=lambda kv: kv[1])]
    w(f"constexpr unsigned char LU_COL[{NLU}] = {{{', '.join(map(str, cols_flat))}}};")
    w(f"constexpr int LU_DIAG[{N}] = {{{', '.join(map(str, diag))}}};")
    w(f"constexpr int PHOTO_IDX[{len(photo_idx)}] = {{{', '.join(map(str, photo_idx))}}};")
    w(f"constexpr double PHOTO_C[{len(photo_idx)}] = {{{', '.join(fx.c_double(mech.photo[r]) for r in photo_idx)}}};")
    w("")

    # ------------------------------------------------------------------ FUN
    w("// CBM4_FUN (:320-449): VDOT = f(V) with RCT = (constants, PH = photolysis rates at the current time)")
    w("template <int SV> CBM4_CELL_FN void fun(const double* __restrict__ V, const double F, const double* __restrict__ PH, "
      "double* __restrict__ VDOT) {")
    w("  double A[81];")
    for a in body("CBM4_FUN"):
        w(f"  {ref(a.lhs.name, a.lhs.idx)} = {fx.emit_c(a.rhs, ref)};")
    w("}")
    w("")

    # ------------------------------------------------------------------ E = ghinv*I - J on the union pattern
    jac = body("CBM4_JAC")
    w("// CBM4_JAC (:451-1016) fused with lss_decomp's  e = -fjac; e(j,j) = e(j,j) + 1/(h*gamma)  (LS_Solver.F90:36-40):")
    w("// writes E on the union pattern (fill positions = 0)")
    w("template <int SV> CBM4_CELL_FN void build_E(const double* __restrict__ V, const double F, const double* __restrict__ PH, "
      "const double ghinv, double* __restrict__ lu) {")
    w("  double B[138];")
    written = set()
    for a in jac:
        if a.lhs.name == "B":
            w(f"  B[{a.lhs.idx[0]-1}] = {fx.emit_c(a.rhs, ref)};")
    for a in jac:
        if a.lhs.name == "JV":
            r, c = a.lhs.idx[0] - 1, a.lhs.idx[1] - 1
            e = f"(-{fx.emit_c(a.rhs, ref)})"
            if r == c:
                e = f"({e} + ghinv)"
            w(f"  lu[{idx[(r, c)]} * SV] = {e};")
            written.add((r, c))
    assert len(written) == len(jent)
    for (r, c), v in sorted(idx.items(), key=lambda kv: kv[1]):
        if (r, c) not in written:
            w(f"  lu[{v} * SV] = 0.0;")
    w("}")
    w("")

    # ------------------------------------------------------------------ dense E for the full-pivoting fallback
    w("// Same as build_E but on a dense column-major 32 x 32 matrix, element (r,c) at E[(r + 32*c) * SV] (structural zeros")
    w("// written as 0.0): input of the general DGETF2 / ZGETF2 of the implicit families' fallback policy.")
    w("template <int SV> CBM4_CELL_FN void build_E_dense(const double* __restrict__ V, const double F, const double* __restrict__ PH, "
      "const double ghinv, double* __restrict__ E) {")
    w("  double B[138];")
    for a in jac:
        if a.lhs.name == "B":
            w(f"  B[{a.lhs.idx[0]-1}] = {fx.emit_c(a.rhs, ref)};")
    w("  for (int q = 0; q < 1024; ++q) E[q * SV] = 0.0;")
    dense_written = set()
    for a in jac:
        if a.lhs.name == "JV":
            r, c = a.lhs.idx[0] - 1, a.lhs.idx[1] - 1
            e = f"(-{fx.emit_c(a.rhs, ref)})"
            if r == c:
                e = f"({e} + ghinv)"
            w(f"  E[{r + 32 * c} * SV] = {e};")
            dense_written.add((r, c))
    for i in range(N):
        assert (i, i) in dense_written
    w("}")
    w("")
    # ------------------------------------------------------------------ DGETF2 on the union pattern
    w("// |x| as an ordered integer (IEEE doubles of equal sign order like their bit patterns)")
    w("CBM4_CELL_INL unsigned long long absbits(double x) {")
    w("#ifdef __CUDA_ARCH__")
    w("  return (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull;")
    w("#else")
    w("  unsigned long long u; __builtin_memcpy(&u, &x, 8); return u & 0x7fffffffffffffffull;")
    w("#endif")
    w("}")
    w("// DGETF2 restricted to the candidate interchanges.  lu: E in, L\\U out (row-major union pattern, rows in")
    w("// pivoted order like LAPACK's in-place result).  dinv[k] = 1/U(k,k).  swaps: bit b set = interchange b taken.")
    w("// Returns 0 when netlib's DGETF2 would have made exactly these choices with pivots >= DLAMCH('S'); non-zero")
    w("// otherwise (unforeseen pivot row / tiny or zero pivot): the factors are then meaningless.")
    w("template <int SV> CBM4_CELL_FN int lu_factor(double* __restrict__ lu, double* __restrict__ dinv, unsigned& swaps) {")
    w("  int bad = 0;")
    w("  unsigned sw = 0u;")
    for s in steps:
        k = s["k"]
        cand = s["cand"]
        w(f"  {{  // column {k}")
        if s["p"] is not None:
            p = s["p"]
            b = [i for i, (kk, _) in enumerate(SWAPS) if kk == k][0]
            w(f"    unsigned long long dmax = absbits(lu[{idx[(k, k)]} * SV]); int jp = {k};")
            for i in cand[1:]:
                w(f"    {{ const unsigned long long v = absbits(lu[{idx[(i, k)]} * SV]); if (v > dmax) {{ dmax = v; jp = {i}; }} }}")
            w(f"    const bool take = (jp == {p});")
            w(f"    bad |= (jp != {k}) & (jp != {p});")
            w(f"    if (take) {{")
            w(f"      sw |= {1 << b}u;")
            for c in s["swcols"]:
                w(f"      {{ const double t = lu[{idx[(k, c)]} * SV]; lu[{idx[(k, c)]} * SV] = lu[{idx[(p, c)]} * SV]; lu[{idx[(p, c)]} * SV] = t; }}")
            w("    }")
        elif len(cand) > 1:
            w(f"    const unsigned long long d = absbits(lu[{idx[(k, k)]} * SV]);")
            terms = " | ".join(f"(absbits(lu[{idx[(i, k)]} * SV]) > d)" for i in cand[1:])
            w(f"    bad |= {terms};")
        w(f"    const double piv = lu[{idx[(k, k)]} * SV];")
        w(f"    bad |= !(fabs(piv) >= DBL_MIN);")
        w(f"    const double r = 1.0 / piv;")
        w(f"    dinv[{k} * SV] = r;")
        if k < N - 1:
            for i in s["rows"]:
                w(f"    const double l{i} = r * lu[{idx[(i, k)]} * SV]; lu[{idx[(i, k)]} * SV] = l{i};")
            for c in s["cols"]:
                w(f"    {{ const double t = -lu[{idx[(k, c)]} * SV];")
                for i in s["rows"]:
                    w(f"      lu[{idx[(i, c)]} * SV] = lu[{idx[(i, c)]} * SV] + l{i} * t;")
                w("    }")
        w("  }")
    w("  swaps = sw;")
    w("  return bad;")
    w("}")
    w("")

    # ------------------------------------------------------------------ DGETRS('N')
    w("// DGETRS('N'), one right-hand side, on the factors of lu_factor: DLASWP, DTRSM(L,L,N,U), DTRSM(L,U,N,N)")
    w("template <int SV> CBM4_CELL_FN void lu_solve(const double* __restrict__ lu, const unsigned swaps, double* __restrict__ b) {")
    for bi, (k, p) in enumerate(SWAPS):
        w(f"  if (swaps & {1 << bi}u) {{ const double t = b[{k} * SV]; b[{k} * SV] = b[{p} * SV]; b[{p} * SV] = t; }}")
    for k in range(N - 1):
        rows = [i for i in range(k + 1, N) if P[i][k]]
        if not rows:
            continue
        w(f"  {{ const double bk = b[{k} * SV];")
        for i in rows:
            w(f"    b[{i} * SV] = b[{i} * SV] - bk * lu[{idx[(i, k)]} * SV];")
        w("  }")
    for k in range(N - 1, -1, -1):
        rows = [i for i in range(k) if P[i][k]]
        w(f"  {{ const double bk = b[{k} * SV] / lu[{idx[(k, k)]} * SV]; b[{k} * SV] = bk;")
        for i in rows:
            w(f"    b[{i} * SV] = b[{i} * SV] - bk * lu[{idx[(i, k)]} * SV];")
        w("  }")
    w("}")
    w("")

    # ------------------------------------------------------------------ ZGETF2 / ZGETRS('N') on the same pattern
    w("// ---- COMPLEX*16 factorisation of the implicit Runge-Kutta families (FWD/RK/LS_Solver.F90:22-52): the matrix")
    w("// -J + (alpha + i beta) I lives in two planes zr/zi on the union pattern.  netlib ZGETF2 restricted to the candidate")
    w("// interchanges: IZAMAX on |re|+|im| (first maximum), ZSCAL by the reciprocal ONE/pivot, ZGERU update")
    w("// a(i,c) += l(i)*(-y(c)); products in the plain four-multiplication form, quotients in the range-reduced (Smith)")
    w("// form gfortran emits under -fcx-fortran-rules.  Agrees with the reference bit for bit except for the sign of zeros.")
    w("// imaginary plane of  -J + (alpha + i beta) I  on the union pattern: beta on the diagonal, zero elsewhere")
    w("template <int SV> CBM4_CELL_FN void zimag_init(double* __restrict__ zi, const double beta) {")
    dset = set(diag)
    for e in range(NLU):
        w(f"  zi[{e} * SV] = {'beta' if e in dset else '0.0'};")
    w("}")
    w("CBM4_CELL_INL void zdiv(const double a, const double b, const double c, const double e, double& qr, double& qi) {")
    w("  if (fabs(c) < fabs(e)) {")
    w("    const double ratio = c / e, div = c * ratio + e;")
    w("    qr = (a * ratio + b) / div; qi = (b * ratio - a) / div;")
    w("  } else {")
    w("    const double ratio = e / c, div = e * ratio + c;")
    w("    qr = (b * ratio + a) / div; qi = (b - a * ratio) / div;")
    w("  }")
    w("}")
    w("template <int SV> CBM4_CELL_FN int zlu_factor(double* __restrict__ zr, double* __restrict__ zi, unsigned& swaps) {")
    w("  int bad = 0;")
    w("  unsigned sw = 0u;")
    cabs1 = lambda e: f"(fabs(zr[{e} * SV]) + fabs(zi[{e} * SV]))"
    for s_ in steps:
        k = s_["k"]
        cand = s_["cand"]
        w(f"  {{  // column {k}")
        if s_["p"] is not None:
            p_ = s_["p"]
            b = [i for i, (kk, _) in enumerate(SWAPS) if kk == k][0]
            w(f"    double dmax = {cabs1(idx[(k, k)])}; int jp = {k};")
            for i in cand[1:]:
                w(f"    {{ const double v = {cabs1(idx[(i, k)])}; if (v > dmax) {{ dmax = v; jp = {i}; }} }}")
            w(f"    const bool take = (jp == {p_});")
            w(f"    bad |= (jp != {k}) & (jp != {p_});")
            w(f"    if (take) {{")
            w(f"      sw |= {1 << b}u;")
            for c in s_["swcols"]:
                a_, b_ = idx[(k, c)], idx[(p_, c)]
                w(f"      {{ const double tr = zr[{a_} * SV], ti = zi[{a_} * SV]; zr[{a_} * SV] = zr[{b_} * SV]; zi[{a_} * SV] = zi[{b_} * SV]; "
                  f"zr[{b_} * SV] = tr; zi[{b_} * SV] = ti; }}")
            w("    }")
        elif len(cand) > 1:
            w(f"    const double d = {cabs1(idx[(k, k)])};")
            terms = " | ".join(f"({cabs1(idx[(i, k)])} > d)" for i in cand[1:])
            w(f"    bad |= {terms};")
        w(f"    const double pr = zr[{idx[(k, k)]} * SV], pi = zi[{idx[(k, k)]} * SV];")
        w(f"    bad |= !(fabs(pr) + fabs(pi) >= 4.0 * DBL_MIN);")
        if k < N - 1 and s_["rows"]:
            w(f"    double rr, ri; zdiv(1.0, 0.0, pr, pi, rr, ri);")
            for i in s_["rows"]:
                e = idx[(i, k)]
                w(f"    const double lr{i} = rr * zr[{e} * SV] - ri * zi[{e} * SV], li{i} = rr * zi[{e} * SV] + ri * zr[{e} * SV]; "
                  f"zr[{e} * SV] = lr{i}; zi[{e} * SV] = li{i};")
            for c in s_["cols"]:
                e = idx[(k, c)]
                w(f"    {{ const double tr = -zr[{e} * SV], ti = -zi[{e} * SV];")
                for i in s_["rows"]:
                    q = idx[(i, c)]
                    w(f"      zr[{q} * SV] = zr[{q} * SV] + (lr{i} * tr - li{i} * ti); zi[{q} * SV] = zi[{q} * SV] + (lr{i} * ti + li{i} * tr);")
                w("    }")
        w("  }")
    w("  swaps = sw;")
    w("  return bad;")
    w("}")
    w("")
    w("// ZGETRS('N'), one right-hand side (br + i bi): ZLASWP, ZTRSM(L,L,N,U), ZTRSM(L,U,N,N)")
    w("template <int SV> CBM4_CELL_FN void zlu_solve(const double* __restrict__ zr, const double* __restrict__ zi, const unsigned swaps, "
      "double* __restrict__ br, double* __restrict__ bi) {")
    for b, (k, p_) in enumerate(SWAPS):
        w(f"  if (swaps & {1 << b}u) {{ const double tr = br[{k} * SV], ti = bi[{k} * SV]; br[{k} * SV] = br[{p_} * SV]; bi[{k} * SV] = bi[{p_} * SV]; "
          f"br[{p_} * SV] = tr; bi[{p_} * SV] = ti; }}")
    for k in range(N - 1):
        rows = [i for i in range(k + 1, N) if P[i][k]]
        if not rows:
            continue
        w(f"  {{ const double kr = br[{k} * SV], ki = bi[{k} * SV];")
        for i in rows:
            e = idx[(i, k)]
            w(f"    br[{i} * SV] = br[{i} * SV] - (kr * zr[{e} * SV] - ki * zi[{e} * SV]); bi[{i} * SV] = bi[{i} * SV] - (kr * zi[{e} * SV] + ki * zr[{e} * SV]);")
        w("  }")
    for k in range(N - 1, -1, -1):
        rows = [i for i in range(k) if P[i][k]]
        d = idx[(k, k)]
        w(f"  {{ double kr, ki; zdiv(br[{k} * SV], bi[{k} * SV], zr[{d} * SV], zi[{d} * SV], kr, ki); br[{k} * SV] = kr; bi[{k} * SV] = ki;")
        for i in rows:
            e = idx[(i, k)]
            w(f"    br[{i} * SV] = br[{i} * SV] - (kr * zr[{e} * SV] - ki * zi[{e} * SV]); bi[{i} * SV] = bi[{i} * SV] - (kr * zi[{e} * SV] + ki * zr[{e} * SV]);")
        w("  }")
    w("}")
    w("")

    # ------------------------------------------------------------------ DGETRS('T') for the backward sweep
    w("// x <- (P L U)^-T x with the vector in registers (backward sweep: lane = adjoint column; lu/dinv are per-cell")
    w("// arrays read with warp-uniform addresses).  U^T and L^T substitutions in axpy form, then the interchanges in")
    w("// reverse order (DGETRS('T'): DTRSM(L,U,T,N), DTRSM(L,L,T,U), DLASWP(-1)).  Uses fma and 1/U(k,k): agrees with")
    w("// netlib to rounding, not bitwise.")
    w("CBM4_CELL_INL void lut_solve(const double* __restrict__ lu, const double* __restrict__ dinv, const unsigned swaps, "
      "double (&x)[32]) {")
    nf = 0
    for k in range(N):
        cols = [c for c in range(k + 1, N) if P[k][c]]
        w(f"  {{ const double wk = x[{k}] * dinv[{k}]; x[{k}] = wk;")
        for c in cols:
            w(f"    x[{c}] = fma(-lu[{idx[(k, c)]}], wk, x[{c}]);")
            nf += 1
        w("  }")
    for k in range(N - 1, 0, -1):
        cols = [c for c in range(k) if P[k][c]]
        if not cols:
            continue
        w(f"  {{ const double zk = x[{k}];")
        for c in cols:
            w(f"    x[{c}] = fma(-lu[{idx[(k, c)]}], zk, x[{c}]);")
            nf += 1
        w("  }")
    for bi in range(len(SWAPS) - 1, -1, -1):
        k, p_ = SWAPS[bi]
        w(f"  if (swaps & {1 << bi}u) {{ const double t = x[{k}]; x[{k}] = x[{p_}]; x[{p_}] = t; }}")
    w("}")
    w(f"constexpr int LUT_SOLVE_FMA = {nf};")
    w("")
    # the same substitution for ONE interchange outcome: entries of the union pattern that are structural zeros of that
    # outcome are skipped (they only ever add exact zeros), so the result is bit-identical to lut_solve
    def variant_pattern(mask):
        Pm = [[False] * N for _ in range(N)]
        for (r, c) in jent:
            Pm[r][c] = True
        for i in range(N):
            Pm[i][i] = True
        for k in range(N):
            for bi, (kk, p_) in enumerate(SWAPS):
                if kk == k and (mask >> bi) & 1:
                    Pm[k], Pm[p_] = Pm[p_], Pm[k]
            rows = [i for i in range(k + 1, N) if Pm[i][k]]
            cols = [c for c in range(k + 1, N) if Pm[k][c]]
            for i in rows:
                for c in cols:
                    Pm[i][c] = True
        assert all(P[i][c] for i in range(N) for c in range(N) if Pm[i][c])
        return Pm
    DOM = 1   # column 12 <-> row 19 taken, the others not: 98.8 % of the factorisations along CBM-IV trajectories
    Pd = variant_pattern(DOM)
    w(f"// lut_solve for the dominant interchange outcome (swaps == {DOM}): {sum(sum(r) for r in Pd)} instead of {NLU} entries")
    w(f"constexpr unsigned LUT_DOM_SWAPS = {DOM}u;")
    w("CBM4_CELL_INL void lut_solve_dom(const double* __restrict__ lu, const double* __restrict__ dinv, double (&x)[32]) {")
    nfd = 0
    for k in range(N):
        cols = [c for c in range(k + 1, N) if Pd[k][c]]
        w(f"  {{ const double wk = x[{k}] * dinv[{k}]; x[{k}] = wk;")
        for c in cols:
            w(f"    x[{c}] = fma(-lu[{idx[(k, c)]}], wk, x[{c}]);")
            nfd += 1
        w("  }")
    for k in range(N - 1, 0, -1):
        cols = [c for c in range(k) if Pd[k][c]]
        if not cols:
            continue
        w(f"  {{ const double zk = x[{k}];")
        for c in cols:
            w(f"    x[{c}] = fma(-lu[{idx[(k, c)]}], zk, x[{c}]);")
            nfd += 1
        w("  }")
    for bi in range(len(SWAPS) - 1, -1, -1):
        if (DOM >> bi) & 1:
            k, p_ = SWAPS[bi]
            w(f"  {{ const double t = x[{k}]; x[{k}] = x[{p_}]; x[{p_}] = t; }}")
    w("}")
    w(f"constexpr int LUT_SOLVE_DOM_FMA = {nfd};")
    w("")
    # ------------------------------------------------------------------ adjoint-side helpers
    w("// J(T,Y) values in the entry order of cbm4_dev::JAC_ROW/JAC_COL (by column, then row)")
    w("template <int SV> CBM4_CELL_FN void jac_values(const double* __restrict__ V, const double F, const double* __restrict__ PH, "
      "double* __restrict__ JV) {")
    w("  double B[138];")
    for a in jac:
        if a.lhs.name == "B":
            w(f"  B[{a.lhs.idx[0]-1}] = {fx.emit_c(a.rhs, ref)};")
    jexpr = [None] * len(jent)
    for a in jac:
        if a.lhs.name == "JV":
            r, c = a.lhs.idx[0] - 1, a.lhs.idx[1] - 1
            jexpr[jpos[(r, c)]] = fx.emit_c(a.rhs, ref)
    emit_paired(w, "JV", jexpr)
    w("}")
    # time-dependent entries only
    tset = {jent[i]: n for n, i in enumerate(mech.jtime)}
    need_b = set()

    def refs_of(e, name, acc):
        if isinstance(e, fx.Ref):
            if e.name == name:
                acc.add(e.idx[0] - 1)
        elif isinstance(e, fx.Neg):
            refs_of(e.a, name, acc)
        elif isinstance(e, fx.Bin):
            refs_of(e.a, name, acc)
            refs_of(e.b, name, acc)
    for a in jac:
        if a.lhs.name == "JV" and (a.lhs.idx[0] - 1, a.lhs.idx[1] - 1) in tset:
            refs_of(a.rhs, "B", need_b)
    w("// the NJT entries of J that depend on time (through a photolysis rate), same arithmetic as jac_values;")
    w("// JT[n] is entry cbm4_dev::JAC_TIME[n]")
    w("template <int SV> CBM4_CELL_FN void jac_time(const double* __restrict__ V, const double F, const double* __restrict__ PH, "
      "double* __restrict__ JT) {")
    w("  double B[138];")
    for a in jac:
        if a.lhs.name == "B" and (a.lhs.idx[0] - 1) in need_b:
            w(f"  B[{a.lhs.idx[0]-1}] = {fx.emit_c(a.rhs, ref)};")
    for a in jac:
        if a.lhs.name == "JV":
            rc = (a.lhs.idx[0] - 1, a.lhs.idx[1] - 1)
            if rc in tset:
                w(f"  JT[{tset[rc]}] = {fx.emit_c(a.rhs, ref)};")
    w("}")
    w("// Hessian (:1190-1853)")
    w("template <int SV> CBM4_CELL_FN void hess_values(const double* __restrict__ V, const double F, const double* __restrict__ PH, "
      "double* __restrict__ HESS) {")
    w("  double D2A[59];")
    for a in body("Hessian"):
        w(f"  {ref(a.lhs.name, a.lhs.idx)} = {fx.emit_c(a.rhs, ref)};")
    w("}")
    w("// M'[e] = sum_t HESS[h]*K[c] (CBM4_HessTR_Vec :1869-2023 regrouped as the sparse matrix (Hess x K))")
    w("//         + hg * dJdt[n] on the time-dependent Jacobian entries; entry order = cbm4_dev's MV order.")
    w("// CBM-IV is at most bimolecular: HESS depends on the rate coefficients only (no V, no SUN), i.e. it is a per-call")
    w("// constant vector that the includer provides as CBM4_CELL_HESS(h) (filled once from hess_values).")
    w("template <int SV> CBM4_CELL_FN void build_M(const double* __restrict__ K, "
      "const double* __restrict__ DJ, const double hg, double* __restrict__ MV) {")
    jt_of = {jent[i]: n for n, i in enumerate(mech.jtime)}
    mexpr = []
    for e, kr in enumerate(mech.mpat):
        terms = mech.M.get(kr, [])
        sx = None
        for (h, c) in terms:
            t = f"CBM4_CELL_HESS({h}) * K[{c} * SV]"
            sx = t if sx is None else f"({sx} + {t})"
        if kr in jt_of:
            t = f"hg * DJ[{jt_of[kr]}]"
            sx = t if sx is None else f"({t} + {sx})"
        mexpr.append(sx if sx is not None else "0.0")
    emit_paired(w, "MV", mexpr)
    w("}")
    w("// AP[p] = ARP_p(Y_i) (dFun_dRcoeff :1135-1175) + sum_t JVRP_t(Y0)*K[col_t] (dJac_dRcoeff :2330-2380), p = 0..28")

    def l1_expr(ent, vec):
        lit, rct, ops = ent
        assert rct < 0
        e = fx.c_double(lit)
        for op in ops:
            if op == gd.OP_ONE:
                continue
            e = f"({e} * {'F' if op == gd.OP_F else ('2.0' if op == gd.OP_TWO else f'{vec}[{op} * SV]')})"
        return e
    w("template <int SV> CBM4_CELL_FN void param_terms(const double* __restrict__ YI, const double* __restrict__ Y0, const double F, "
      "const double* __restrict__ K, double* __restrict__ AP) {")
    pexpr = []
    for p in range(gd.NP):
        j = gd.JCOEFF[p]
        e = l1_expr(mech.ARP[j - 1], "YI")
        for k in range(mech.CROW[j - 1] - 1, mech.CROW[j] - 1):
            e = f"({e} + {l1_expr(mech.JVRP[k], 'Y0')} * K[{mech.ICOL[k] - 1} * SV])"
        pexpr.append(e)
    emit_paired(w, "AP", pexpr)
    w("}")
    # ------------------------------------------------------------------ Runge-Kutta adjoint helpers
    w("")
    w("// ---- operators of the Runge-Kutta discrete adjoint (rk_DadjInt, ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:1229-1524)")
    w("// AP[p] = ARP_p(Y): reactant product of reaction JCOEFF[p]; column p of JACP is AP[p] * STOICM(:, JCOEFF[p])")
    w("// (dFun_dRcoeff, cbm4_Parameters.F90:1121-1161)")
    w("template <int SV> CBM4_CELL_FN void arp_values(const double* __restrict__ YI, const double F, double* __restrict__ AP) {")
    pexpr = [l1_expr(mech.ARP[gd.JCOEFF[p] - 1], "YI") for p in range(gd.NP)]
    emit_paired(w, "AP", pexpr)
    w("}")
    w("// Division by a complex pivot c + i e in Smith's form (zdiv) splits into a pivot-only part — ratio = c/e or e/c, div =")
    w("// c*ratio + e or e*ratio + c, and which branch — and two real divisions by div.  zdiag_prep tabulates the pivot-only part")
    w("// of the 32 diagonal entries (contiguous outputs) plus RN(1/div); bit k of the result = branch |c| < |e| taken.")
    w("template <int SV> CBM4_CELL_FN unsigned zdiag_prep(const double* __restrict__ zr, const double* __restrict__ zi, double* __restrict__ ZRATIO, "
      "double* __restrict__ ZDIV, double* __restrict__ ZRDIV) {")
    w("  unsigned mask = 0u;")
    for k0 in range(0, N, 4):
        w("  {")
        for j in range(4):
            d = diag[k0 + j]
            w(f"    const double c{j} = zr[{d} * SV], e{j} = zi[{d} * SV]; const bool f{j} = fabs(c{j}) < fabs(e{j});")
            w(f"    const double t{j} = f{j} ? c{j} / e{j} : e{j} / c{j}; const double d{j} = f{j} ? c{j} * t{j} + e{j} : e{j} * t{j} + c{j};")
            w(f"    if (f{j}) mask |= {1 << (k0 + j)}u;")
        w(f"    CBM4_CELL_ST4(ZRATIO + {k0}, t0, t1, t2, t3);")
        w(f"    CBM4_CELL_ST4(ZDIV + {k0}, d0, d1, d2, d3);")
        w(f"    CBM4_CELL_ST4(ZRDIV + {k0}, 1.0 / d0, 1.0 / d1, 1.0 / d2, 1.0 / d3);")
        w("  }")
    w("  return mask;")
    w("}")
    w("// a / b, correctly rounded, from rb = RN(1/b): one Newton correction with exact residual (Markstein); three operations")
    w("// instead of a full division.  Normal-range operands only.")
    w("CBM4_CELL_INL double exact_div(const double a, const double b, const double rb) {")
    w("  const double q = a * rb;")
    w("  return fma(fma(-q, b, a), rb, q);")
    w("}")
    w("// out = J^T x  (LSS_Mul_Jactr1..3), Jv in the JAC_ROW/JAC_COL entry order")
    w("CBM4_CELL_INL void jt_vec(const double* __restrict__ Jv, const double (&x)[32], double (&out)[32]) {")
    for c in range(N):
        ents = [(e, rc[0]) for e, rc in enumerate(jent) if rc[1] == c]
        w("  { double a0 = 0.0, a1 = 0.0;")
        for t, (e, r) in enumerate(ents):
            w(f"    a{t % 2} = fma(Jv[{e}], x[{r}], a{t % 2});")
        w(f"    out[{c}] = a0 + a1; }}")
    w("}")
    w("// (xr + i xi) <- (P L U)^-T (xr + i xi) for the complex factors of zlu_factor (ZGETRS('T'): plain transpose).")
    w("// zdr + i zdi = 1/U(k,k).  Same structure as lut_solve; fma arithmetic, agrees with netlib to rounding.")
    w("CBM4_CELL_INL void zlut_solve(const double* __restrict__ zr, const double* __restrict__ zi, const double* __restrict__ zdr, "
      "const double* __restrict__ zdi, const unsigned swaps, double (&xr)[32], double (&xi)[32]) {")
    for k in range(N):
        cols = [c for c in range(k + 1, N) if P[k][c]]
        w(f"  {{ const double wr = fma(xr[{k}], zdr[{k}], -(xi[{k}] * zdi[{k}])), wi = fma(xr[{k}], zdi[{k}], xi[{k}] * zdr[{k}]); xr[{k}] = wr; xi[{k}] = wi;")
        for c in cols:
            e = idx[(k, c)]
            w(f"    xr[{c}] = fma(-zr[{e}], wr, fma(zi[{e}], wi, xr[{c}])); xi[{c}] = fma(-zr[{e}], wi, fma(-zi[{e}], wr, xi[{c}]));")
        w("  }")
    for k in range(N - 1, 0, -1):
        cols = [c for c in range(k) if P[k][c]]
        if not cols:
            continue
        w(f"  {{ const double wr = xr[{k}], wi = xi[{k}];")
        for c in cols:
            e = idx[(k, c)]
            w(f"    xr[{c}] = fma(-zr[{e}], wr, fma(zi[{e}], wi, xr[{c}])); xi[{c}] = fma(-zr[{e}], wi, fma(-zi[{e}], wr, xi[{c}]));")
        w("  }")
    for bi in range(len(SWAPS) - 1, -1, -1):
        k, p_ = SWAPS[bi]
        w(f"  if (swaps & {1 << bi}u) {{ double t = xr[{k}]; xr[{k}] = xr[{p_}]; xr[{p_}] = t; t = xi[{k}]; xi[{k}] = xi[{p_}]; xi[{p_}] = t; }}")
    w("}")
    # ------------------------------------------------------------------ reference-order variants (RK adjoint)
    w("")
    w("// ---- reference-order variants for the Runge-Kutta adjoint sweep.  Its update Lambda <- Lambda + U1 + U2 + U3 cancels")
    w("// up to nine digits for stiff components, so rounding differences of a reordered evaluation are amplified beyond the")
    w("// 1e-9 parity bound; these variants perform exactly the reference's operations (no fma, true divisions) in its order:")
    w("// matmul(transpose(fjac), g) as sequential dot products; DGETRS('T') = DTRSM(L,U,T,N) + DTRSM(L,L,T,U) + DLASWP(-1) in")
    w("// dot-product form; ZGETRS('T') likewise with four-multiplication products and Smith division.  Skipped structural")
    w("// zeros add exact zeros.  Compile without FMA contraction.")
    w("CBM4_CELL_INL void jt_vec_seq(const double* __restrict__ Jv, const double (&x)[32], double (&out)[32]) {")
    for c in range(N):
        ents = [(e, rc[0]) for e, rc in enumerate(jent) if rc[1] == c]
        w("  { double a = 0.0;")
        for (e, r) in ents:
            w(f"    a = a + Jv[{e}] * x[{r}];")
        w(f"    out[{c}] = a; }}")
    w("}")
    w("CBM4_CELL_INL void lut_solve_seq(const double* __restrict__ lu, const double* __restrict__ dinv, const unsigned swaps, double (&x)[32]) {")
    for i in range(N):
        rows = [k for k in range(i) if P[k][i]]
        w(f"  {{ double t = x[{i}];")
        for k in rows:
            w(f"    t = t - lu[{idx[(k, i)]}] * x[{k}];")
        w(f"    x[{i}] = exact_div(t, lu[{idx[(i, i)]}], dinv[{i}]); }}")
    for i in range(N - 1, -1, -1):
        rows = [k for k in range(i + 1, N) if P[k][i]]
        if not rows:
            continue
        w(f"  {{ double t = x[{i}];")
        for k in rows:
            w(f"    t = t - lu[{idx[(k, i)]}] * x[{k}];")
        w(f"    x[{i}] = t; }}")
    for bi in range(len(SWAPS) - 1, -1, -1):
        k, p_ = SWAPS[bi]
        w(f"  if (swaps & {1 << bi}u) {{ const double t = x[{k}]; x[{k}] = x[{p_}]; x[{p_}] = t; }}")
    w("}")
    w("CBM4_CELL_INL void zlut_solve_seq(const double* __restrict__ zr, const double* __restrict__ zi, const double* __restrict__ zratio, "
      "const double* __restrict__ zdiv, const double* __restrict__ zrdiv, const unsigned zmask, const unsigned swaps, "
      "double (&xr)[32], double (&xi)[32]) {")
    for i in range(N):
        rows = [k for k in range(i) if P[k][i]]
        w(f"  {{ double tr = xr[{i}], ti = xi[{i}];")
        for k in rows:
            e = idx[(k, i)]
            w(f"    {{ const double pr = zr[{e}] * xr[{k}] - zi[{e}] * xi[{k}], pi = zr[{e}] * xi[{k}] + zi[{e}] * xr[{k}]; tr = tr - pr; ti = ti - pi; }}")
        w(f"    const double rt = zratio[{i}]; const bool f = (zmask >> {i}) & 1u;")
        w(f"    const double nr = f ? tr * rt + ti : ti * rt + tr, ni = f ? ti * rt - tr : ti - tr * rt;")
        w(f"    xr[{i}] = exact_div(nr, zdiv[{i}], zrdiv[{i}]); xi[{i}] = exact_div(ni, zdiv[{i}], zrdiv[{i}]); }}")
    for i in range(N - 1, -1, -1):
        rows = [k for k in range(i + 1, N) if P[k][i]]
        if not rows:
            continue
        w(f"  {{ double tr = xr[{i}], ti = xi[{i}];")
        for k in rows:
            e = idx[(k, i)]
            w(f"    {{ const double pr = zr[{e}] * xr[{k}] - zi[{e}] * xi[{k}], pi = zr[{e}] * xi[{k}] + zi[{e}] * xr[{k}]; tr = tr - pr; ti = ti - pi; }}")
        w(f"    xr[{i}] = tr; xi[{i}] = ti; }}")
    for bi in range(len(SWAPS) - 1, -1, -1):
        k, p_ = SWAPS[bi]
        w(f"  if (swaps & {1 << bi}u) {{ double t = xr[{k}]; xr[{k}] = xr[{p_}]; xr[{p_}] = t; t = xi[{k}]; xi[{k}] = xi[{p_}]; xi[{p_}] = t; }}")
    w("}")
    # ------------------------------------------------------------------ reference-order variants (SDIRK TLM)
    w("// reference-order operators of the SDIRK tangent-linear sweep (same reason as above: Y_tlm <- Y_tlm + sum d_i Z_tlm_i")
    w("// cancels for stiff components): matmul(fjac1, g) as sequential row sums, DGETRS('N') = DLASWP + DTRSM(L,L,N,U) +")
    w("// DTRSM(L,U,N,N) in netlib's column-sweep order with the quotient by U(k,k) through exact_div (dinv = RN(1/U(k,k))).")
    w("CBM4_CELL_INL void j_vec_seq(const double* __restrict__ Jv, const double (&x)[32], double (&out)[32]) {")
    for r in range(N):
        ents = sorted([(rc[1], e) for e, rc in enumerate(jent) if rc[0] == r])
        w("  { double a = 0.0;")
        for (c, e) in ents:
            w(f"    a = a + Jv[{e}] * x[{c}];")
        w(f"    out[{r}] = a; }}")
    w("}")
    w("CBM4_CELL_INL void lun_solve_seq(const double* __restrict__ lu, const double* __restrict__ dinv, const unsigned swaps, double (&b)[32]) {")
    for bi, (k, p_) in enumerate(SWAPS):
        w(f"  if (swaps & {1 << bi}u) {{ const double t = b[{k}]; b[{k}] = b[{p_}]; b[{p_}] = t; }}")
    for k in range(N - 1):
        rows = [i for i in range(k + 1, N) if P[i][k]]
        if not rows:
            continue
        w(f"  {{ const double bk = b[{k}];")
        for i in rows:
            w(f"    b[{i}] = b[{i}] - bk * lu[{idx[(i, k)]}];")
        w("  }")
    for k in range(N - 1, -1, -1):
        rows = [i for i in range(k) if P[i][k]]
        w(f"  {{ const double bk = exact_div(b[{k}], lu[{idx[(k, k)]}], dinv[{k}]); b[{k}] = bk;")
        for i in rows:
            w(f"    b[{i}] = b[{i}] - bk * lu[{idx[(i, k)]}];")
        w("  }")
    w("}")
    # ------------------------------------------------------------------ tangent-linear sweep (lane = TLM column)
    w("")
    w("// ---- operators of the tangent-linear sweep (ros_TLM_Int, TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:551-633): the same")
    w("// per-stage operands as the backward sweep, applied NON-transposed to a register vector.")
    w("// out = J x  (LSS_Mul_Jac / LSS_Mul_Jac1), Jv in the JAC_ROW/JAC_COL entry order")
    w("CBM4_CELL_INL void j_vec(const double* __restrict__ Jv, const double (&x)[32], double (&out)[32]) {")
    for r in range(N):
        ents = [(e, rc[1]) for e, rc in enumerate(jent) if rc[0] == r]
        w("  { double a0 = 0.0, a1 = 0.0;")
        for t, (e, c) in enumerate(ents):
            w(f"    a{t % 2} = fma(Jv[{e}], x[{c}], a{t % 2});")
        w(f"    out[{r}] = a0 + a1; }}")
    w("}")
    w("// out += (Hess x K) x  (Hess_Vec(K, x): the matrix M of build_M with hg = 0, applied non-transposed), Mv in MV order")
    w("CBM4_CELL_INL void m_vec_add(const double* __restrict__ Mv, const double (&x)[32], double (&out)[32]) {")
    for k in range(N):
        ents = [(e, kr[1]) for e, kr in enumerate(mech.mpat) if kr[0] == k]
        if not ents:
            continue
        w("  { double a0 = 0.0, a1 = 0.0;")
        for t, (e, r) in enumerate(ents):
            w(f"    a{t % 2} = fma(Mv[{e}], x[{r}], a{t % 2});")
        w(f"    out[{k}] += a0 + a1; }}")
    w("}")
    w("// out += hg * (dJ/dt) x  on the NJT time-dependent Jacobian entries (LSS_Mul_Jac2), Dj[n] = d/dt of entry JAC_TIME[n]")
    w("CBM4_CELL_INL void dj_vec_add(const double* __restrict__ Dj, const double hg, const double (&x)[32], double (&out)[32]) {")
    byrow = {}
    for n, i in enumerate(mech.jtime):
        r, c = jent[i]
        byrow.setdefault(r, []).append((n, c))
    for r, ents in sorted(byrow.items()):
        e = None
        for (n, c) in ents:
            t = f"Dj[{n}] * x[{c}]"
            e = t if e is None else f"fma(Dj[{n}], x[{c}], {e})"
        w(f"  out[{r}] = fma(hg, {e}, out[{r}]);")
    w("}")
    w("// x <- (P L U)^-1 x with the vector in registers: DGETRS('N') on the static pattern with fma and 1/U(k,k)")
    w("CBM4_CELL_INL void lun_solve(const double* __restrict__ lu, const double* __restrict__ dinv, const unsigned swaps, "
      "double (&x)[32]) {")
    for bi, (k, p_) in enumerate(SWAPS):
        w(f"  if (swaps & {1 << bi}u) {{ const double t = x[{k}]; x[{k}] = x[{p_}]; x[{p_}] = t; }}")
    for k in range(N - 1):
        rows = [i for i in range(k + 1, N) if P[i][k]]
        if not rows:
            continue
        w(f"  {{ const double bk = x[{k}];")
        for i in rows:
            w(f"    x[{i}] = fma(-lu[{idx[(i, k)]}], bk, x[{i}]);")
        w("  }")
    for k in range(N - 1, -1, -1):
        rows = [i for i in range(k) if P[i][k]]
        w(f"  {{ const double bk = x[{k}] * dinv[{k}]; x[{k}] = bk;")
        for i in rows:
            w(f"    x[{i}] = fma(-lu[{idx[(i, k)]}], bk, x[{i}]);")
        w("  }")
    w("}")
    # ------------------------------------------------------------------ reference-order variants (ROS_TLM lock-step kernel)
    w("// ---- reference-order operators of the lock-step Rosenbrock tangent-linear kernel (ICNTRL(12) = 1: the columns' error")
    w("// norms decide accept/reject, so their arithmetic follows ros_TLM_Int operation by operation, no fma):")
    w("// tmp = (dJ/dt) x as matmul's column sweep restricted to the NJT time-dependent entries (LSS_Mul_Jac2, :620), then the")
    w("// DAXPY out = out + hg * tmp on the rows that have such entries (the others would add hg * 0)")
    w("CBM4_CELL_INL void dj_vec_seq(const double* __restrict__ Dj, const double hg, const double (&x)[32], double (&out)[32]) {")
    for r, ents in sorted(byrow.items()):
        w("  { double a = 0.0;")
        for (n, c) in sorted(ents, key=lambda nc: nc[1]):
            w(f"    a = a + Dj[{n}] * x[{c}];")
        w(f"    out[{r}] = out[{r}] + hg * a; }}")
    w("}")
    w("// out = Hess_Vec(U1, x) = (f_yy x U1) . x: the term list of CBM4_HessTR_Vec (:1869-2023) contracted the other way, terms")
    w("// of output k in the order of that routine (by HTU row, then textual order) — the order the CPU restatement uses")
    w("CBM4_CELL_INL void hess_vec_seq(const double* __restrict__ U1, const double (&x)[32], double (&out)[32]) {")
    hv = {}
    for a in body("CBM4_HessTR_Vec"):
        r = a.lhs.idx[0] - 1

        def hterms(e):
            if isinstance(e, fx.Bin) and e.op == "+":
                return hterms(e.a) + [e.b]
            if isinstance(e, fx.Num):
                return []
            return [e]
        for t in hterms(a.rhs):
            hv.setdefault(t.b.a.idx[0] - 1, []).append((t.a.idx[0] - 1, r, t.b.b.idx[0] - 1))
    for k in range(N):
        e = None
        for (h, r, c) in hv.get(k, []):
            t = f"CBM4_CELL_HESS({h}) * (U1[{r}] * x[{c}])"
            e = t if e is None else f"({e} + {t})"
        w(f"  out[{k}] = {e if e else '0.0'};")
    w("}")
    # ------------------------------------------------------------------ forward sweep with the right-hand side in registers
    w("// DGETRS('N') with the right-hand side in REGISTERS and the thread's own factors in its strided slots (lu[e * SV],")
    w("// dinv[k * SV] = RN(1/U(k,k)) as lu_factor leaves it): same operations in the same order as lu_solve, the quotient by")
    w("// U(k,k) through exact_div (the correctly rounded quotient in three operations instead of a division sequence)")
    w("// exact_div with a guard: outside the range where Markstein's correction is proven (tiny, huge or non-finite quotients)")
    w("// the true division is used")
    w("CBM4_CELL_INL double exact_div_safe(const double a, const double b, const double rb) {")
    w("  const double q = a * rb;")
    w("  const double aq = fabs(q);")
    w("  if (aq != 0.0 && !(aq > 1.0e-280 && aq < 1.0e280)) return a / b;")
    w("  return fma(fma(-q, b, a), rb, q);")
    w("}")
    w("template <int SV> CBM4_CELL_INL void lu_solve_reg(const double* __restrict__ lu, const double* __restrict__ dinv, "
      "const unsigned swaps, double (&b)[32]) {")
    for bi, (k, p_) in enumerate(SWAPS):
        w(f"  if (swaps & {1 << bi}u) {{ const double t = b[{k}]; b[{k}] = b[{p_}]; b[{p_}] = t; }}")
    for k in range(N - 1):
        rows = [i for i in range(k + 1, N) if P[i][k]]
        if not rows:
            continue
        w(f"  {{ const double bk = b[{k}];")
        for i in rows:
            w(f"    b[{i}] = b[{i}] - bk * lu[{idx[(i, k)]} * SV];")
        w("  }")
    for k in range(N - 1, -1, -1):
        rows = [i for i in range(k) if P[i][k]]
        w(f"  {{ const double bk = exact_div_safe(b[{k}], lu[{idx[(k, k)]} * SV], dinv[{k} * SV]); b[{k}] = bk;")
        for i in rows:
            w(f"    b[{i}] = b[{i}] - bk * lu[{idx[(i, k)]} * SV];")
        w("  }")
    w("}")
    w("}  // namespace cbm4_cell")

    outdir = os.path.join(HERE, "..", "fatode_b200", "csrc", "gen")
    open(os.path.join(outdir, "cbm4_cell_gen.h"), "w").write("\n".join(o) + "\n")
    nl = sum(1 for (i, c) in idx if c < i)
    print(f"NLU={NLU} (L {nl}, U {NLU - nl - N}) updates={nupd} lines={len(o)}")


if __name__ == "__main__":
    main()
