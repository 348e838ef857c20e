#!/usr/bin/env python3
"""Generate fatode_b200/csrc/gen/cbm4_dev_gen.h — the CBM-IV mechanism laid out for
warp-cooperative evaluation on the GPU (PRODUCT side; no oracle code involved).

The KPP-generated module of the reference (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90)
is read as *data*: every routine is a two-level "sum of products"
    level 1:  P_m = ((c_m * o1) * o2) * o3      c_m literal or RCONST(r); o_k in {V(i), F(1), 1, 2}
    level 2:  R_e = sum_t s_{e,t} * P_{m(e,t)}  (textual order kept -> bit-faithful to the Fortran)
which is what lets 32 lanes share one evaluation instead of each thread running 600 scalar flops.
Emitted:
  * L1/L2 tables for FUN (A -> Vdot, :320-449), JAC (B -> JV, :451-1016), Hessian (D2A -> HESS,
    :1190-1853), reactant products and their Jacobian for the 29 rate-coefficient parameters
    (ReactantProd :1028-1121, JacReactantProd :2035-2316, CROW/ICOL_JVRP :20-41, STOICM :45-187),
    lane-interleaved so table reads coalesce;
  * the rate-coefficient recipe of Update_RCONST (:227-310): Arrhenius pairs (evaluated once on
    the host), photolysis factors (x SUN on the device);
  * static sparse per-lane code: rows of J^T x, of (Hess x K)^T x (CBM4_HessTR_Vec :1869-2023
    regrouped as a sparse matrix M[k][r] = sum HESS(h)*K(c)), and stoichiometric column sums
    for the parameter sensitivities.
REAL(4) literals keep their single-precision rounding.
"""
import os
import re
import sys
from collections import OrderedDict, defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90expr as fx  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90")
N, NREACT, NP = 32, 81, 29
JCOEFF = [4, 11, 18, 21, 24, 25, 36, 37, 41, 44, 49, 50, 52, 54, 55,
          59, 64, 65, 66, 67, 68, 70, 73, 75, 76, 77, 78, 79, 81]      # cbm4_ros_adj_dr.F90:31-32
OP_F, OP_ONE, OP_TWO = 32, 33, 34   # operand codes (0..31 = V index)


# ---------------------------------------------------------------- parsing helpers
def flatten_product(e):
    """left-assoc product chain -> list of leaf factors"""
    if isinstance(e, fx.Bin) and e.op == "*":
        return flatten_product(e.a) + [e.b]
    return [e]


def parse_l1(e):
    """-> (lit_value or None, rct_index or -1, [op codes])"""
    fac = flatten_product(e)
    lit, rct, ops = None, -1, []
    first = fac[0]
    rest = fac[1:]
    if isinstance(first, fx.Num):
        lit = first.value
    elif isinstance(first, fx.Ref) and first.name == "RCT":
        rct = first.idx[0] - 1
    else:
        lit = 1.0
        rest = fac
    for f in rest:
        if isinstance(f, fx.Num):
            assert f.kind == "int" and f.value == 2.0, f
            ops.append(OP_TWO)
        elif isinstance(f, fx.Ref) and f.name == "V":
            ops.append(f.idx[0] - 1)
        elif isinstance(f, fx.Ref) and f.name == "F":
            assert f.idx == (1,)
            ops.append(OP_F)
        else:
            raise ValueError(f)
    assert len(ops) <= 3, ops
    while len(ops) < 3:
        ops.append(OP_ONE)
    return lit, rct, ops


def linear_terms(e, src_name):
    """sum chain  [-]t1 +/- t2 ...  with t = [num*]SRC(i)  ->  [(signed coef, i-1)]"""
    def leaf(t, sign):
        if isinstance(t, fx.Neg):
            return leaf(t.a, -sign)
        if isinstance(t, fx.Ref) and t.name == src_name:
            return (sign * 1.0, t.idx[0] - 1)
        if isinstance(t, fx.Bin) and t.op == "*" and isinstance(t.a, fx.Num) and isinstance(t.b, fx.Ref) \
                and t.b.name == src_name:
            return (sign * t.a.value, t.b.idx[0] - 1)
        raise ValueError(t)

    if isinstance(e, fx.Bin) and e.op in "+-":
        return linear_terms(e.a, src_name) + [leaf(e.b, 1.0 if e.op == "+" else -1.0)]
    if isinstance(e, fx.Num) and e.value == 0.0:
        return []
    return [leaf(e, 1.0)]


class SumProd:
    """one two-level routine"""

    def __init__(self, name, l1, l2, out_pos, n_out):
        self.name, self.l1, self.l2, self.out_pos, self.n_out = name, l1, l2, out_pos, n_out


def schedule_l2(l2_lists):
    """Assign outputs to 32 lanes (longest first, least-loaded lane) and flatten each lane's work
    into a sequence of (scoef, src, first, last, out).  Returns (niter, per-iter-per-lane rows)."""
    order = sorted(range(len(l2_lists)), key=lambda e: -len(l2_lists[e]))
    load = [0] * 32
    seqs = [[] for _ in range(32)]
    for e in order:
        terms = l2_lists[e]
        if not terms:
            continue
        lane = min(range(32), key=lambda l: (load[l], l))
        for t, (sc, src) in enumerate(terms):
            seqs[lane].append((sc, src, t == 0, t == len(terms) - 1, e))
        load[lane] += len(terms)
    niter = max(load)
    return niter, seqs


# ---------------------------------------------------------------- main
def parse_mechanism(text):
    """Reads the KPP module text into the two-level sum-of-products description (shared by the warp and the
    per-thread generators).  Returns a namespace with A, VD, B, JV, jent, D2A, HE, M, mpat, jtime, photo, arr,
    ARP, JVRP, CROW, ICOL, CCOL, IROW, st, NJ, NM."""
    body = lambda nm: fx.parse_assignments(fx.routine_body(text, nm))

    # ---- rate coefficients (Update_RCONST)
    arr = {}      # r -> (A0, B0)
    photo = {}    # r -> c
    for l in fx.routine_body(text, "Update_RCONST"):
        m = re.match(r"RCONST\((\d+)\)\s*=\s*\(ARR2\(([^,]+),([^)]+)\)\)", l, re.I)
        if m:
            bt = m.group(3).strip()
            b = fx.parse_number(bt.lstrip("-")).value * (-1 if bt.startswith("-") else 1)
            arr[int(m.group(1)) - 1] = (fx.parse_number(m.group(2).strip()).value, b)
            continue
        m = re.match(r"RCONST\((\d+)\)\s*=\s*\(([^*]+)\*SUN\)", l, re.I)
        if m:
            photo[int(m.group(1)) - 1] = fx.parse_number(m.group(2).strip()).value
    assert len(arr) == 41 and len(photo) == 11

    # ---- FUN
    A = OrderedDict()
    VD = [[] for _ in range(N)]
    for a in body("CBM4_FUN"):
        if a.lhs.name == "A":
            A[a.lhs.idx[0] - 1] = parse_l1(a.rhs)
        elif a.lhs.name == "VDOT":
            VD[a.lhs.idx[0] - 1] = linear_terms(a.rhs, "A")
    assert list(A.keys()) == list(range(NREACT))
    # ---- JAC
    B = OrderedDict()
    JV = OrderedDict()
    for a in body("CBM4_JAC"):
        if a.lhs.name == "B":
            B[a.lhs.idx[0] - 1] = parse_l1(a.rhs)
        elif a.lhs.name == "JV":
            JV[(a.lhs.idx[0] - 1, a.lhs.idx[1] - 1)] = linear_terms(a.rhs, "B")
    bpos = {m: i for i, m in enumerate(sorted(B.keys()))}      # B has gaps (dA/dF skipped)
    JV = OrderedDict((rc, [(sc, bpos[src]) for sc, src in lst]) for rc, lst in JV.items())
    B = OrderedDict((bpos[m], v) for m, v in B.items())
    # entry order: by column (= output row of J^T x), then row
    jent = sorted(JV.keys(), key=lambda rc: (rc[1], rc[0]))
    NJ = len(jent)
    # ---- Hessian
    D2A = OrderedDict()
    HE = OrderedDict()
    for a in body("Hessian"):
        if a.lhs.name == "D2A":
            D2A[a.lhs.idx[0] - 1] = parse_l1(a.rhs)
        elif a.lhs.name == "HESS":
            HE[a.lhs.idx[0] - 1] = linear_terms(a.rhs, "D2A")
    assert list(HE.keys()) == list(range(258)) and list(D2A.keys()) == list(range(59))
    # ---- HessTR_Vec: HTU(r) = sum HESS(h)*(U1(k)*U2(c))  ->  M[k][r] = sum HESS(h)*U2(c)
    M = defaultdict(list)   # (k, r) -> [(h, c)]
    for a in body("CBM4_HessTR_Vec"):
        r = a.lhs.idx[0] - 1

        def terms(e):
            if isinstance(e, fx.Bin) and e.op == "+":
                return terms(e.a) + [e.b]
            if isinstance(e, fx.Num):
                return []
            return [e]
        for t in terms(a.rhs):
            assert isinstance(t, fx.Bin) and t.op == "*" and t.a.name == "HESS"
            u = t.b
            assert isinstance(u, fx.Bin) and u.op == "*" and u.a.name == "U1" and u.b.name == "U2"
            M[(u.a.idx[0] - 1, r)].append((t.a.idx[0] - 1, u.b.idx[0] - 1))
    # J entries that depend on time (through a photolysis rate): they carry the H*gamma*dJ/dt term
    photoB = {m for m, (lit, rct, ops) in B.items() if rct in photo}
    jtime = [i for i, rc in enumerate(jent) if any(src in photoB for _, src in JV[rc])]
    # M' pattern = pattern(M) U time-dependent J entries; ordered by output row r, then k
    mpat = sorted(set(M.keys()) | {jent[i] for i in jtime}, key=lambda kr: (kr[1], kr[0]))
    NM = len(mpat)
    # ---- parameter terms
    ARP = {a.lhs.idx[0] - 1: parse_l1(a.rhs) for a in body("ReactantProd")}
    JVRP = {a.lhs.idx[0] - 1: parse_l1(a.rhs) for a in body("JacReactantProd")}
    tab = lambda nm: [int(x) for x in re.findall(r"-?\d+", re.sub("&", " ", re.search(
        r"::\s*" + nm + r"\s*=\s*\(/(.*?)/\)", text, re.S | re.I).group(1)))]
    CROW, ICOL, CCOL, IROW = tab("CROW_JVRP"), tab("ICOL_JVRP"), tab("CCOL_STOICM"), tab("IROW_STOICM")
    st = []
    for nm in ("STOICM_0", "STOICM_1", "STOICM_2"):
        bodytxt = re.sub("&", " ", re.search(r"::\s*" + nm + r"\s*=\s*\(/(.*?)/\)", text, re.S | re.I).group(1))
        for x in bodytxt.split(","):
            x = x.strip()
            if x:
                st.append(-fx.parse_number(x[1:]).value if x.startswith("-") else fx.parse_number(x).value)
    assert len(st) == 329

    import types
    return types.SimpleNamespace(**{k: v for k, v in locals().items() if k not in ("text", "body", "tab", "terms")})


def main():
    text = open(SRC).read()
    mech = parse_mechanism(text)
    arr = mech.arr
    photo = mech.photo
    A = mech.A
    VD = mech.VD
    B = mech.B
    JV = mech.JV
    bpos = mech.bpos
    jent = mech.jent
    NJ = mech.NJ
    D2A = mech.D2A
    HE = mech.HE
    M = mech.M
    photoB = mech.photoB
    jtime = mech.jtime
    mpat = mech.mpat
    NM = mech.NM
    ARP = mech.ARP
    JVRP = mech.JVRP
    CROW = mech.CROW
    ICOL = mech.ICOL
    CCOL = mech.CCOL
    IROW = mech.IROW
    st = mech.st

    o = []
    w = o.append
    w("// GENERATED by tools/gen_cbm4_device.py from the reference's CBM-IV mechanism")
    w("// (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90) — do not edit.  PRODUCT side:")
    w("// mechanism tables for warp-cooperative evaluation + static sparse per-lane code.")
    w("#pragma once")
    w("#ifdef __CUDACC__")
    w("#define CBM4_TAB static __device__ const")
    w("#else")
    w("#define CBM4_TAB static const")
    w("#endif")
    w("namespace cbm4_dev {")
    w(f"constexpr int NVAR = {N}, NREACT = {NREACT}, NPAR = {NP}, NJ = {NJ}, NM = {NM}, NHESS = 258, NJT = {len(jtime)};")
    w("constexpr int OP_F = 32, OP_ONE = 33, OP_TWO = 34;")
    w("struct L1 { double lit; short rct; unsigned char o1, o2, o3; unsigned char pad[3]; };")
    w("constexpr int TERM_BUF = 353;   // products buffer: MAXT + the zero sentinel")
    w("struct TermSP { double sc; unsigned int off; unsigned int pad; };")

    def emit_l1(name, entries):
        ms = sorted(entries.keys())
        n = len(ms)
        npad = (n + 31) // 32 * 32
        w(f"constexpr int {name}_N = {n}, {name}_NPAD = {npad};")
        w(f"CBM4_TAB L1 {name}_TAB[{npad}] = {{")
        for i in range(npad):
            if i < n:
                lit, rct, ops = entries[ms[i]]
                w(f"  {{{fx.c_double(lit if lit is not None else 0.0)}, {rct}, {ops[0]}, {ops[1]}, {ops[2]}, {{0,0,0}}}},")
            else:
                w(f"  {{0.0, -1, {OP_ONE}, {OP_ONE}, {OP_ONE}, {{0,0,0}}}},")
        w("};")

    def emit_l2(name, lists, fixed_lanes=False, mat=False, out_ids=None):
        """Level 2 as (a) a flat table of products TERM_k = sc_k * P[src_k] evaluated 32 at a time and (b) per
        lane an ordered list of term indices to add up (the Fortran's textual order), so the sequential part
        is one shared-memory load and one DADD per term.  fixed_lanes: output e belongs to lane e (FUN: the
        sum stays in that lane's register).  mat: TERM_k = HS[h_k] * K[c_k] (no literal coefficient)."""
        terms = []        # (sc, src) or (h, c)
        per_out = []      # per output: list of term ids, in order
        for lst in lists:
            ids = []
            for (sc, src) in lst:
                terms.append((sc, src))
                ids.append(len(terms) - 1)
            per_out.append(ids)
        nt = len(terms)
        ntpad = (nt + 31) // 32 * 32
        ZERO = ntpad      # T[ntpad] = 0.0 (sentinel for padded slots)
        seqs = [[] for _ in range(32)]
        if fixed_lanes:
            assert len(lists) == 32
            for e, ids in enumerate(per_out):
                seqs[e] = [(k, False, e) for k in ids]
        else:
            order = sorted(range(len(per_out)), key=lambda e: -len(per_out[e]))
            load = [0] * 32
            for e in order:
                ids = per_out[e]
                if not ids:
                    continue
                lane = min(range(32), key=lambda l: (load[l], l))
                for t, k in enumerate(ids):
                    seqs[lane].append((k, t == len(ids) - 1, e))
                load[lane] += len(ids)
        niter = max(len(q) for q in seqs)
        niter4 = (niter + 3) // 4 * 4
        w(f"constexpr int {name}_NT = {nt}, {name}_NTPAD = {ntpad}, {name}_NITER = {niter4};")
        if mat:
            w(f"CBM4_TAB ushort2 {name}_TERM[{ntpad}] = {{   // (HESS index, K index) byte offsets")
            w("  " + ", ".join(f"{{{8 * int(terms[k][1])}, {8 * int(terms[k][0])}}}" if k < nt else "{0, 0}" for k in range(ntpad)))
        else:
            w(f"CBM4_TAB TermSP {name}_TERM[{ntpad}] = {{   // sc, byte offset of the level-1 product")
            w("  " + ", ".join(f"{{{fx.c_double(terms[k][0])}, {8 * terms[k][1]}, 0}}" if k < nt else "{0.0, 0, 0}"
                               for k in range(ntpad)))
        w("};")
        # ordered sums: ushort per (iter, lane), 4 iterations packed per 64-bit word: [iter/4][lane] as ushort4
        w(f"CBM4_TAB ushort4 {name}_SUM[{niter4 // 4} * 32] = {{   // term index (low 14 bits) | 0x8000 = last term of an output")
        rows = []
        for q4 in range(niter4 // 4):
            row = []
            for lane in range(32):
                vals = []
                for it in range(4 * q4, 4 * q4 + 4):
                    if it < len(seqs[lane]):
                        k, last, e = seqs[lane][it]
                        vals.append(k | (0x8000 if (last and not fixed_lanes) else 0))
                    else:
                        vals.append(ZERO)
                row.append("{" + ", ".join(map(str, vals)) + "}")
            rows.append("  " + ", ".join(row) + ",")
        w("\n".join(rows))
        w("};")
        if not fixed_lanes:
            nout = max(sum(1 for (_, last, _) in q if last) for q in seqs)
            w(f"constexpr int {name}_NOUT = {nout};")
            w(f"CBM4_TAB unsigned short {name}_OUT[{nout} * 32] = {{   // n-th output finished by a lane -> output index")
            for n in range(nout):
                row = []
                for lane in range(32):
                    outs = [e for (_, last, e) in seqs[lane] if last]
                    row.append(str(out_ids[outs[n]] if out_ids else outs[n]) if n < len(outs) else "0")
                w("  " + ", ".join(row) + ",")
            w("};")
        return niter4

    MAXT = 352   # products buffer (doubles) shared by all evaluators; larger routines run in groups

    def emit_l2_groups(name, lists, mat=False):
        """split the outputs (kept in order, global output numbering) into groups of <= MAXT terms"""
        groups, cur, cnt = [], [], 0
        for e, lst in enumerate(lists):
            if cnt + len(lst) > MAXT and cur:
                groups.append(cur)
                cur, cnt = [], 0
            cur.append(e)
            cnt += len(lst)
        if cur:
            groups.append(cur)
        w(f"constexpr int {name}_NG = {len(groups)};")
        its = []
        for g, outs in enumerate(groups):
            sub = [lists[e] for e in outs]
            its.append(emit_l2(f"{name}{g}", sub, mat=mat, out_ids=outs))
        return its

    stats = {}
    # FUN
    emit_l1("FUN_L1", A)
    stats["fun_iter"] = emit_l2("FUN_L2", VD, fixed_lanes=True)
    # JAC: outputs numbered in jent order
    emit_l1("JAC_L1", B)
    stats["jac_iter"] = emit_l2_groups("JAC_L2", [JV[rc] for rc in jent])
    w(f"CBM4_TAB unsigned char JAC_ROW[{NJ}] = {{{', '.join(str(rc[0]) for rc in jent)}}};")
    w(f"CBM4_TAB unsigned char JAC_COL[{NJ}] = {{{', '.join(str(rc[1]) for rc in jent)}}};")
    w(f"CBM4_TAB short JAC_TIME[{len(jtime)}] = {{{', '.join(map(str, jtime))}}};   // J entries depending on SUN")
    # HESS
    emit_l1("HES_L1", D2A)
    stats["hes_iter"] = emit_l2_groups("HES_L2", [HE[h] for h in range(258)])
    # M' assembly: M'_e = sum_t HESS[h]*K[c]  (+ H*gamma*djdt[jt] if the entry is a time-dependent J entry)
    mlists = []
    for kr in mpat:
        mlists.append([(float(c), h) for (h, c) in M.get(kr, [])])      # sc field carries the K index
    stats["m_iter"] = emit_l2_groups("MAT_L2", mlists, mat=True)
    jt_of = {jent[i]: n for n, i in enumerate(jtime)}
    w(f"CBM4_TAB short MAT_JT[{NM}] = {{{', '.join(str(jt_of.get(kr, -1)) for kr in mpat)}}};   // index into JAC_TIME or -1")
    # rate coefficients
    w(f"static const short ARR_IDX[41] = {{{', '.join(str(r) for r in sorted(arr))}}};")
    w(f"static const double ARR_A0[41] = {{{', '.join(fx.c_double(arr[r][0]) for r in sorted(arr))}}};")
    w(f"static const double ARR_B0[41] = {{{', '.join(fx.c_double(arr[r][1]) for r in sorted(arr))}}};")
    w(f"CBM4_TAB short PHOTO_IDX[11] = {{{', '.join(str(r) for r in sorted(photo))}}};")
    w(f"CBM4_TAB double PHOTO_C[11] = {{{', '.join(fx.c_double(photo[r]) for r in sorted(photo))}}};")
    # parameter terms: ARP_p and aj_p = sum_k JVRP(k)*K(ICOL(k))
    w("// parameter p = rate coefficient of reaction JCOEFF[p]; ARP = reactant product (L1 form),")
    w("// AJ term list: aj_p = sum_t JVRP_t(Y0) * K[col_t]  (<= 2 terms here; third slot = padding)")
    w(f"CBM4_TAB short JCOEFF[{NP}] = {{{', '.join(str(j - 1) for j in JCOEFF)}}};")
    w(f"CBM4_TAB L1 ARP_TAB[32] = {{")
    for p in range(32):
        if p < NP:
            lit, rct, ops = ARP[JCOEFF[p] - 1]
            w(f"  {{{fx.c_double(lit)}, {rct}, {ops[0]}, {ops[1]}, {ops[2]}, {{0,0,0}}}},")
        else:
            w(f"  {{0.0, -1, {OP_ONE}, {OP_ONE}, {OP_ONE}, {{0,0,0}}}},")
    w("};")
    maxaj = max(CROW[j] - CROW[j - 1] for j in JCOEFF)
    w(f"constexpr int AJ_NT = {maxaj};")
    w(f"CBM4_TAB L1 AJ_TAB[{maxaj} * 32] = {{   // [term][p]")
    ajc = []
    for t in range(maxaj):
        for p in range(32):
            if p < NP and t < CROW[JCOEFF[p]] - CROW[JCOEFF[p] - 1]:
                k = CROW[JCOEFF[p] - 1] - 1 + t
                lit, rct, ops = JVRP[k]
                w(f"  {{{fx.c_double(lit)}, {rct}, {ops[0]}, {ops[1]}, {ops[2]}, {{0,0,0}}}},")
                ajc.append(ICOL[k] - 1)
            else:
                w(f"  {{0.0, -1, {OP_ONE}, {OP_ONE}, {OP_ONE}, {{0,0,0}}}},")
                ajc.append(0)
    w("};")
    w(f"CBM4_TAB unsigned char AJ_COL[{maxaj} * 32] = {{{', '.join(map(str, ajc))}}};")

    # ---------------- static sparse per-lane code
    w("#ifdef __CUDACC__")
    w("// Row r of  v = J^T x  and  h = M'^T x  for the calling lane's private vector x.")
    w("// Jv / Mv are the per-cell value arrays in shared memory (broadcast reads), entry order =")
    w("// (output row, source row); f(IC<r>, v, h) consumes the pair.")
    w("template <int R> struct IC { static constexpr int value = R; };")
    w("template <class F> __device__ __forceinline__ void jt_mt_rows(const double* __restrict__ Jv,")
    w("    const double* __restrict__ Mv, const double (&x)[32], F&& f) {")
    je = 0
    me = 0
    for r in range(N):
        w("  {")
        w("    double v = 0.0, h = 0.0;")
        while je < NJ and jent[je][1] == r:
            w(f"    v = fma(Jv[{je}], x[{jent[je][0]}], v);")
            je += 1
        while me < NM and mpat[me][1] == r:
            w(f"    h = fma(Mv[{me}], x[{mpat[me][0]}], h);")
            me += 1
        w(f"    f(IC<{r}>(), v, h);")
        w("  }")
    w("}")
    w("// v = J^T x as a register vector (static indices), and rows of h = M'^T x through a callback")
    w("__device__ __forceinline__ void jt_vec(const double* __restrict__ Jv, const double (&x)[32], double (&v)[32]) {")
    je = 0
    for r in range(N):
        terms = []
        while je < NJ and jent[je][1] == r:
            terms.append((je, jent[je][0]))
            je += 1
        # two partial sums per row shorten the dependent FMA chain
        w(f"  {{ double a0 = 0.0, a1 = 0.0;")
        for t, (e, k) in enumerate(terms):
            w(f"    a{t % 2} = fma(Jv[{e}], x[{k}], a{t % 2});")
        w(f"    v[{r}] = a0 + a1; }}")
    w("}")
    w("template <class F> __device__ __forceinline__ void mt_rows(const double* __restrict__ Mv, const double (&x)[32], F&& f) {")
    me = 0
    for r in range(N):
        terms = []
        while me < NM and mpat[me][1] == r:
            terms.append((me, mpat[me][0]))
            me += 1
        w(f"  {{ double a0 = 0.0, a1 = 0.0;")
        for t, (e, k) in enumerate(terms):
            w(f"    a{t % 2} = fma(Mv[{e}], x[{k}], a{t % 2});")
        w(f"    f(IC<{r}>(), a0 + a1); }}")
    w("}")
    # ---- variants for the two-warps-per-cell backward sweep (ros_cell_adj_pair.cuh): fewer instructions per row
    w("// v = J^T x with one product + multiply-add chain per row (no 0.0 seeds, no final add of partial sums)")
    w("__device__ __forceinline__ void jt_vec1(const double* __restrict__ Jv, const double (&x)[32], double (&v)[32]) {")
    je = 0
    for r in range(N):
        terms = []
        while je < NJ and jent[je][1] == r:
            terms.append((je, jent[je][0]))
            je += 1
        assert terms
        if len(terms) <= 4:
            (e0, k0) = terms[0]
            expr = f"Jv[{e0}] * x[{k0}]"
            for (e, k) in terms[1:]:
                expr = f"fma(Jv[{e}], x[{k}], {expr})"
            w(f"  v[{r}] = {expr};")
        else:   # long rows: two chains seeded with products
            w(f"  {{ double a0 = Jv[{terms[0][0]}] * x[{terms[0][1]}], a1 = Jv[{terms[1][0]}] * x[{terms[1][1]}];")
            for t, (e, k) in enumerate(terms[2:]):
                w(f"    a{t % 2} = fma(Jv[{e}], x[{k}], a{t % 2});")
            w(f"    v[{r}] = a0 + a1; }}")
    w("}")
    w("// s_p = sum_r STOICM(r, JCOEFF[p]) * x_r  (column p of the stoichiometric matrix; REAL(4) literals)")
    w("template <class F> __device__ __forceinline__ void stoich_cols(const double (&x)[32], F&& f) {")
    nst = 0
    for p in range(NP):
        j = JCOEFF[p]
        w("  {")
        w("    double s = 0.0;")
        for k in range(CCOL[j - 1], CCOL[j]):
            w(f"    s = fma({fx.c_double(st[k - 1])}, x[{IROW[k - 1] - 1}], s);")
            nst += 1
        w(f"    f(IC<{p}>(), s);")
        w("  }")
    w("}")
    w("#endif  // __CUDACC__")
    w("}  // namespace cbm4_dev")
    outdir = os.path.join(HERE, "..", "fatode_b200", "csrc", "gen")
    os.makedirs(outdir, exist_ok=True)
    open(os.path.join(outdir, "cbm4_dev_gen.h"), "w").write("\n".join(o) + "\n")
    print(f"NJ={NJ} NM={NM} (pattern(M)={len(M)}) time-dep J entries={len(jtime)} stoich nnz={nst} iters={stats}")


if __name__ == "__main__":
    main()
