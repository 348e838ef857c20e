// CBM-IV (32 species, 81 reactions, 29 rate-coefficient parameters) as warp-cooperative device
// functions: the 32 lanes of the warp that owns a cell share ONE evaluation of the mechanism.
//
// Replaces the reference's model callbacks for this example
//   fun / jac / hesstr_vec / jacp / hesstr_vec_f_py / adjinit  (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_ros_adj_dr.F90:2-117)
//   Update_SUN / Update_RCONST / CBM4_FUN / CBM4_JAC / Hessian / CBM4_HessTR_Vec / dFun_dRcoeff / dJac_dRcoeff
//                                                               (cbm4_Parameters.F90:196-2392)
// The mechanism itself comes from gen/cbm4_dev_gen.h (two-level sum-of-products tables, see
// tools/gen_cbm4_device.py).  FUN and JAC keep the Fortran's operation order with unfused
// multiplies/adds; COS comes from portable_math.h (fixed-order, < 1 ulp), so the forward sweep is
// reproducible bit for bit on the host.
#pragma once
#include "fatode_common.cuh"
#include "gen/cbm4_dev_gen.h"
#include "portable_math.h"

namespace fatode {

struct Cbm4Const {
  double rc_base[81];   // RCONST with the Arrhenius entries evaluated on the host (exp from libm), photolysis = 0
  double fix;           // FIX(1)
  double emis;          // emission added to p(26), p(31)  (dr.F90:96-97)
  double temp;
};

struct Cbm4Warp {
  static constexpr int N = 32, NP = 29;
  static constexpr int NJ = cbm4_dev::NJ, NM = cbm4_dev::NM, NHESS = cbm4_dev::NHESS, NJT = cbm4_dev::NJT;
  // model scratch inside the per-warp workspace (doubles)
  static constexpr int OFF_YB = 0;      // [36]  operands: V(0..31), F(1), 1.0, 2.0
  static constexpr int OFF_RC = 36;     // [84]  rate coefficients at the current time
  static constexpr int OFF_P = 120;     // [160] level-1 products
  static constexpr int OFF_T = 280;     // [354] level-2 products sc*P (+ zero sentinel)
  static constexpr int WS_DOUBLES = 634;
  typedef Cbm4Const Const;

  // Update_SUN, cbm4_Parameters.F90:196-220
  __device__ static double sun(double T) {
    const double PI = 3.14159265358979;
    const double SunRise = 4.5, SunSet = 19.5;
    double Thour = __ddiv_rn(T, 3600.0);
    double Tlocal = __dsub_rn(Thour, (double)((((int)Thour) / 24) * 24));
    if (Tlocal >= SunRise && Tlocal <= SunSet) {
      double Ttmp = __ddiv_rn(__dsub_rn(__dsub_rn(__dmul_rn(2.0, Tlocal), SunRise), SunSet), __dsub_rn(SunSet, SunRise));
      if (Ttmp > 0) Ttmp = __dmul_rn(Ttmp, Ttmp); else Ttmp = __dmul_rn(-Ttmp, Ttmp);
      return __ddiv_rn(__dadd_rn(1.0, fpm::cos_cr(__dmul_rn(PI, Ttmp))), 2.0);
    }
    return 0.0;
  }

  // once per cell: constant part of RCONST, constant operands
  __device__ static void init(double* ws, const Const& mc, int lane) {
    for (int r = lane; r < 81; r += 32) ws[OFF_RC + r] = mc.rc_base[r];
    if (lane == 0) { ws[OFF_YB + 32] = mc.fix; ws[OFF_YB + 33] = 1.0; ws[OFF_YB + 34] = 2.0; }
    __syncwarp();
  }
  // Update_SUN(t) + photolysis part of Update_RCONST (:228,235,236,241,250,261,265,266,272,298,301)
  __device__ static void set_time(double* ws, double t, int lane) {
    double s = sun(t);
    __syncwarp();
    if (lane < 11) ws[OFF_RC + cbm4_dev::PHOTO_IDX[lane]] = __dmul_rn(cbm4_dev::PHOTO_C[lane], s);
    __syncwarp();
  }
  // publish the state vector (lane i holds component i)
  __device__ static void set_y(double* ws, double y_lane, int lane) {
    __syncwarp();
    ws[OFF_YB + lane] = y_lane;
    __syncwarp();
  }

  template <int NPAD>
  __device__ static void eval_l1(const cbm4_dev::L1* __restrict__ tab, const double* ws, double* P, int lane) {
    const double* yb = ws + OFF_YB;
    const double* rc = ws + OFF_RC;
#pragma unroll
    for (int b = 0; b < NPAD; b += 32) {
      cbm4_dev::L1 e = tab[b + lane];
      double c = (e.rct >= 0) ? rc[e.rct] : e.lit;
      c = __dmul_rn(c, yb[e.o1]);
      c = __dmul_rn(c, yb[e.o2]);
      c = __dmul_rn(c, yb[e.o3]);
      P[b + lane] = c;
    }
    __syncwarp();
  }
  // level 2, part 1: all products TERM_k = sc_k * P[src_k], 32 per round (independent -> fully parallel)
  template <int NTPAD>
  __device__ static void eval_terms(const cbm4_dev::TermSP* __restrict__ tab, const double* P, double* T, int lane) {
#pragma unroll
    for (int b = 0; b < NTPAD; b += 32) {
      cbm4_dev::TermSP e = tab[b + lane];
      T[b + lane] = __dmul_rn(e.sc, *reinterpret_cast<const double*>(reinterpret_cast<const char*>(P) + e.off));
    }
    if (lane == 0) T[NTPAD] = 0.0;   // sentinel used by padded slots
    __syncwarp();
  }
  // level 2, part 2: ordered sums (the Fortran's textual order).  One output per lane, kept in a register.
  template <int NITER>
  __device__ static double sum_fixed(const ushort4* __restrict__ tab, const double* T, int lane) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < NITER / 4; ++q) {
      const ushort4 w = tab[q * 32 + lane];
      acc = __dadd_rn(acc, T[w.x]);
      acc = __dadd_rn(acc, T[w.y]);
      acc = __dadd_rn(acc, T[w.z]);
      acc = __dadd_rn(acc, T[w.w]);
    }
    return acc;
  }
  // several outputs per lane: bit 15 marks the last term of an output; ACCUM adds to out[] instead of storing
  template <int NITER, bool ACCUM>
  __device__ static void sum_multi(const ushort4* __restrict__ tab, const unsigned short* __restrict__ outmap, const double* T,
                                   double* out, int lane) {
    double acc = 0.0;
    int n = 0;
    auto step = [&](unsigned v) {
      acc = __dadd_rn(acc, T[v & 0x3fffu]);
      if (v & 0x8000u) {
        const int o = outmap[n * 32 + lane];
        if (ACCUM) out[o] += acc; else out[o] = acc;
        acc = 0.0;
        ++n;
      }
    };
#pragma unroll
    for (int q = 0; q < NITER / 4; ++q) {
      const ushort4 w = tab[q * 32 + lane];
      step(w.x); step(w.y); step(w.z); step(w.w);
    }
    __syncwarp();
  }

  // fun (dr.F90:83-98): returns component `lane` of f(t, y); set_time/set_y must have been called
  __device__ static double fun(double* ws, const Const& mc, int lane) {
    eval_l1<cbm4_dev::FUN_L1_NPAD>(cbm4_dev::FUN_L1_TAB, ws, ws + OFF_P, lane);
    eval_terms<cbm4_dev::FUN_L2_NTPAD>(cbm4_dev::FUN_L2_TERM, ws + OFF_P, ws + OFF_T, lane);
    double f = sum_fixed<cbm4_dev::FUN_L2_NITER>(cbm4_dev::FUN_L2_SUM, ws + OFF_T, lane);
    __syncwarp();
    if (lane == 30 || lane == 25) f = __dadd_rn(f, mc.emis);
    return f;
  }
  // jac (dr.F90:69-80): the 276 structural entries, order (col,row) = cbm4_dev::JAC_COL/JAC_ROW
  __device__ static void jac(double* ws, double* jv, int lane) {
    static_assert(cbm4_dev::JAC_L2_NG == 2, "generator emitted a different number of JAC groups");
    eval_l1<cbm4_dev::JAC_L1_NPAD>(cbm4_dev::JAC_L1_TAB, ws, ws + OFF_P, lane);
    eval_terms<cbm4_dev::JAC_L20_NTPAD>(cbm4_dev::JAC_L20_TERM, ws + OFF_P, ws + OFF_T, lane);
    sum_multi<cbm4_dev::JAC_L20_NITER, false>(cbm4_dev::JAC_L20_SUM, cbm4_dev::JAC_L20_OUT, ws + OFF_T, jv, lane);
    eval_terms<cbm4_dev::JAC_L21_NTPAD>(cbm4_dev::JAC_L21_TERM, ws + OFF_P, ws + OFF_T, lane);
    sum_multi<cbm4_dev::JAC_L21_NITER, false>(cbm4_dev::JAC_L21_SUM, cbm4_dev::JAC_L21_OUT, ws + OFF_T, jv, lane);
  }
  // dense scatter of E = ghinv*I - J into a column-major 32x32 tile with leading dimension LD (for the LU)
  template <int LD>
  __device__ static void scatter_E(const double* jv, double ghinv, double* tile, int lane) {
#pragma unroll 8
    for (int j = 0; j < 32; ++j) tile[j * LD + lane] = (j == lane) ? ghinv : 0.0;
    __syncwarp();
    for (int e = lane; e < NJ; e += 32) {
      int r = cbm4_dev::JAC_ROW[e], c = cbm4_dev::JAC_COL[e];
      double v = -jv[e];
      // e(j,j) = -fjac(j,j) + hgamma   (LS_Solver.F90:36-40)
      tile[c * LD + r] = (r == c) ? __dadd_rn(v, ghinv) : v;
    }
    __syncwarp();
  }
  // Hessian(T) (:1190-1853): HESS depends on the rate coefficients only (bimolecular mechanism)
  __device__ static void hess(double* ws, double* hs, int lane) {
    static_assert(cbm4_dev::HES_L2_NG == 1, "generator emitted a different number of HESS groups");
    eval_l1<cbm4_dev::HES_L1_NPAD>(cbm4_dev::HES_L1_TAB, ws, ws + OFF_P, lane);
    eval_terms<cbm4_dev::HES_L20_NTPAD>(cbm4_dev::HES_L20_TERM, ws + OFF_P, ws + OFF_T, lane);
    sum_multi<cbm4_dev::HES_L20_NITER, false>(cbm4_dev::HES_L20_SUM, cbm4_dev::HES_L20_OUT, ws + OFF_T, hs, lane);
  }
  // M'[e] = sum_t HESS[h]*K[c]  +  hg * dJdt[e]   on pattern(M) U time-dependent J entries.
  // (CBM4_HessTR_Vec regrouped: HTU = M^T U1;  non-autonomous term ADJ :1377-1403 folded per stage)
  template <int NTPAD>
  __device__ static void mat_terms(const ushort2* __restrict__ tab, const double* hs, const double* kb, double* T, int lane) {
#pragma unroll
    for (int b = 0; b < NTPAD; b += 32) {
      const ushort2 e = tab[b + lane];
      T[b + lane] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(hs) + e.x) *
                    *reinterpret_cast<const double*>(reinterpret_cast<const char*>(kb) + e.y);
    }
    if (lane == 0) T[NTPAD] = 0.0;
    __syncwarp();
  }
  __device__ static void build_M(double* ws, const double* hs, const double* kb, const double* djdt, double hg, double* mv,
                                 int lane) {
    static_assert(cbm4_dev::MAT_L2_NG == 2, "generator emitted a different number of M groups");
    for (int e = lane; e < NM; e += 32) {
      int jt = cbm4_dev::MAT_JT[e];
      mv[e] = (jt >= 0) ? hg * djdt[jt] : 0.0;
    }
    __syncwarp();
    mat_terms<cbm4_dev::MAT_L20_NTPAD>(cbm4_dev::MAT_L20_TERM, hs, kb, ws + OFF_T, lane);
    sum_multi<cbm4_dev::MAT_L20_NITER, true>(cbm4_dev::MAT_L20_SUM, cbm4_dev::MAT_L20_OUT, ws + OFF_T, mv, lane);
    mat_terms<cbm4_dev::MAT_L21_NTPAD>(cbm4_dev::MAT_L21_TERM, hs, kb, ws + OFF_T, lane);
    sum_multi<cbm4_dev::MAT_L21_NITER, true>(cbm4_dev::MAT_L21_SUM, cbm4_dev::MAT_L21_OUT, ws + OFF_T, mv, lane);
  }
  // ap[p] = ARP_p(Y_i)  (dFun_dRcoeff :1135-1175)  +  aj_p(Y0, K_i) (dJac_dRcoeff :2330-2380)
  // yb must hold Y_i when arp is evaluated and Y0 when aj is evaluated -> two calls
  __device__ static double arp(const double* ws, int lane) {
    const double* yb = ws + OFF_YB;
    cbm4_dev::L1 e = cbm4_dev::ARP_TAB[lane];
    double c = e.lit;
    c = c * yb[e.o1]; c = c * yb[e.o2]; c = c * yb[e.o3];
    return (lane < NP) ? c : 0.0;
  }
  __device__ static double aj(const double* ws, const double* kb, int lane) {
    const double* yb = ws + OFF_YB;
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < cbm4_dev::AJ_NT; ++t) {
      cbm4_dev::L1 e = cbm4_dev::AJ_TAB[t * 32 + lane];
      double c = e.lit;
      c = c * yb[e.o1]; c = c * yb[e.o2]; c = c * yb[e.o3];
      a = a + c * kb[cbm4_dev::AJ_COL[t * 32 + lane]];
    }
    return (lane < NP) ? a : 0.0;
  }
  __device__ static int jac_time_entry(int n) { return cbm4_dev::JAC_TIME[n]; }
  template <class F>
  __device__ __forceinline__ static void jt_mt_rows(const double* Jv, const double* Mv, const double (&x)[32], F&& f) {
    cbm4_dev::jt_mt_rows(Jv, Mv, x, f);
  }
  __device__ __forceinline__ static void jt_vec(const double* Jv, const double (&x)[32], double (&v)[32]) {
    cbm4_dev::jt_vec(Jv, x, v);
  }
  template <class F>
  __device__ __forceinline__ static void mt_rows(const double* Mv, const double (&x)[32], F&& f) { cbm4_dev::mt_rows(Mv, x, f); }
  __device__ __forceinline__ static void jt_vec1(const double* Jv, const double (&x)[32], double (&v)[32]) { cbm4_dev::jt_vec1(Jv, x, v); }
  template <class F>
  __device__ __forceinline__ static void stoich_cols(const double (&x)[32], F&& f) { cbm4_dev::stoich_cols(x, f); }
  // adjinit (dr.F90:101-117): Lambda = I, Mu = 0
  __device__ static double adjinit_lambda(int i, int m, double /*y_i*/) { return (i == m) ? 1.0 : 0.0; }
  __device__ static double adjinit_mu(int /*p*/, int /*m*/, const double* /*yb*/) { return 0.0; }
};

}  // namespace fatode
