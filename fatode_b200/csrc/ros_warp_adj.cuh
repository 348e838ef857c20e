// Discrete Rosenbrock adjoint, backward sweep — one cell per warp, lane m = adjoint column m.
//
// Replaces ros_DadjInt (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:1177-1441; Lambda-only twin
// :3508-3700) together with its LSS_* calls (ADJ/ROS_ADJ/LS_Solver.F90: lapack_decomp :31-42,
// DGETRS('t') :50, lss_mul_jactr :378-381, lss_jac_time_derivative :355-360, lss_mul_djdttr
// :388-391) and the model callbacks HESSTR_VEC / JACP / HESSTR_VEC_F_PY it drives.
//
// Data flow per popped tape entry (T, H, Ystage, K):
//   cooperative part (lane = species):  J(T,Y0) -> E -> LU (registers, same code as the forward
//     sweep) -> row-major L\U + 1/diag(U) + pivot directory in shared memory; Hessian(T);
//     dJ/dt by the reference's forward difference; per stage: J(T_i,Y_i) values, the sparse matrix
//     M'_i = Hess x K_i + h*gamma_i*dJ/dt, and a_p = dF/dp reactant products + dJ/dp x K_i.
//   private part (lane = adjoint column): x = m_i*lambda + W_i in 32 registers, U^T/L^T
//     substitutions with broadcast shared-memory reads of L\U, static sparse J_i^T x, M'^T x and
//     stoichiometric column sums, running sums W_j += a_ij*V_i + c_ij/h*U_i for the earlier stages.
// The columns never couple, so there is no cross-lane traffic in the private part.
//
// Algebraic regrouping vs the reference (results agree to rounding, not bitwise):
//   * U_j, V_j are not stored; their contributions to earlier stages are accumulated in W_j.
//   * sum_i gamma_i U_i times dJ/dt^T (:1392-1403) is folded into M'_i per stage.
//   * HESSTR_VEC(U,K_i) is applied as the sparse matrix (Hess x K_i)^T U.
//   * d(f_p)/dt is identically zero for models whose JACP does not depend on t (cbm4).
#pragma once
#include "fatode_common.cuh"
#include "ros_warp_fwd.cuh"

namespace fatode {

template <class Model>
struct AdjSmem {
  // offsets in doubles inside one warp's slice
  static constexpr int OFF_MODEL = 0;
  static constexpr int OFF_HS = Model::OFF_OUT;                       // HESS aliases the model's L2 scratch
  static constexpr int OFF_JV = Model::WS_DOUBLES;                    // [NJ]
  static constexpr int OFF_MV = OFF_JV + ((Model::NJ + 1) / 2) * 2;   // [NM]
  static constexpr int OFF_DJ = OFF_MV + ((Model::NM + 1) / 2) * 2;   // [32]
  static constexpr int OFF_KB = OFF_DJ + 32;
  static constexpr int OFF_AP = OFF_KB + 32;
  static constexpr int OFF_DINV = OFF_AP + 32;
  static constexpr int OFF_WHO = OFF_DINV + 32;                       // 32 ints
  static constexpr int OFF_LU = OFF_WHO + 16;                         // [32][34] row-major, logical rows
  static constexpr int OFF_LAM = OFF_LU + 32 * 34;                    // [i][m]
  static constexpr int OFF_MU = OFF_LAM + 1024;                       // [p][m]
  static constexpr int OFF_W = OFF_MU + Model::NP * 32;               // [(S-1)][r][m]
  // at least 3 slabs: between steps W doubles as the LU tile ([32][33]) and the dJ/dt scratch (at +2048)
  __host__ __device__ static constexpr int doubles(int S) { return OFF_W + (S - 1 > 3 ? S - 1 : 3) * 1024; }
};

// x <- (L U)^-T x on the row-major factors; 1056 DFMA, all L\U reads are warp-uniform
__device__ __forceinline__ void solve_transposed(const double* __restrict__ LUs, const double* __restrict__ dinv,
                                                 double (&x)[32]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {                 // U^T w = b   (DTRSM Left Upper Trans NonUnit, axpy form)
    double wk = x[k] * dinv[k];
    x[k] = wk;
    const double* row = LUs + k * 34;
#pragma unroll
    for (int ip = 0; ip < 16; ++ip) {
      const int i0 = 2 * ip, i1 = i0 + 1;
      if (i1 > k) {
        double2 t = *reinterpret_cast<const double2*>(row + i0);
        if (i0 > k) x[i0] = fma(-t.x, wk, x[i0]);
        x[i1] = fma(-t.y, wk, x[i1]);
      }
    }
  }
#pragma unroll
  for (int k = 31; k >= 1; --k) {                // L^T z = w   (DTRSM Left Lower Trans Unit, axpy form)
    double zk = x[k];
    const double* row = LUs + k * 34;
#pragma unroll
    for (int ip = 0; ip < 16; ++ip) {
      const int i0 = 2 * ip, i1 = i0 + 1;
      if (i0 < k) {
        double2 t = *reinterpret_cast<const double2*>(row + i0);
        x[i0] = fma(-t.x, zk, x[i0]);
        if (i1 < k) x[i1] = fma(-t.y, zk, x[i1]);
      }
    }
  }
}

// Backward sweep for one cell.  smem = this warp's AdjSmem slice.  tape_base = first entry of the cell.
template <class Model>
__device__ void ros_adjoint_warp(const RosParams& rp, const typename Model::Const& mc, double* smem,
                                 const double* __restrict__ tape_base, int nsteps, int nadj, bool with_mu,
                                 double y_final, int lane) {
  typedef AdjSmem<Model> L;
  const int S = rp.S;
  const int NE = tape_entry_doubles(S, 32);
  double* ws = smem + L::OFF_MODEL;
  double* HS = smem + L::OFF_HS;
  double* JV = smem + L::OFF_JV;
  double* MV = smem + L::OFF_MV;
  double* DJ = smem + L::OFF_DJ;
  double* KB = smem + L::OFF_KB;
  double* AP = smem + L::OFF_AP;
  double* DINV = smem + L::OFF_DINV;
  int* WHO = reinterpret_cast<int*>(smem + L::OFF_WHO);
  double* LUs = smem + L::OFF_LU;
  double* LAM = smem + L::OFF_LAM;
  double* MU = smem + L::OFF_MU;
  double* W = smem + L::OFF_W;
  const int Direction = (rp.tend >= rp.tstart) ? 1 : -1;
  const double dDir = (double)Direction;
  const double DeltaMin = 1.0e-5;   // ros_DadjInt :1212

  Model::init(ws, mc, lane);
  // ADJINIT (dr.F90:101-117): lane = column m
  Model::set_y(ws, y_final, lane);
#pragma unroll 4
  for (int i = 0; i < 32; ++i) LAM[i * 32 + lane] = Model::adjinit_lambda(i, lane, ws[Model::OFF_YB + i]);
  for (int p = 0; p < Model::NP; ++p) MU[p * 32 + lane] = Model::adjinit_mu(p, lane, ws + Model::OFF_YB);
  __syncwarp();

  for (int step = nsteps - 1; step >= 0; --step) {
    const double* ent = tape_base + (long)step * NE;
    const double T = ent[0], H = ent[1];
    const double* Yst = ent + 4;
    const double* Kst = ent + 4 + S * 32;
    const double y0 = Yst[lane];
    // ---- cooperative: J(T,Y0), LU of E = 1/(h gamma_1) - J   (:1261-1267)
    Model::set_time(ws, T, lane);
    Model::set_y(ws, y0, lane);
    Model::jac(ws, JV, lane);
    {
      const double ghinv = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[0]));
      double* tile = W;   // [32][LDA]: W is free between steps
      Model::template scatter_E<LDA>(JV, ghinv, tile, lane);
      int pos;
      warp_lu_factor(tile, WHO, pos, lane);       // same code as the forward sweep -> identical factors
      // repack for the private phase: logical rows, row-major, ld 34 (lane = column j)
      for (int r = 0; r < 32; ++r) LUs[r * 34 + lane] = tile[lane * LDA + WHO[r]];
      __syncwarp();
      DINV[lane] = 1.0 / LUs[lane * 34 + lane];
    }
    Model::hess(ws, HS, lane);                   // Hessian(T): rate coefficients only
    if (!rp.autonomous) {                        // dJ/dt (:1379-1382) on the time-dependent entries
      const double Delta = sqrt(rp.roundoff) * fmax(DeltaMin, fabs(T));
      double* Jd = W + 2048;                       // scratch (S >= 4) / beyond the tile
      Model::set_time(ws, T + Delta, lane);
      Model::jac(ws, Jd, lane);
      if (lane < Model::NJT) {
        int e = Model::jac_time_entry(lane);
        DJ[lane] = (Jd[e] - JV[e]) * (1.0 / Delta);
      }
      __syncwarp();
    } else {
      DJ[lane] = 0.0;
      __syncwarp();
    }
    const bool ident = __all_sync(FATODE_FULL_MASK, WHO[lane] == lane);
    const double inv_dH = 1.0 / (dDir * H);

    double lacc[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) lacc[r] = 0.0;

    for (int is = S - 1; is >= 0; --is) {
      // ---- cooperative: stage Jacobian, M', a_p
      const double Tau = T + rp.Alpha[is] * dDir * H;
      const double yi = Yst[is * 32 + lane];
      const double ki = Kst[is * 32 + lane];
      Model::set_time(ws, Tau, lane);
      Model::set_y(ws, yi, lane);
      Model::jac(ws, JV, lane);                  // J(T_i, Y_i)  (:1311)
      double ap = 0.0;
      if (with_mu) ap = Model::arp(ws, lane);    // JACP(T_i, Y_i) (:1348)
      KB[lane] = ki;
      Model::set_y(ws, y0, lane);
      if (with_mu) { ap += Model::aj(ws, KB, lane); AP[lane] = ap; }   // HESSTR_VEC_F_PY(T,Y0,.,K_i) (:1363)
      const double hg = rp.autonomous ? 0.0 : H * rp.Gamma[is];
      Model::build_M(HS, KB, DJ, hg, MV, lane);  // HESSTR_VEC(T,Y0,.,K_i) (:1330) + non-autonomous term
      // ---- private: lane = adjoint column
      double x[32];
      const double mi = rp.M[is];
      if (is < S - 1) {
        const double* Wi = W + is * 1024;
#pragma unroll
        for (int r = 0; r < 32; ++r) x[r] = fma(mi, LAM[r * 32 + lane], Wi[r * 32 + lane]);
      } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) x[r] = mi * LAM[r * 32 + lane];
      }
      solve_transposed(LUs, DINV, x);            // LSS_Solve(Transp, U_i(m)) (:1301-1304)
      if (!ident) {                              // x = P^T z : natural index who[r] <- logical r
        double* sc = W + ((is < S - 1) ? is : (S > 1 ? S - 2 : 0)) * 1024;
#pragma unroll
        for (int r = 0; r < 32; ++r) sc[WHO[r] * 32 + lane] = x[r];
#pragma unroll
        for (int r = 0; r < 32; ++r) x[r] = sc[r * 32 + lane];
      }
      const bool first = (is == S - 1);
      double ca[5], cc[5];                        // a_{is,j}, c_{is,j}/h for the earlier stages j < is (:1294-1295)
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const bool on = j < is;
        ca[j] = on ? rp.A[is * (is - 1) / 2 + j] : 0.0;
        cc[j] = on ? rp.C[is * (is - 1) / 2 + j] * inv_dH : 0.0;
      }
      Model::jt_mt_rows(JV, MV, x, [&](auto rc, double v, double h) {
        constexpr int r = decltype(rc)::value;
        lacc[r] += v + h;                        // Lambda += V_i + (Hess x K_i)^T U_i (+ h*gamma_i*dJdt^T U_i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {            // U_j seed sums for the earlier stages (:1292-1300)
          if (j < is) {
            double* wj = W + j * 1024 + r * 32 + lane;
            const double t = fma(ca[j], v, cc[j] * x[r]);
            *wj = first ? t : (*wj + t);
          }
        }
      });
      if (with_mu) {
        Model::stoich_cols(x, [&](auto pc, double s) {
          constexpr int p = decltype(pc)::value;
          MU[p * 32 + lane] = fma(AP[p], s, MU[p * 32 + lane]);
        });
      }
      __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) LAM[r * 32 + lane] += lacc[r];
    __syncwarp();
  }
  (void)nadj;
}

}  // namespace fatode
