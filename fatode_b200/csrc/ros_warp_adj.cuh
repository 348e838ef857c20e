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

}  // namespace fatode
