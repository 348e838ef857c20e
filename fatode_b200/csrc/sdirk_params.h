// Parsed SDIRK options, identical for every system of a batch (passed by value to kernels); shared by the per-thread kernels of
// sdirk_cell_fwd.cuh and the lock-step tangent linear kernel of sdirk_tlm_lock.cu.
#pragma once
#include <cuda_runtime.h>

namespace fatode {

struct SdirkParams {
  int S;
  double Gamma, ELO, C[5], D[5], E[5], Theta[5][5], Alpha[5][5];
  int vector_tol, max_no_steps, newton_maxit, start_newton;
  double roundoff, hmin, hmax, hstart, facmin, facmax, facrej, facsafe, thetamin, newtontol, qmin, qmax;
  double tstart, tend;
};

// tape record of one accepted step written by k_sdirk_fwd_cell in its TLM / ADJ modes (sdirk_cell_fwd.cuh)
constexpr int SD_REC = 196;   // T, H, packed saveNiter (4 bits per stage), pad, Y(32), Z(5 x 32)

// coefficients of the discrete adjoint (ADJ/SDIRK_ADJ: rkA, rkB of the transposed scheme)
struct SdirkAdjCoef { double A[5][5], B[5]; };

__device__ __forceinline__ double sdirk_powi(double x, int m) {   // gfortran's x**m for INTEGER m: libgcc __powidf2
  unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) {
    x = __dmul_rn(x, x);
    if (n % 2) y = __dmul_rn(y, x);
  }
  return m < 0 ? __ddiv_rn(1.0, y) : y;
}

}  // namespace fatode
