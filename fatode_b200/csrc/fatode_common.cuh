// fatode_b200 — shared definitions for the device kernels and the host-side option parser.
// All arithmetic is IEEE FP64.  The translation units are compiled with -fmad=false so that
// a*b+c is never contracted behind our back: the forward sweep spells its operations with
// __dmul_rn/__dadd_rn in the reference's order (bit-faithful up to cos/pow), the adjoint
// sweep spells fma() explicitly where DFMA throughput matters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fatode {

// ISTATUS / RSTATUS slots (0-based views of the reference's Nfun..Nsng, Ntexit..Nhnew;
// ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:26-29)
enum { Nfun = 0, Njac = 1, Nstp = 2, Nacc = 3, Nrej = 4, Ndec = 5, Nsol = 6, Nsng = 7 };
enum { Ntexit = 0, Nhexit = 1, Nhnew = 2 };

enum RosFamily { FAM_FWD = 0, FAM_ADJ = 1, FAM_TLM = 2 };

// Parsed Rosenbrock options, identical for every cell of a batch (passed by value to kernels).
struct RosParams {
  int S;
  double A[15], C[15], M[6], E[6], Alpha[6], Gamma[6], ELO;
  int NewF[6];
  int autonomous, vector_tol, max_no_steps, family;
  double roundoff, hmin, hmax, hstart, facmin, facmax, facrej, facsafe;
  double deltamin_start, deltamin_fd, tenth;
  double tstart, tend;
};

// Tape entry (one accepted step): [T, H, pad, pad, Ystage(S*N), K(S*N)]   (ros_DPush, ADJ :735-755)
__host__ __device__ inline int tape_entry_doubles(int S, int N) { return 4 + 2 * S * N; }

#define FATODE_FULL_MASK 0xffffffffu

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FATODE_FULL_MASK, v, src); }

}  // namespace fatode
