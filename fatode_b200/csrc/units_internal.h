// Host-side declarations shared by the translation units of libfatode_b200.so (not part of the C ABI).
// Each family added after round 1 lives in its own .cu (compiled in parallel by fatode_b200/build.py); fatode_capi.cu keeps the
// C entry points, argument checks and host<->device staging and calls the unit's device-level driver declared here.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "fatode_common.cuh"
#include "model_cbm4.cuh"
#include "sdirk_params.h"
#include "erk_swe.cuh"

namespace fatode {

// CUDA call guard for the units: the message goes to `err`, the return value is the C ABI's error code
#define UNIT_TRY(x)                                                                                           \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess) {                                                                                  \
      err = std::string(#x) + ": " + cudaGetErrorString(e_);                                                  \
      return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? -102 /*FATODE_ENODEVICE*/ : -103 /*FATODE_ECUDA*/; \
    }                                                                                                         \
  } while (0)

struct UnitTiming { double ms = 0.0; long launches = 0; };

// ros_tlm_lock.cu — ros_TLM_Int (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:462-707) with state and columns in lock-step and
// general partial pivoting.  ids == nullptr: cells 0..ncells-1; otherwise the listed cells (device array).
// trunc: TLM truncation-error control (ICNTRL(12) != 0), d_atol_tlm / d_rtol_tlm [ntlm][32] device arrays then required.
int ros_tlm_lock_device(const RosParams& rp, const Cbm4Const& mc, long ncells, const int* d_ids, int ntlm, bool trunc,
                        double* d_y, double* d_ytlm, const double* d_atol, const double* d_rtol, const double* d_atol_tlm,
                        const double* d_rtol_tlm, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st,
                        UnitTiming& tm, std::string& err);

// sdirk_tlm_lock.cu — SDIRK_IntegratorTLM (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:474-911) with state and columns in
// lock-step: direct TLM solve (ICNTRL(7) = 1), TLM Newton control (ICNTRL(9)), TLM truncation-error control (ICNTRL(12)).
int sdirk_tlm_lock_device(const SdirkParams& sp, const Cbm4Const& mc, long ncells, const int* d_ids, int ntlm, bool direct,
                          bool newton_est, bool trunc, double* d_y, double* d_ytlm, const double* d_atol, const double* d_rtol,
                          const double* d_atol_tlm, const double* d_rtol_tlm, int* d_istatus, double* d_rstatus, int* d_ierr,
                          cudaStream_t st, UnitTiming& tm, std::string& err);

// sdirk_adj_newton.cu — backward sweep of ADJ/SDIRK_ADJ with simplified-Newton adjoint stages (ICNTRL(7) /= 1,
// ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:786-995), one warp per system over the forward sweep's tape records; all device arrays
int sdirk_adj_newton_device(const SdirkParams& sp, const Cbm4Const& mc, const SdirkAdjCoef& co, int nsys, const long long* d_rec_off,
                            const int* d_slot, const int* d_cell, const double* d_tape, int tape_cap, int nadj, int with_mu, int np,
                            double* d_lambda, double* d_mu, const double* d_atol_adj, const double* d_rtol_adj, int* d_istatus,
                            int* d_ierr, unsigned long long* d_queue, cudaStream_t st, UnitTiming& tm, std::string& err);

// erk_adj_swe.cu — INTEGRATE_ADJ of ADJ/ERK_ADJ (no parameters) on the shallow-water ensemble: forward sweep with checkpoints,
// ADJINIT (Lambda(k,k) = 1) and ERK_DadjInt with the matrix-free JAC^T of the example; one member per CTA.  Device arrays.
// tape_slot: checkpoints per CTA in the first pass (0: 128); members that need more are integrated again with Max_no_steps.
int erk_adj_swe_device(const ErkParams& ep, long nmembers, int nadj, double* d_y, double* d_lambda, const double* d_atol,
                       const double* d_rtol, int* d_istatus, double* d_rstatus, int* d_ierr, int tape_slot, size_t budget,
                       cudaStream_t st, UnitTiming& tm, long& retried, std::string& err);

// ros_small_adj.cu — INTEGRATE_ADJ (ADJ/ROS_ADJ, discrete adjoint) for small_strato / roberts, one system per thread; device arrays
int ros_small_cadj_epilogue_device(int model, const RosParams& rp, const double* fix, long ncells, int nadj, bool with_mu, const double* d_y,
                                   double* d_lambda, double* d_mu, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st);
int ros_small_adj_device(int model, const RosParams& rp, const double* fix, long ncells, int nadj, bool with_mu, double* d_y,
                         double* d_lambda, double* d_mu, const double* d_atol, const double* d_rtol, int* d_istatus,
                         double* d_rstatus, int* d_ierr, size_t tape_budget, cudaStream_t st, UnitTiming& tm, std::string& err);

// fatode_comm.cu — the library's NCCL communicator (optional): sum of a device buffer over the ranks, stream-ordered
bool comm_active();
int comm_allreduce_device(double* d_buf, size_t count, cudaStream_t st, std::string& err);

}  // namespace fatode
