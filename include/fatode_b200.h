/* fatode_b200 — C ABI of the B200-native batched stiff-integration engine.
 *
 * Drop-in boundary for FATODE's data-parallel hot path.  Each entry point is the
 * batched form of one Fortran90 interface of the reference; `ncells` independent systems are
 * integrated in one call (ncells = 1 reproduces the reference call).  What a Fortran binding
 * (fatode_b200/fortran/fatode_b200_mod.F90, ISO_C_BINDING) passes here is exactly what the
 * reference's INTEGRATE* routines receive, except that the EXTERNAL callbacks FUN/JAC/HESSTR_VEC/
 * JACP/HESSTR_VEC_F_PY/ADJINIT/HESS_VEC are replaced by a registered model id: the shipped example
 * models are compiled as device functions.
 *
 * Array conventions (Fortran order, 0-based description):
 *   y       [ncells][nvar]              VAR(1:N) of cell c at y + c*nvar                 (in/out)
 *   lambda  [ncells][nadj][nvar]        Lambda(1:N,1:NADJ) of cell c, column-major        (out)
 *   mu      [ncells][nadj][np]          Mu(1:NP,1:NADJ) of cell c, column-major           (out)
 *   atol, rtol [nvar]                   shared by all cells (ATOL(1:N), RTOL(1:N))
 *   icntrl_u[20], rcntrl_u[20]          ICNTRL_U / RCNTRL_U, may be NULL (all defaults)
 *   istatus [ncells][20], rstatus [ncells][20], ierr [ncells]   per-cell ISTATUS_U / RSTATUS_U / IERR
 * `mem` says where those arrays live: FATODE_MEM_HOST (the library stages them through the GPU,
 * host<->device copies included) or FATODE_MEM_DEVICE (device pointers, results stay resident;
 * icntrl_u/rcntrl_u/atol/rtol are always host pointers).
 *
 * Return value: 0 on success of the call itself (per-cell outcomes are in ierr[]: 1 = success,
 * negative = the reference's error codes), or a negative FATODE_E* code; fatode_last_error()
 * returns a message.  There is no CPU fallback: without a usable CUDA device every compute entry
 * point fails with FATODE_ENODEVICE.
 */
#ifndef FATODE_B200_H
#define FATODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* model registry (replaces the FUN/JAC/... callback arguments) */
enum {
  FATODE_MODEL_ROBERTS = 1,      /* EXAMPLES/roberts/ROBERTS_ROS_ADJ/roberts_ros_adj_dr.F90 : N=3,  NP=3  */
  FATODE_MODEL_SMALL_STRATO = 2, /* EXAMPLES/small_strato/small_ros{,_adj}/User_Parameters.F90 : N=5, NP=0 */
  FATODE_MODEL_CBM4 = 3,         /* EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90          : N=32, NP=29 */
  FATODE_MODEL_SWE = 4           /* EXAMPLES/swe/SWE_ERK/swe2D_upwind.F90 (40 x 40 grid)    : N=4800, NP=0; ERK entry points only */
};

enum { FATODE_MEM_HOST = 0, FATODE_MEM_DEVICE = 1 };

enum {
  FATODE_OK = 0,
  FATODE_EINVAL = -101,    /* bad argument (unknown model, dimension mismatch, NULL required pointer) */
  FATODE_ENODEVICE = -102, /* no CUDA device / driver */
  FATODE_ECUDA = -103,     /* CUDA runtime error (message in fatode_last_error) */
  FATODE_ENOMEM = -104,    /* trajectory tape does not fit in device memory even for one cell */
  FATODE_EUNSUPPORTED = -105 /* option of the reference outside this engine's scope */
};

const char* fatode_last_error(void);
const char* fatode_version(void);

/* dimensions of a registered model: nvar, np (number of parameters for Mu) */
int fatode_model_dims(int model, int* nvar, int* np);
/* initial state of the reference's example driver for this model (host array [nvar]) */
int fatode_model_initial_state(int model, double* y);

/* Optional per-call model constants; NULL = the example driver's values.
 *  CBM4: params[0] = TEMP (298), params[1] = FIX(1) (1.25e8*2.55e10), params[2] = emission (2e3)
 *        (cbm4_ros_adj_dr.F90:96-97,202,268).  Other models take none. */

/* INTEGRATE of FWD/ROS (FWD/ROS/ROS_f90_Integrator.F90:32):
 *   INTEGRATE(TIN,TOUT,N,NNZERO,VAR,RTOL,ATOL,FUN,JAC,ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,IERR_U)
 * ICNTRL_U/RCNTRL_U entries override the defaults only when > 0 (:67-72). */
int fatode_ros_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                               const double* atol, const double* rtol, const int* icntrl_u,
                               const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                               const double* model_params, int mem, void* stream);

/* INTEGRATE_ADJ of ADJ/ROS_ADJ (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:34):
 *   INTEGRATE_ADJ(NVAR,NP,NADJ,NNZERO,Y,Lambda,Mu,TIN,TOUT,ATOL_adj,RTOL_adj,ATOL,RTOL,FUN,JAC,ADJINIT,
 *                 HESSTR_VEC,JACP,DRDY,DRDP,HESSTR_VEC_F_PY,HESSTR_VEC_R_PY,HESSTR_VEC_R,
 *                 ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,Q,QFUN)
 * Discrete adjoint (ICNTRL(7) = 0 or 2) or forward only (1).  ICNTRL_U/RCNTRL_U override when >= 0
 * (:91-96).  mu == NULL or np == 0 selects the Lambda-only mode (RosenbrockADJ2).  ADJINIT is the
 * model's registered routine.  ATOL_adj/RTOL_adj are accepted for signature parity and unused (the
 * reference reads them only in the continuous adjoint, which never gets that far).  ICNTRL(7) = 3 (continuous adjoint) behaves
 * as the reference does: INTEGRATE_ADJ enters ros_CadjInt backwards in time with a negative H, whose first test
 * `H <= Roundoff` (:1517) returns IERR -7 at once — per cell: the forward solution, ADJINIT's Lambda / Mu, the forward
 * counters plus the last checkpoint's FUN / JAC, RSTATUS(Nhexit) = 0, IERR -7 (-6 if the forward sweep used up Max_no_steps).
 * ICNTRL(7) = 4 (simplified continuous adjoint) returns FATODE_EUNSUPPORTED: ros_SimpleCadjInt multiplies its first stage by J
 * instead of J^T (:1741) and overflows within tens of steps (restated and shown in oracle/ros.hpp, tests/test_ros_cadj.py).
 * Quadrature arguments (Q, QFUN, DRDY, DRDP, HESSTR_VEC_R*) are not accepted: the reference's quadrature weights ros_W come
 * from a loop that never executes (:1241-1248).
 * Models: cbm4 (nadj <= 32; the headline path), small_strato (RosenbrockADJ2, EXAMPLES/small_strato/small_ros_adj) and
 * roberts (NP = 3, EXAMPLES/roberts/ROBERTS_ROS_ADJ without its quadrature terms), nadj <= 8 for the two small models.
 * mu_sum (may be NULL): [nadj][np] sum over all cells of this call with ierr == 1, reduced in a
 * fixed order on the device (host pointer unless mem == FATODE_MEM_DEVICE); with a communicator
 * (fatode_comm_init) it is additionally summed over the cells of all ranks. */
int fatode_ros_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y,
                                   double* lambda, double* mu, double tin, double tout,
                                   const double* atol_adj, const double* rtol_adj, const double* atol,
                                   const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                   int* istatus, double* rstatus, int* ierr, double* mu_sum,
                                   const double* model_params, int mem, void* stream);

/* INTEGRATE_TLM of TLM/ROS_TLM (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:35):
 *   INTEGRATE_TLM(N,NTLM,NNZERO,Y,Y_tlm,TIN,TOUT,ATOL_tlm,RTOL_tlm,ATOL,RTOL,FUN,JAC,HESS_VEC,
 *                 ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U)
 * y_tlm [ncells][ntlm][nvar] = Y_tlm(1:N,1:NTLM) of cell c, column-major (in: sensitivities at TIN, out: at TOUT).
 * ICNTRL_U/RCNTRL_U override the presets (ICNTRL(2)=1, (3)=5, (12)=1) when >= 0 (:77-82).  Both error-control modes are
 * built: ICNTRL(12) = 0 (forward error control: state sweep + column sweep over the tape) and ICNTRL(12) != 0 — the PRESET
 * when icntrl_u == NULL — where the columns' error norms (ros_ErrorNorm_tlm, scaled with ATOL_tlm/RTOL_tlm [ntlm][nvar], then
 * required) enter the accept/reject decision: state and columns advance in lock-step in one kernel whose column arithmetic
 * follows the reference operation by operation (csrc/ros_tlm_lock.cu).  RCNTRL(3) = STEPMIN — a SAVEd variable of the
 * reference that leaks the last step size of one call into the next one as Hstart — is 0 for every cell (every system of a
 * batch is a first call); pass RCNTRL_U(3) = RSTATUS(Nhexit) of the previous call to reproduce the carry-over.
 * ntlm <= 32, cbm4 only; HESS_VEC is the model's registered routine.  A cell whose state sweep ends with an error code
 * returns Y and Y_tlm as of its last accepted step, as the reference.  Systems whose factorisations leave the static pivot
 * pattern of the per-thread kernels (almost all at rtol >= 1e-3, none at rtol <= 1e-5) are integrated by the lock-step kernel
 * with netlib's partial pivoting (fatode_stats.cells_general counts them). */
int fatode_ros_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm,
                                   double tin, double tout, const double* atol_tlm, const double* rtol_tlm,
                                   const double* atol, const double* rtol, const int* icntrl_u,
                                   const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                   const double* model_params, int mem, void* stream);

/* INTEGRATE of FWD/SDIRK (FWD/SDIRK/SDIRK_f90_Integrator.F90:25):
 *   INTEGRATE(TIN,TOUT,N,NNZERO,VAR,RTOL,ATOL,FUN,JAC,ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,Ierr_U)
 * Singly diagonally implicit Runge-Kutta methods (ICNTRL(3) = 1..5: Sdirk2a, 2b, 3a, 4a (default), 4b) with simplified
 * Newton iterations; ICNTRL_U/RCNTRL_U override the defaults only when > 0 (:57-64); RCNTRL(8..11) = ThetaMin,
 * NewtonTol, Qmin, Qmax, ICNTRL(5) = NewtonMaxit, ICNTRL(6) = zero starting values for Newton.  Models: cbm4,
 * small_strato, roberts.  Per-cell IERR: 1, or the reference's -1..-8.  cbm4 systems whose factorisations leave the static
 * pivot pattern are integrated a second time with general partial pivoting (fatode_stats.cells_general counts them). */
int fatode_sdirk_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                                 const double* atol, const double* rtol, const int* icntrl_u,
                                 const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                 const double* model_params, int mem, void* stream);

/* INTEGRATE_TLM of TLM/RK_TLM (TLM/RK_TLM/RK_TLM_f90_Integrator.F90:33):
 *   INTEGRATE_TLM(N,NTLM,NNZERO,Y,Y_TLM,TIN,TOUT,ATOL_TLM,RTOL_TLM,ATOL,RTOL,FUN,JAC,ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,IERR_U)
 * Tangent linear model of the 3-stage implicit Runge-Kutta methods (ICNTRL(3) = 1 Radau-2A, 0/2 Lobatto-3C, 3 Gauss, 4 Radau-1A)
 * in the reference's preset mode: ICNTRL(5) = 8, (7) = 1 DIRECT TLM solve (one factorisation of the 3N x 3N system
 * I - H (A x J_b) per accepted step and one solve per column, :761-793 — the modified-Newton TLM, ICNTRL(7) = 0, cannot be
 * selected through INTEGRATE_TLM because of the `> 0` override rule), (10) = 1 SDIRK error estimator, (11) = 1 classic
 * controller, (12) = 0 state-only truncation error; ICNTRL(12) != 0 returns FATODE_EUNSUPPORTED.  y_tlm [ncells][ntlm][nvar],
 * ntlm <= 32, cbm4 only.  Per-cell IERR: 1, the reference's -1..-13, or -12 when the big matrix is singular (the reference
 * STOPs there).  A cell whose state sweep fails returns Y and Y_tlm as of its last accepted step, as the reference.  Systems
 * whose step factorisations leave the static pivot pattern take the general-pivoting second pass of the RK family.  State,
 * counters and Y_tlm are bit-identical to the reference's operation order (oracle/rk_tlm.hpp). */
int fatode_rk_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm,
                                  double tin, double tout, const double* atol_tlm, const double* rtol_tlm,
                                  const double* atol, const double* rtol, const int* icntrl_u,
                                  const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                  const double* model_params, int mem, void* stream);

/* INTEGRATE_TLM of TLM/SDIRK_TLM (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:31):
 *   INTEGRATE_TLM(N,NTLM,NNZERO,Y,Y_TLM,TIN,TOUT,ATOL_TLM,RTOL_TLM,ATOL,RTOL,FUN,JAC,
 *                 ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,IERR_U)
 * y_tlm [ncells][ntlm][nvar] as for ROS_TLM.  Presets: ICNTRL(5) = 8 Newton iterations, everything else 0; user values
 * override when > 0 (:60-77); an ICNTRL(3) outside 0..5 selects Sdirk2a in this family (:262-275).  Every mode of the
 * reference is built.  Preset mode — modified-Newton TLM (ICNTRL(7) = 0) with the iteration count of the state stage, columns
 * outside the step control (ICNTRL(9) = ICNTRL(12) = 0): state sweep + preparation + column sweep over a tape; systems whose
 * factorisations leave the static pivot pattern (all of them at rtol >= 1e-3) are integrated a second time with general
 * partial pivoting (fatode_stats.cells_general counts them).  ICNTRL(7) = 1 (direct TLM solve: a second factorisation per
 * stage, :659-705), ICNTRL(9) /= 0 (the columns' Newton iterations are convergence-controlled with ATOL_TLM / RTOL_TLM and
 * can reject the attempt, :735-800) and ICNTRL(12) /= 0 (the columns' error estimates enter the step control, :825-837), in
 * any combination: one lock-step kernel advances the state and the columns of a system together with general pivoting
 * (cells_general = ncells).  atol_tlm / rtol_tlm [ntlm][nvar] (the Fortran (N, NTLM) arrays, one set per call) are read
 * when ICNTRL(9) or ICNTRL(12) is set and may be NULL otherwise (FATODE_EINVAL if they are needed and NULL).
 * ntlm <= 32, cbm4 only.  Per-cell IERR: 1 or the reference's -1..-8.  State, ISTATUS, RSTATUS and Y_tlm are bit-identical
 * to the reference's operation order in every mode.  A cell whose integration fails returns Y and Y_tlm as of its last
 * accepted step, as the reference. */
int fatode_sdirk_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm,
                                     double tin, double tout, const double* atol_tlm, const double* rtol_tlm,
                                     const double* atol, const double* rtol, const int* icntrl_u,
                                     const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                     const double* model_params, int mem, void* stream);

/* INTEGRATE_ADJ of ADJ/SDIRK_ADJ (ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:29):
 *   INTEGRATE_ADJ(NVAR,NP,NADJ,NNZERO,Y,Lambda,Mu,TIN,TOUT,ATOL_adj,RTOL_adj,ATOL,RTOL,FUN,JAC,ADJINIT,JACP,DRDY,DRDP,
 *                 ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,Ierr_U,Q,QFUN)
 * Discrete adjoint of the five SDIRK methods.  Presets ICNTRL(5) = 8, ICNTRL(7) = 1 (direct adjoint solve: one factorisation per
 * stage), user values override when > 0; Max_no_steps defaults to 10000 and Newton starting values are always interpolated in
 * this family.  Built: both adjoint solution modes, NADJ <= 32, Mu for the model's NP parameters, ADJINIT of the drivers
 * (Lambda = I, Mu = 0), cbm4 only; the quadrature terms return FATODE_EUNSUPPORTED.  ICNTRL(7) /= 1 selects the
 * simplified-Newton adjoint stages (:890-951): one factorisation per step, the columns iterated until they converge in their own
 * scale 1 / (ATOL_adj + RTOL_adj |Lambda|) — atol_adj / rtol_adj [nadj][nvar] (the Fortran (N, NADJ) arrays, one set per call) are
 * then required (FATODE_EINVAL if NULL).  A column that does not converge takes the reference's fall-back, LSS_Decomp(HGamma)
 * with a HGamma this mode never assigns upstream (:804, :855, :941): reproduced with HGamma = 0, the value a zero-initialised stack
 * gives — the results of such columns, and of the columns after them (they iterate with the new factors), are as meaningless as
 * the reference's, and fall-backs are frequent; use the preset direct mode for sensitivities.
 * Reference behaviour kept: a system whose forward sweep fails still gets ADJINIT and the backward sweep over its accepted
 * steps, and its IERR is the backward sweep's (1) — RSTATUS(1) < TOUT tells (:519-531).  Systems that leave the static pivot
 * pattern take a second pass with general partial pivoting, as for SDIRK_TLM.  Arguments as fatode_ros_integrate_adj_batch. */
int fatode_sdirk_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda,
                                     double* mu, double tin, double tout, const double* atol_adj, const double* rtol_adj,
                                     const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                     int* istatus, double* rstatus, int* ierr, double* mu_sum, const double* model_params,
                                     int mem, void* stream);

/* INTEGRATE of FWD/ERK (FWD/ERK/ERK_f90_Integrator.F90:28):
 *   INTEGRATE(TIN,TOUT,NVAR,VAR,RTOL,ATOL,FUN,ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,Ierr_U)   (no NNZERO, no JAC)
 * Explicit Runge-Kutta pairs, ICNTRL(3) = 1 Erk23, 2 Erk3_Heun, 3 Erk43, 0/4 Dopri5 (default), 5 Verner, 6 Dopri853; user
 * values override the defaults when > 0 (:56-61).  Built for the shallow-water ensemble (model FATODE_MODEL_SWE, one member
 * per system, y [ncells][4800] in the reference's grid2vec order).  Per-member IERR: 1 or the reference's -1..-7.
 * ISTATUS: Nfun = S * Nstp, Njac = Ndec = Nsol = Nsng = 0. */
int fatode_erk_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                               const double* atol, const double* rtol, const int* icntrl_u,
                               const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                               const double* model_params, int mem, void* stream);

/* INTEGRATE_ADJ of ADJ/ERK_ADJ (ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:33):
 *   INTEGRATE_ADJ(NVAR,NP,NADJ,NNZ,Y,Lambda,TIN,TOUT,ATOL,RTOL,FUN,JAC,AdjInit,ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,Ierr_U,
 *                 Mu,JACP,DRDY,DRDP,QFUN,Q)
 * Discrete adjoint of the six explicit pairs on the shallow-water ensemble (model FATODE_MODEL_SWE, the family's only example,
 * EXAMPLES/swe/SWE_ERK_ADJ): forward sweep with the checkpoint of every accepted step (ERK_FwdInt :1600-1740), ADJINIT of the
 * driver (Lambda(k,k) = 1, swe2D_erk_adj_dr.F90:31-43) and ERK_DadjInt, U_i = JAC^T (h sum a_ji U_j + h b_i Lambda) per stage with
 * the example's JAC callback (compute_JacF, swe2D_upwind.F90:483-2023) applied matrix-free.  Options as fatode_erk_integrate_batch
 * (ICNTRL(2), (3), (4), RCNTRL(1..7), (10)); the call without parameters (ERKADJ2) is built: np > 0 with mu != NULL and the
 * quadrature terms return FATODE_EUNSUPPORTED (the example has no parameters).  lambda [ncells][nadj][nvar] (column k of the
 * Fortran Lambda(NVAR,NADJ) at lambda[cell][k][:]), nadj <= nvar.  Per-member IERR: 1 (as upstream, also after a forward sweep that
 * stopped with -6 / -7: ADJINIT and the backward sweep still run over the accepted steps, RSTATUS(1) < TOUT tells), or -6 when
 * more than Max_no_steps steps are accepted (the reference program STOPs in ERK_Push, :766-770; Lambda is then left as given).
 * ISTATUS: Nfun = S * attempts, Njac = S * Nacc, Nstp = attempts + Nacc.  State, ISTATUS, RSTATUS bit-identical to the reference's
 * operation order, Lambda within 1e-9 (rounding level observed).  fatode_set_erk_tape_slot: checkpoints a CTA reserves in the first
 * pass (default 128); members with more accepted steps are integrated again with Max_no_steps checkpoints
 * (fatode_stats.cells_deferred counts them). */
int fatode_erk_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda,
                                   double* mu, double tin, double tout, const double* atol, const double* rtol,
                                   const int* icntrl_u, const double* rcntrl_u, int* istatus, double* rstatus,
                                   int* ierr, const double* model_params, int mem, void* stream);
int fatode_set_erk_tape_slot(int checkpoints);

/* INTEGRATE_TLM of TLM/ERK_TLM (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:28):
 *   INTEGRATE_TLM(TIN,TOUT,NVAR,NTLM,VAR,VAR_TLM,ATOL_TLM,RTOL_TLM,ATOL,RTOL,FUN,JAC,
 *                 ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,IERR_U)
 * (the C ABI keeps the argument order of the other TLM entry points).  Built: the preset mode, ICNTRL(5) = 0 — Jacobian-vector
 * products by forward finite differences (FDJACV :1071-1097, increment RCNTRL(12) or sqrt(max(RTOL(1), eps))), one extra
 * FUN per column per stage, Nfun = S * (1 + NTLM) * Nstp — with or without the TLM truncation-error control (ICNTRL(12) /= 0:
 * the columns' error norms, scaled with ATOL_TLM / RTOL_TLM [ntlm][nvar] — host arrays shared by all members — enter the step
 * control, :493-501).  ICNTRL(5) /= 0 selects the dense-Jacobian mode (:460-474): one JAC per stage and
 * K_tlm = MATMUL(FJAC, S_tlm) per column — the example's JAC callback (compute_JacF, EXAMPLES/swe/SWE_ERK_TLM/swe2D_upwind.F90:
 * 483-2007) is compiled as a device function and applied matrix-free (27 structural entries per row); Nfun = S * Nstp = Njac;
 * state bit-identical, Y_tlm within 1e-9 of the reference's operation order (pow(x, 1.5) is CUDA's).
 * y_tlm [ncells][ntlm][4800]. */
int fatode_erk_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm,
                                   double tin, double tout, const double* atol_tlm, const double* rtol_tlm,
                                   const double* atol, const double* rtol, const int* icntrl_u,
                                   const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                   const double* model_params, int mem, void* stream);

/* Initial state of the shallow-water drivers: Gaussian bump of the given amplitude on h = 100, constant velocities
 * (initializeGaussianGrid, EXAMPLES/swe/SWE_ERK/swe2D_upwind.F90:86-126; swe2D_erk_dr.F90:100-103 uses 30, 2, 2). */
int fatode_swe_initial_state(double amplitude, double u0, double v0, double* y);

/* INTEGRATE of FWD/RK (FWD/RK/RK_f90_Integrator.F90:32):
 *   INTEGRATE(TIN,TOUT,N,NNZERO,VAR,RTOL,ATOL,FUN,JAC,ICNTRL_U,RCNTRL_U,ISTATUS_U,RSTATUS_U,IERR_U)
 * Fully implicit 3-stage Runge-Kutta methods with simplified Newton iterations on the transformed (one real + one complex)
 * system: ICNTRL(3) = 1 Radau-2A, 0/2 Lobatto-3C (the default), 3 Gauss, 4 Radau-1A, 5 Lobatto-3A.  INTEGRATE presets
 * ICNTRL(5) = 8 (NewtonMaxit) and ICNTRL(10) = 1 (embedded SDIRK error estimator); ICNTRL_U/RCNTRL_U override only when
 * > 0 (:67-74).  ICNTRL(6) = 1: zero starting values, ICNTRL(11) = 1: classic instead of Gustafsson controller,
 * RCNTRL(8..11) = ThetaMin, NewtonTol, Qmin, Qmax.  ISTATUS as the reference counts it (the three FUN calls per Newton
 * iteration are not counted, a real+complex decomposition counts once, RK_Solve counts twice).  Models: cbm4,
 * small_strato, roberts.  Per-cell IERR: 1, or the reference's -1..-13.  cbm4 systems whose factorisations leave the static
 * pivot pattern (seen for rtol >= 1e-5) are integrated a second time with general partial pivoting
 * (fatode_stats.cells_general counts them). */
int fatode_rk_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                              const double* atol, const double* rtol, const int* icntrl_u,
                              const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                              const double* model_params, int mem, void* stream);

/* INTEGRATE_ADJ of ADJ/RK_ADJ (ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:20):
 *   INTEGRATE_ADJ(NVAR,NP,NADJ,NNZERO,Y,Lambda,TIN,TOUT,ATOL_adj,RTOL_adj,ATOL,RTOL,FUN,JAC,AdjInit,ICNTRL_U,RCNTRL_U,
 *                 ISTATUS_U,RSTATUS_U,IERR_U,Mu,JACP,DRDY,DRDP,QFUN,Q)
 * Discrete adjoint of the 3-stage implicit Runge-Kutta methods (ICNTRL(3) = 1 Radau-2A, 0/2 Lobatto-3C, 3 Gauss,
 * 4 Radau-1A) with the presets of the reference (ICNTRL(5) = 8, (7) = 1 fixed-sweep adjoint solve, (9) = 2 discrete,
 * (10) = 1 SDIRK error estimator, (11) = 1 classic controller, Max_no_steps = 10000); ICNTRL(9) = 1: forward sweep only.
 * Same argument layout as fatode_ros_integrate_adj_batch; ADJINIT of the example drivers (Lambda = I, Mu = 0) is applied
 * after the forward sweep, systems whose forward sweep fails keep the caller's Lambda/Mu.  cbm4 only.  ICNTRL(7) = 2 selects
 * the direct adjoint solve (LSS_Decomp_Big / LSS_Solve_Big, ADJ/RK_ADJ/LS_Solver.F90:59-80,115-124: one factorisation of the
 * 3N x 3N system per accepted step, one DGETRS('T') per column; block-cooperative 96 x 96 LU on the GPU).  The adaptive
 * adjoint solve (ICNTRL(7) = 3) and the quadrature terms return FATODE_EUNSUPPORTED.  Per-cell IERR: 1, or
 * the reference's -1..-13.  Systems whose real or complex factorisations leave the static pivot pattern (none at rtol 1e-6,
 * all at rtol >= 1e-4) are integrated a second time with general partial pivoting (fatode_stats.cells_general counts them);
 * the forward sweep, the counters and Lambda are bit-identical to the reference's operation order on both paths. */
int fatode_rk_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda,
                                  double* mu, double tin, double tout, const double* atol_adj,
                                  const double* rtol_adj, const double* atol, const double* rtol,
                                  const int* icntrl_u, const double* rcntrl_u, int* istatus, double* rstatus,
                                  int* ierr, double* mu_sum, const double* model_params, int mem, void* stream);

/* Engine knobs (process-wide): trajectory-tape budget in bytes (0 = 60% of free device memory) and
 * the number of cells per wave (0 = automatic). */
int fatode_set_tape_budget(uint64_t bytes);
/* RK_ADJ: accepted steps per system the first pass reserves tape for (0 = 1024); systems with more steps are integrated
 * again with a Max_no_steps slot.  Results do not depend on it. */
int fatode_set_rk_tape_slot(int steps);
/* SDIRK_TLM: the same knob for the state sweep's tape (0 = 4608 accepted steps per system). */
int fatode_set_sdirk_tape_slot(int steps);
int fatode_set_wave_cells(int64_t cells);
/* Execution path: 0 (default) = per-thread kernels (one cell per thread, static-pattern LU) with automatic
 * hand-back of irregular cells to the general kernels; 1 = general warp-per-cell kernels only.  Both give
 * identical forward bits.  fatode_debug_irregular_every(k > 0) is a test hook that pushes every k-th cell
 * through the hand-back. */
int fatode_set_path(int mode);
/* 0 (default): the forward, expand and backward kernels of a wave run back to back; 1, 2: the forward sweep of the
 * next wave is overlapped with the backward sweep of the current one on separate streams (double-buffered tape).
 * Results are identical in all modes. */
int fatode_set_pipeline(int mode);
int fatode_debug_irregular_every(int k);
/* frees the device workspace (trajectory tape etc.) the engine caches between calls */
int fatode_release_workspace(void);

/* Per-call telemetry of the last batch call on this thread (kernel launches, waves, device time in
 * ms of each phase measured with CUDA events on the launch stream, tape bytes).  ms_forward: forward
 * sweeps; ms_forward_tape: tape replay (general path) / k_expand_tape (per-thread path); ms_adjoint:
 * backward sweeps. */
typedef struct {
  int64_t launches;
  int64_t waves;
  double ms_forward, ms_forward_tape, ms_adjoint, ms_other;
  uint64_t tape_bytes_peak;
  int64_t accepted_steps, attempts;
  int64_t cells_general;  /* cells integrated by the general (warp-per-cell, full pivoting) kernels */
  int64_t cells_deferred; /* cells re-queued because the tape pool of their wave was exhausted */
  int64_t launches_forward, launches_expand, launches_adjoint; /* kernel launches behind ms_forward(+_tape), ms_adjoint */
} fatode_stats;
int fatode_get_last_stats(fatode_stats* out);

/* pinned host memory helpers for callers that want asynchronous staging */
void* fatode_alloc_pinned(uint64_t bytes);
void fatode_free_pinned(void* p);

/* FP64 DFMA micro-benchmark on the current device: returns achieved TFLOP/s (<0 on error). */
double fatode_measure_fp64_peak(int iters);

/* ---- Multi-GPU (one process per GPU).  Cells are sharded by the host; the path's one collective — the sum of
 * the parameter sensitivities Mu over all cells — is done inside the library over NCCL once a communicator exists: every
 * *_integrate_adj_batch call then returns mu_sum summed over ALL ranks' cells.  NCCL is resolved at run time (dlopen).
 *  fatode_comm_unique_id : rank 0 creates the 128-byte NCCL id; the host hands it to the other ranks (file, MPI, socket)
 *  fatode_comm_init      : every rank, after selecting its GPU (cudaSetDevice)
 *  fatode_comm_allreduce_sum : in-place sum of a host or device buffer over the ranks (no-op without a communicator) */
int fatode_comm_unique_id(void* id128);
int fatode_comm_init(int nranks, int rank, const void* id128);
int fatode_comm_size(void);
int fatode_comm_rank(void);
int fatode_comm_allreduce_sum(double* buf, int64_t count, int mem, void* stream);
int fatode_comm_finalize(void);
const char* fatode_comm_last_error(void);

/* ---- Batched structure-of-arrays container "FATODEB1" (on disk / on the wire) and the reference's text files.
 * The reference writes one system's result as text, one E24.16 number per line (cbm4_ros_adj_dr.F90:134-164); the container
 * holds the arrays of a whole batch exactly in the layouts this header uses, so a section can be passed to the *_batch entry
 * points without re-layout (format: fatode_b200/csrc/fatode_io.cu).  Host code only, no GPU needed.
 *  fatode_file_write : any of y, lambda, mu, istatus, rstatus, ierr may be NULL (left out)
 *  fatode_file_info  : header fields (any pointer may be NULL); has[6] = presence of the six arrays in that order
 *  fatode_file_read  : cells [first_cell, first_cell + ncells) of one named array into dst
 *  fatode_text_*     : cbm4_ros_sol.txt / cbm4_output_sens.txt of ONE system (lambda[nadj][nvar], mu[nadj][np]) */
int fatode_file_write(const char* path, int model, int64_t ncells, int nvar, int np, int nadj, double tin, double tout,
                      const double* y, const double* lambda, const double* mu, const int* istatus, const double* rstatus,
                      const int* ierr);
int fatode_file_info(const char* path, int* model, int64_t* ncells, int* nvar, int* np, int* nadj, double* tin, double* tout,
                     int* has);
int fatode_file_read(const char* path, const char* name, int64_t first_cell, int64_t ncells, void* dst);
int fatode_text_write_solution(const char* path, int nvar, const double* y);
int fatode_text_write_sens(const char* path, int nvar, int np, int nadj, const double* lambda, const double* mu);
int fatode_text_read_solution(const char* path, int nvar, double* y);
int fatode_text_read_sens(const char* path, int nvar, int np, int nadj, double* lambda, double* mu);

/* Unit-level probes used by the parity tests (one warp, one system).
 *  fatode_probe_cbm4: f[32] = fun(t,y); jac_dense[32*32] column-major = jac(t,y); hess[258] = Hessian(t)
 *  fatode_probe_lu32: DGETF2 + DGETRS('N') of a column-major 32x32 matrix on the warp LU;
 *                     lu_out row-major in pivoted row order, who[k] = original row at position k */
int fatode_probe_cbm4(double t, const double* y, double* f, double* jac_dense, double* hess);
int fatode_probe_lu32(const double* A, const double* b, double* lu_out, int* who, double* x, int* info);
/*  fatode_probe_math: the correctly rounded COS (Update_SUN) and Err**(1/ELO) (step controller) kernels evaluated on the device:
 *                     out_cos[i] = cos(x[i]) for |x[i]| <= 5 pi / 4, out_root[i] = |x[i]| ** fl(1/elo), elo = 2, 3 or 4 */
int fatode_probe_math(int n, const double* x, double elo, double* out_cos, double* out_root);

#ifdef __cplusplus
}
#endif
#endif /* FATODE_B200_H */
