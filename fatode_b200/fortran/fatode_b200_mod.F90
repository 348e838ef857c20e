!~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
!  fatode_b200_mod — ISO_C_BINDING shim that keeps FATODE's INTEGRATE / INTEGRATE_ADJ host API and
!  forwards it to the B200 engine (include/fatode_b200.h).  Source only: the build image has no
!  Fortran compiler; a maintainer compiles this next to the application and links libfatode_b200.so.
!
!  Replaces, for the registered example models:
!     MODULE ROS_adj_f90_Integrator :: INTEGRATE_ADJ   (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:34)
!     MODULE ROS_f90_Integrator     :: INTEGRATE       (FWD/ROS/ROS_f90_Integrator.F90:32)
!     MODULE ROS_tlm_f90_Integrator :: INTEGRATE_TLM   (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:35)
!     MODULE SDIRK_f90_Integrator   :: INTEGRATE       (FWD/SDIRK/SDIRK_f90_Integrator.F90:25)   after fatode_b200_set_family(FATODE_FAMILY_SDIRK)
!     MODULE RK_f90_Integrator      :: INTEGRATE       (FWD/RK/RK_f90_Integrator.F90:32)         after fatode_b200_set_family(FATODE_FAMILY_RK)
!     MODULE RK_adj_f90_Integrator  :: INTEGRATE_ADJ   (ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:20) after fatode_b200_set_family(FATODE_FAMILY_RK)
!     MODULE RK_tlm_f90_Integrator  :: INTEGRATE_TLM   (TLM/RK_TLM/RK_TLM_f90_Integrator.F90:33) after fatode_b200_set_family(FATODE_FAMILY_RK)
!     MODULE SDIRK_ADJ_f90_Integrator :: INTEGRATE_ADJ (ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:29) after fatode_b200_set_family(FATODE_FAMILY_SDIRK)
!     MODULE SDIRK_TLM_f90_Integrator :: INTEGRATE_TLM (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:31) after fatode_b200_set_family(FATODE_FAMILY_SDIRK)
!     MODULE ERK_f90_Integrator     :: INTEGRATE       (FWD/ERK/ERK_f90_Integrator.F90:28)       as INTEGRATE_ERK (rename on the USE line)
!     MODULE ERK_TLM_f90_Integrator :: INTEGRATE_TLM   (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:28) as INTEGRATE_ERK_TLM
!     MODULE ERK_ADJ_f90_Integrator :: INTEGRATE_ADJ   (ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:33) as INTEGRATE_ERK_ADJ
!  (upstream the forward families are separate modules that all export INTEGRATE with the same argument list; the
!  application picks one with its USE line, here with fatode_b200_set_family)
!  Dummy-argument names and order are the reference's, so keyword calls such as
!     call integrate_adj(tin=tstart, tout=tend, nvar=nvar, nnzero=nnz, np=np, nadj=nadj, y=var, &
!                        lambda=lambda, mu=mu, ..., fun=fun, jac=jac, ..., icntrl_u=icntrl)
!  (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_ros_adj_dr.F90:282-287) compile unchanged.  The EXTERNAL callback
!  arguments are accepted and ignored: the model is selected with fatode_b200_set_model(); the
!  shipped models (roberts, small_strato, cbm4, swe) run as device functions.
!  Batched entry points (INTEGRATE_ADJ_BATCH / INTEGRATE_BATCH) take NCELLS systems at once:
!  Y(NVAR,NCELLS), Lambda(NVAR,NADJ,NCELLS), Mu(NP,NADJ,NCELLS), ISTATUS(20,NCELLS), ... — the
!  column-major Fortran layout is exactly the C ABI's [cell][nadj][nvar] layout.
!~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
MODULE fatode_b200_mod
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: INTEGRATE, INTEGRATE_ADJ, INTEGRATE_TLM, INTEGRATE_BATCH, INTEGRATE_ADJ_BATCH, fatode_b200_set_model
  PUBLIC :: INTEGRATE_ERK, INTEGRATE_ERK_TLM, INTEGRATE_ERK_ADJ
  PUBLIC :: fatode_b200_set_family, FATODE_FAMILY_ROS, FATODE_FAMILY_SDIRK, FATODE_FAMILY_RK
  PUBLIC :: FATODE_MODEL_ROBERTS, FATODE_MODEL_SMALL_STRATO, FATODE_MODEL_CBM4, FATODE_MODEL_SWE

  INTEGER(C_INT), PARAMETER :: FATODE_MODEL_ROBERTS = 1, FATODE_MODEL_SMALL_STRATO = 2, FATODE_MODEL_CBM4 = 3, FATODE_MODEL_SWE = 4
  INTEGER(C_INT), PARAMETER :: FATODE_MEM_HOST = 0
  INTEGER, PARAMETER :: FATODE_FAMILY_ROS = 0, FATODE_FAMILY_SDIRK = 1, FATODE_FAMILY_RK = 2
  INTEGER(C_INT), SAVE :: current_model = FATODE_MODEL_CBM4
  INTEGER, SAVE :: current_family = FATODE_FAMILY_ROS

  INTERFACE
    FUNCTION fatode_ros_integrate_batch(model, ncells, nvar, y, tin, tout, atol, rtol, icntrl_u, rcntrl_u, &
                                        istatus, rstatus, ierr, model_params, mem, stream) BIND(C, NAME='fatode_ros_integrate_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), rstatus(*)
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      TYPE(C_PTR), VALUE :: icntrl_u, rcntrl_u, model_params, stream
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_ros_integrate_batch
    END FUNCTION
    FUNCTION fatode_sdirk_integrate_batch(model, ncells, nvar, y, tin, tout, atol, rtol, icntrl_u, rcntrl_u, &
                                        istatus, rstatus, ierr, model_params, mem, stream) BIND(C, NAME='fatode_sdirk_integrate_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), rstatus(*)
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      TYPE(C_PTR), VALUE :: icntrl_u, rcntrl_u, model_params, stream
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_sdirk_integrate_batch
    END FUNCTION
    FUNCTION fatode_erk_integrate_batch(model, ncells, nvar, y, tin, tout, atol, rtol, icntrl_u, rcntrl_u, &
                                        istatus, rstatus, ierr, model_params, mem, stream) BIND(C, NAME='fatode_erk_integrate_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), rstatus(*)
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      TYPE(C_PTR), VALUE :: icntrl_u, rcntrl_u, model_params, stream
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_erk_integrate_batch
    END FUNCTION
    FUNCTION fatode_erk_integrate_tlm_batch(model, ncells, nvar, ntlm, y, y_tlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, &
                                            icntrl_u, rcntrl_u, istatus, rstatus, ierr, model_params, mem, stream) &
                                            BIND(C, NAME='fatode_erk_integrate_tlm_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, ntlm, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), y_tlm(*), rstatus(*)
      TYPE(C_PTR), VALUE :: atol_tlm, rtol_tlm, icntrl_u, rcntrl_u, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_erk_integrate_tlm_batch
    END FUNCTION
    FUNCTION fatode_erk_integrate_adj_batch(model, ncells, nvar, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl_u, rcntrl_u, &
                                            istatus, rstatus, ierr, model_params, mem, stream) &
                                            BIND(C, NAME='fatode_erk_integrate_adj_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, np, nadj, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), lambda(*), rstatus(*)
      TYPE(C_PTR), VALUE :: mu, icntrl_u, rcntrl_u, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_erk_integrate_adj_batch
    END FUNCTION
    FUNCTION fatode_rk_integrate_batch(model, ncells, nvar, y, tin, tout, atol, rtol, icntrl_u, rcntrl_u, &
                                        istatus, rstatus, ierr, model_params, mem, stream) BIND(C, NAME='fatode_rk_integrate_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), rstatus(*)
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      TYPE(C_PTR), VALUE :: icntrl_u, rcntrl_u, model_params, stream
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_rk_integrate_batch
    END FUNCTION
    FUNCTION fatode_ros_integrate_adj_batch(model, ncells, nvar, np, nadj, y, lambda, mu, tin, tout, atol_adj, rtol_adj, &
                                            atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus, ierr, mu_sum, model_params, &
                                            mem, stream) BIND(C, NAME='fatode_ros_integrate_adj_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, np, nadj, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), lambda(*), rstatus(*)
      TYPE(C_PTR), VALUE :: mu, atol_adj, rtol_adj, icntrl_u, rcntrl_u, mu_sum, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_ros_integrate_adj_batch
    END FUNCTION
    FUNCTION fatode_rk_integrate_adj_batch(model, ncells, nvar, np, nadj, y, lambda, mu, tin, tout, atol_adj, rtol_adj, &
                                            atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus, ierr, mu_sum, model_params, &
                                            mem, stream) BIND(C, NAME='fatode_rk_integrate_adj_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, np, nadj, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), lambda(*), rstatus(*)
      TYPE(C_PTR), VALUE :: mu, atol_adj, rtol_adj, icntrl_u, rcntrl_u, mu_sum, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_rk_integrate_adj_batch
    END FUNCTION
    FUNCTION fatode_sdirk_integrate_adj_batch(model, ncells, nvar, np, nadj, y, lambda, mu, tin, tout, atol_adj, rtol_adj, &
                                            atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus, ierr, mu_sum, model_params, &
                                            mem, stream) BIND(C, NAME='fatode_sdirk_integrate_adj_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, np, nadj, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), lambda(*), rstatus(*)
      TYPE(C_PTR), VALUE :: mu, atol_adj, rtol_adj, icntrl_u, rcntrl_u, mu_sum, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_sdirk_integrate_adj_batch
    END FUNCTION
    FUNCTION fatode_ros_integrate_tlm_batch(model, ncells, nvar, ntlm, y, y_tlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, &
                                            icntrl_u, rcntrl_u, istatus, rstatus, ierr, model_params, mem, stream) &
                                            BIND(C, NAME='fatode_ros_integrate_tlm_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, ntlm, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), y_tlm(*), rstatus(*)
      TYPE(C_PTR), VALUE :: atol_tlm, rtol_tlm, icntrl_u, rcntrl_u, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_ros_integrate_tlm_batch
    END FUNCTION
    FUNCTION fatode_sdirk_integrate_tlm_batch(model, ncells, nvar, ntlm, y, y_tlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, &
                                              icntrl_u, rcntrl_u, istatus, rstatus, ierr, model_params, mem, stream) &
                                              BIND(C, NAME='fatode_sdirk_integrate_tlm_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, ntlm, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), y_tlm(*), rstatus(*)
      TYPE(C_PTR), VALUE :: atol_tlm, rtol_tlm, icntrl_u, rcntrl_u, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_sdirk_integrate_tlm_batch
    END FUNCTION
    FUNCTION fatode_rk_integrate_tlm_batch(model, ncells, nvar, ntlm, y, y_tlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, &
                                           icntrl_u, rcntrl_u, istatus, rstatus, ierr, model_params, mem, stream) &
                                           BIND(C, NAME='fatode_rk_integrate_tlm_batch')
      IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE :: model, nvar, ntlm, mem
      INTEGER(C_INT64_T), VALUE :: ncells
      REAL(C_DOUBLE), VALUE :: tin, tout
      REAL(C_DOUBLE) :: y(*), y_tlm(*), rstatus(*)
      TYPE(C_PTR), VALUE :: atol_tlm, rtol_tlm, icntrl_u, rcntrl_u, model_params, stream
      REAL(C_DOUBLE), INTENT(IN) :: atol(*), rtol(*)
      INTEGER(C_INT) :: istatus(*), ierr(*)
      INTEGER(C_INT) :: fatode_rk_integrate_tlm_batch
    END FUNCTION
  END INTERFACE

CONTAINS

  SUBROUTINE fatode_b200_set_model(model)
    INTEGER, INTENT(IN) :: model
    current_model = INT(model, C_INT)
  END SUBROUTINE

  SUBROUTINE fatode_b200_set_family(family)
    INTEGER, INTENT(IN) :: family
    current_family = family
  END SUBROUTINE

  !~~~> FWD/ROS, FWD/SDIRK or FWD/RK INTEGRATE, one system (FWD/ROS/ROS_f90_Integrator.F90:32-33,
  !     FWD/SDIRK/SDIRK_f90_Integrator.F90:25-26, FWD/RK/RK_f90_Integrator.F90:32-33: identical argument lists)
  SUBROUTINE INTEGRATE( TIN, TOUT, N, NNZERO, VAR, RTOL, ATOL, FUN, JAC, ICNTRL_U, RCNTRL_U, ISTATUS_U, RSTATUS_U, IERR_U )
    INTEGER, INTENT(IN) :: N, NNZERO
    DOUBLE PRECISION, INTENT(INOUT) :: VAR(N)
    DOUBLE PRECISION, INTENT(IN) :: RTOL(N), ATOL(N), TIN, TOUT
    EXTERNAL :: FUN, JAC
    INTEGER, INTENT(IN), OPTIONAL, TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: ISTATUS_U(20), IERR_U
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL :: RSTATUS_U(20)
    INTEGER(C_INT) :: ist(20), ierr(1), rc
    REAL(C_DOUBLE) :: rst(20)
    TYPE(C_PTR) :: pic, prc
    pic = C_NULL_PTR; prc = C_NULL_PTR
    IF (PRESENT(ICNTRL_U)) pic = C_LOC(ICNTRL_U)
    IF (PRESENT(RCNTRL_U)) prc = C_LOC(RCNTRL_U)
    SELECT CASE (current_family)
    CASE (FATODE_FAMILY_SDIRK)
      rc = fatode_sdirk_integrate_batch(current_model, 1_C_INT64_T, INT(N, C_INT), VAR, TIN, TOUT, ATOL, RTOL, pic, prc, &
                                        ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
    CASE (FATODE_FAMILY_RK)
      rc = fatode_rk_integrate_batch(current_model, 1_C_INT64_T, INT(N, C_INT), VAR, TIN, TOUT, ATOL, RTOL, pic, prc, &
                                     ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
    CASE DEFAULT
      rc = fatode_ros_integrate_batch(current_model, 1_C_INT64_T, INT(N, C_INT), VAR, TIN, TOUT, ATOL, RTOL, pic, prc, &
                                      ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
    END SELECT
    IF (rc /= 0) ierr(1) = rc
    IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = ist(:)
    IF (PRESENT(RSTATUS_U)) RSTATUS_U(:) = rst(:)
    IF (PRESENT(IERR_U)) IERR_U = ierr(1)
  END SUBROUTINE INTEGRATE

  !~~~> FWD/ERK INTEGRATE, one system (FWD/ERK/ERK_f90_Integrator.F90:28-29: NVAR instead of N, no NNZERO, no JAC).  Its dummy
  !     names differ from the implicit families', so it is a routine of its own; an application written against
  !     ERK_f90_Integrator keeps its keyword call (swe2D_erk_dr.F90:112-114) with
  !        USE fatode_b200_mod, ONLY: INTEGRATE => INTEGRATE_ERK, fatode_b200_set_model
  SUBROUTINE INTEGRATE_ERK( TIN, TOUT, NVAR, VAR, RTOL, ATOL, FUN, ICNTRL_U, RCNTRL_U, ISTATUS_U, RSTATUS_U, Ierr_U )
    INTEGER, INTENT(IN) :: NVAR
    DOUBLE PRECISION, INTENT(INOUT) :: VAR(NVAR)
    DOUBLE PRECISION, INTENT(IN) :: RTOL(NVAR), ATOL(NVAR), TIN, TOUT
    EXTERNAL :: FUN
    INTEGER, INTENT(IN), OPTIONAL, TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: ISTATUS_U(20), Ierr_U
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL :: RSTATUS_U(20)
    INTEGER(C_INT) :: ist(20), ierr(1), rc
    REAL(C_DOUBLE) :: rst(20)
    TYPE(C_PTR) :: pic, prc
    pic = C_NULL_PTR; prc = C_NULL_PTR
    IF (PRESENT(ICNTRL_U)) pic = C_LOC(ICNTRL_U)
    IF (PRESENT(RCNTRL_U)) prc = C_LOC(RCNTRL_U)
    ist = 0; rst = 0.0d0; ierr = 0
    rc = fatode_erk_integrate_batch(current_model, 1_C_INT64_T, INT(NVAR, C_INT), VAR, TIN, TOUT, ATOL, RTOL, pic, prc, &
                                    ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
    IF (rc /= 0) ierr(1) = rc
    IF (ierr(1) < 0) PRINT *, 'ERK: Unsuccessful exit at T=', TIN, ' (Ierr=', ierr(1), ')'   ! as :75-77
    IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = ist(:)
    IF (PRESENT(RSTATUS_U)) RSTATUS_U(:) = rst(:)
    IF (PRESENT(Ierr_U)) Ierr_U = ierr(1)
  END SUBROUTINE INTEGRATE_ERK

  !~~~> TLM/ERK_TLM INTEGRATE_TLM, one system (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:28-30; argument order and names differ from
  !     the ROS/SDIRK TLM families): USE fatode_b200_mod, ONLY: INTEGRATE_TLM => INTEGRATE_ERK_TLM keeps the keyword call of
  !     swe2D_erk_tlm_dr.F90:131-135.  JAC is accepted and not called: with ICNTRL(5) /= 0 the
  !     engine applies the shallow-water example's own JAC (compute_JacF), compiled as a device function.
  SUBROUTINE INTEGRATE_ERK_TLM( TIN, TOUT, NVAR, NTLM, VAR, VAR_TLM, ATOL_TLM, RTOL_TLM, ATOL, RTOL, FUN, JAC, &
                                ICNTRL_U, RCNTRL_U, ISTATUS_U, RSTATUS_U, IERR_U )
    INTEGER, INTENT(IN) :: NVAR, NTLM
    DOUBLE PRECISION, INTENT(INOUT) :: VAR(NVAR), VAR_TLM(NVAR,NTLM)
    DOUBLE PRECISION, INTENT(INOUT), TARGET :: RTOL_TLM(NVAR,NTLM), ATOL_TLM(NVAR,NTLM)
    DOUBLE PRECISION, INTENT(IN) :: RTOL(NVAR), ATOL(NVAR), TIN, TOUT
    EXTERNAL :: FUN, JAC
    INTEGER, INTENT(IN), OPTIONAL, TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: ISTATUS_U(20), IERR_U
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL :: RSTATUS_U(20)
    INTEGER(C_INT) :: ist(20), ierr(1), rc
    REAL(C_DOUBLE) :: rst(20)
    TYPE(C_PTR) :: pic, prc
    pic = C_NULL_PTR; prc = C_NULL_PTR
    IF (PRESENT(ICNTRL_U)) pic = C_LOC(ICNTRL_U)
    IF (PRESENT(RCNTRL_U)) prc = C_LOC(RCNTRL_U)
    ist = 0; rst = 0.0d0; ierr = 0
    rc = fatode_erk_integrate_tlm_batch(current_model, 1_C_INT64_T, INT(NVAR, C_INT), INT(NTLM, C_INT), VAR, VAR_TLM, TIN, TOUT, &
                                        C_LOC(ATOL_TLM), C_LOC(RTOL_TLM), ATOL, RTOL, pic, prc, ist, rst, ierr, C_NULL_PTR, &
                                        FATODE_MEM_HOST, C_NULL_PTR)
    IF (rc /= 0) ierr(1) = rc
    IF (ierr(1) < 0) PRINT *, 'ERK-TLM: Unsuccessful exit at T=', TIN, ' (Ierr=', ierr(1), ')'   ! as :79-81
    IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = ist(:)
    IF (PRESENT(RSTATUS_U)) RSTATUS_U(:) = rst(:)
    IF (PRESENT(IERR_U)) IERR_U = ierr(1)
  END SUBROUTINE INTEGRATE_ERK_TLM

  !~~~> ADJ/ERK_ADJ INTEGRATE_ADJ, one system (ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:33-35; argument order and names differ from
  !     the implicit adjoint families): USE fatode_b200_mod, ONLY: INTEGRATE_ADJ => INTEGRATE_ERK_ADJ keeps the keyword call of
  !     swe2D_erk_adj_dr.F90:164-167.  FUN, JAC, AdjInit are accepted and not called (model id: the example's callbacks are device
  !     functions; ADJINIT sets Lambda(k,k) = 1).  The call without parameters (ERKADJ2) is served; Mu / JACP / quadrature: -105.
  SUBROUTINE INTEGRATE_ERK_ADJ( NVAR, NP, NADJ, NNZ, Y, Lambda, TIN, TOUT, ATOL, RTOL, FUN, JAC, AdjInit, ICNTRL_U, RCNTRL_U, &
                                ISTATUS_U, RSTATUS_U, Ierr_U, Mu, JACP, DRDY, DRDP, QFUN, Q )
    INTEGER, INTENT(IN) :: NVAR, NP, NADJ
    INTEGER, INTENT(IN), OPTIONAL :: NNZ
    DOUBLE PRECISION, INTENT(INOUT) :: Y(NVAR), Lambda(NVAR,NADJ)
    DOUBLE PRECISION, INTENT(INOUT), OPTIONAL :: Q(NADJ), Mu(NP,NADJ)
    DOUBLE PRECISION, INTENT(IN) :: ATOL(NVAR), RTOL(NVAR), TIN, TOUT
    EXTERNAL :: FUN, JAC, AdjInit
    EXTERNAL :: QFUN, JACP, DRDY, DRDP
    OPTIONAL :: QFUN, JACP, DRDY, DRDP
    INTEGER, INTENT(IN), OPTIONAL, TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: ISTATUS_U(20), Ierr_U
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL :: RSTATUS_U(20)
    INTEGER(C_INT) :: ist(20), ierr(1), rc
    REAL(C_DOUBLE) :: rst(20)
    TYPE(C_PTR) :: pic, prc
    pic = C_NULL_PTR; prc = C_NULL_PTR
    IF (PRESENT(ICNTRL_U)) pic = C_LOC(ICNTRL_U)
    IF (PRESENT(RCNTRL_U)) prc = C_LOC(RCNTRL_U)
    ist = 0; rst = 0.0d0; ierr = 0
    IF ((NP > 0 .AND. PRESENT(Mu) .AND. PRESENT(JACP)) .OR. (PRESENT(Q) .AND. PRESENT(QFUN))) THEN
      ierr(1) = -105   ! FATODE_EUNSUPPORTED: parameter sensitivities / quadrature terms of ERKADJ1 (:100-117)
    ELSE
      rc = fatode_erk_integrate_adj_batch(current_model, 1_C_INT64_T, INT(NVAR, C_INT), 0_C_INT, INT(NADJ, C_INT), Y, Lambda, &
                                          C_NULL_PTR, TIN, TOUT, ATOL, RTOL, pic, prc, ist, rst, ierr, C_NULL_PTR, &
                                          FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
    END IF
    IF (ierr(1) < 0) PRINT *, 'ERK: Unsuccessful exit at T=', TIN, ' (Ierr=', ierr(1), ')'   ! as :129-131
    IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = ist(:)
    IF (PRESENT(RSTATUS_U)) RSTATUS_U(:) = rst(:)
    IF (PRESENT(Ierr_U)) Ierr_U = ierr(1)
  END SUBROUTINE INTEGRATE_ERK_ADJ

  !~~~> ADJ/ROS_ADJ INTEGRATE_ADJ, one system (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:34-37)
  !     With fatode_b200_set_family(FATODE_FAMILY_RK) the same routine stands in for ADJ/RK_ADJ's INTEGRATE_ADJ
  !     (ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:20-23): its dummy arguments are a subset of these names (no Hessian
  !     callbacks, plus IERR_U), so the keyword call of cbm4_rk_adj_dr.F90:231-235 compiles unchanged; only the positional
  !     order differs upstream.
  SUBROUTINE INTEGRATE_ADJ( NVAR, NP, NADJ, NNZERO, Y, Lambda, Mu, TIN, TOUT, ATOL_adj, RTOL_adj, ATOL, RTOL, FUN, JAC, &
                            ADJINIT, HESSTR_VEC, JACP, DRDY, DRDP, HESSTR_VEC_F_PY, HESSTR_VEC_R_PY, HESSTR_VEC_R, &
                            ICNTRL_U, RCNTRL_U, ISTATUS_U, RSTATUS_U, Q, QFUN, IERR_U )
    INTEGER, INTENT(IN) :: NVAR, NNZERO, NP, NADJ
    DOUBLE PRECISION, INTENT(INOUT) :: Y(NVAR), Lambda(NVAR,NADJ)
    DOUBLE PRECISION, INTENT(INOUT), OPTIONAL, TARGET :: Mu(NP,NADJ)
    DOUBLE PRECISION, INTENT(INOUT), OPTIONAL :: Q(NADJ)
    DOUBLE PRECISION, INTENT(IN) :: TIN, TOUT, ATOL(NVAR), RTOL(NVAR)
    DOUBLE PRECISION, INTENT(IN), TARGET :: ATOL_adj(NVAR,NADJ), RTOL_adj(NVAR,NADJ)
    INTEGER, INTENT(IN), OPTIONAL, TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: ISTATUS_U(20), IERR_U
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL :: RSTATUS_U(20)
    EXTERNAL FUN, JAC, ADJINIT, HESSTR_VEC
    EXTERNAL QFUN, JACP, DRDY, DRDP, HESSTR_VEC_F_PY, HESSTR_VEC_R_PY, HESSTR_VEC_R
    OPTIONAL HESSTR_VEC, QFUN, JACP, DRDY, DRDP, HESSTR_VEC_F_PY, HESSTR_VEC_R_PY, HESSTR_VEC_R
    INTEGER(C_INT) :: ist(20), ierr(1), rc, npc
    REAL(C_DOUBLE) :: rst(20)
    TYPE(C_PTR) :: pic, prc, pmu
    pic = C_NULL_PTR; prc = C_NULL_PTR; pmu = C_NULL_PTR; npc = 0
    IF (PRESENT(ICNTRL_U)) pic = C_LOC(ICNTRL_U)
    IF (PRESENT(RCNTRL_U)) prc = C_LOC(RCNTRL_U)
    ! mode selection of the reference (:100-144): Mu is propagated only with NP>0, Mu, JACP and HESSTR_VEC_F_PY present
    IF (NP > 0 .AND. PRESENT(Mu) .AND. PRESENT(JACP) .AND. (PRESENT(HESSTR_VEC_F_PY) .OR. current_family /= FATODE_FAMILY_ROS)) THEN
      pmu = C_LOC(Mu); npc = INT(NP, C_INT)
    END IF
    IF (PRESENT(Q) .AND. PRESENT(QFUN)) THEN
      PRINT *, 'fatode_b200: quadrature cost terms (Q, QFUN) are not supported by the batched engine'
      IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = 0
      RETURN
    END IF
    IF (current_family == FATODE_FAMILY_RK) THEN
      rc = fatode_rk_integrate_adj_batch(current_model, 1_C_INT64_T, INT(NVAR, C_INT), npc, INT(NADJ, C_INT), Y, Lambda, pmu, &
                                         TIN, TOUT, C_NULL_PTR, C_NULL_PTR, ATOL, RTOL, pic, prc, ist, rst, ierr, &
                                         C_NULL_PTR, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
      IF (ierr(1) < 0) PRINT *, 'Runge-Kutta-ADJ: Unsuccessful exit at T=', TIN, ' (IERR=', ierr(1), ')'   ! RK_ADJ :135-137
    ELSE IF (current_family == FATODE_FAMILY_SDIRK) THEN   ! ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:29-31 (cbm4_sdirk_adj_dr.F90:230-234)
      ! ATOL_adj / RTOL_adj are read by the simplified-Newton adjoint stages (ICNTRL(7) /= 1, :869-871)
      rc = fatode_sdirk_integrate_adj_batch(current_model, 1_C_INT64_T, INT(NVAR, C_INT), npc, INT(NADJ, C_INT), Y, Lambda, pmu, &
                                            TIN, TOUT, C_LOC(ATOL_adj), C_LOC(RTOL_adj), ATOL, RTOL, pic, prc, ist, rst, ierr, &
                                            C_NULL_PTR, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
      IF (ierr(1) < 0) PRINT *, 'SDIRK: Unsuccessful exit at T=', TIN, ' (Ierr=', ierr(1), ')'
    ELSE
      rc = fatode_ros_integrate_adj_batch(current_model, 1_C_INT64_T, INT(NVAR, C_INT), npc, INT(NADJ, C_INT), Y, Lambda, pmu, &
                                          TIN, TOUT, C_NULL_PTR, C_NULL_PTR, ATOL, RTOL, pic, prc, ist, rst, ierr, &
                                          C_NULL_PTR, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
      IF (ierr(1) < 0) PRINT *, 'RosenbrockADJ: Unsucessful step at T=', TIN, ' (IERR=', ierr(1), ')'   ! as :146-148
    END IF
    IF (PRESENT(IERR_U)) IERR_U = ierr(1)
    IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = ist(:)
    IF (PRESENT(RSTATUS_U)) RSTATUS_U(:) = rst(:)
  END SUBROUTINE INTEGRATE_ADJ

  !~~~> TLM/ROS_TLM INTEGRATE_TLM, one system (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:35-36): both error-control modes,
  !     including the preset ICNTRL(12) = 1 of a call without ICNTRL_U (ATOL_tlm / RTOL_tlm are then required, as upstream);
  !     ROS_TLM has no IERR_U (failures are printed, :86-89).
  !     With fatode_b200_set_family(FATODE_FAMILY_SDIRK) the same routine stands in for TLM/SDIRK_TLM's INTEGRATE_TLM
  !     (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:31-32: no HESS_VEC, IERR_U last); both keyword sets compile unchanged.
  SUBROUTINE INTEGRATE_TLM( N, NTLM, NNZERO, Y, Y_tlm, TIN, TOUT, ATOL_tlm, RTOL_tlm, ATOL, RTOL, FUN, JAC, HESS_VEC, &
                            ICNTRL_U, RCNTRL_U, ISTATUS_U, RSTATUS_U, IERR_U )
    INTEGER, INTENT(IN) :: N, NNZERO, NTLM
    DOUBLE PRECISION, INTENT(INOUT) :: Y(N), Y_tlm(N,NTLM)
    DOUBLE PRECISION, INTENT(IN) :: TIN, TOUT
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RTOL_tlm(N,NTLM), ATOL_tlm(N,NTLM)
    DOUBLE PRECISION, INTENT(IN) :: ATOL(N), RTOL(N)
    EXTERNAL :: FUN, JAC
    EXTERNAL :: HESS_VEC
    OPTIONAL :: HESS_VEC
    INTEGER, INTENT(IN), OPTIONAL, TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), OPTIONAL, TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: ISTATUS_U(20)
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL :: RSTATUS_U(20)
    INTEGER, INTENT(OUT), OPTIONAL :: IERR_U
    INTEGER(C_INT) :: ist(20), ierr(1), rc
    REAL(C_DOUBLE) :: rst(20)
    TYPE(C_PTR) :: pic, prc, pat, prt
    pic = C_NULL_PTR; prc = C_NULL_PTR; pat = C_NULL_PTR; prt = C_NULL_PTR
    IF (PRESENT(ICNTRL_U)) pic = C_LOC(ICNTRL_U)
    IF (PRESENT(RCNTRL_U)) prc = C_LOC(RCNTRL_U)
    IF (PRESENT(ATOL_tlm)) pat = C_LOC(ATOL_tlm)
    IF (PRESENT(RTOL_tlm)) prt = C_LOC(RTOL_tlm)
    ist = 0; rst = 0.0d0; ierr = 0
    IF (current_family == FATODE_FAMILY_SDIRK) THEN
      rc = fatode_sdirk_integrate_tlm_batch(current_model, 1_C_INT64_T, INT(N, C_INT), INT(NTLM, C_INT), Y, Y_tlm, TIN, TOUT, &
                                            pat, prt, ATOL, RTOL, pic, prc, ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
      IF (ierr(1) < 0) PRINT *, 'SDIRK-TLM: Unsuccessful exit at T=', TIN, ' (IERR=', ierr(1), ')'   ! as :94-96
    ELSE IF (current_family == FATODE_FAMILY_RK) THEN   ! TLM/RK_TLM INTEGRATE_TLM (RK_TLM_f90_Integrator.F90:33)
      rc = fatode_rk_integrate_tlm_batch(current_model, 1_C_INT64_T, INT(N, C_INT), INT(NTLM, C_INT), Y, Y_tlm, TIN, TOUT, &
                                         pat, prt, ATOL, RTOL, pic, prc, ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
      IF (ierr(1) < 0) PRINT *, 'Runge-Kutta-TLM: Unsuccessful exit at T=', TIN, ' (IERR=', ierr(1), ')'   ! as :95-97
    ELSE
      rc = fatode_ros_integrate_tlm_batch(current_model, 1_C_INT64_T, INT(N, C_INT), INT(NTLM, C_INT), Y, Y_tlm, TIN, TOUT, &
                                          pat, prt, ATOL, RTOL, pic, prc, ist, rst, ierr, C_NULL_PTR, FATODE_MEM_HOST, C_NULL_PTR)
      IF (rc /= 0) ierr(1) = rc
      IF (ierr(1) < 0) PRINT *, 'Rosenbrock: Unsucessful step at T=', TIN, ' (IERR=', ierr(1), ')'   ! as :86-89
    END IF
    IF (PRESENT(IERR_U)) IERR_U = ierr(1)
    IF (PRESENT(ISTATUS_U)) ISTATUS_U(:) = ist(:)
    IF (PRESENT(RSTATUS_U)) RSTATUS_U(:) = rst(:)
  END SUBROUTINE INTEGRATE_TLM

  !~~~> batched forms: NCELLS independent systems in one call
  SUBROUTINE INTEGRATE_BATCH( NCELLS, TIN, TOUT, N, VAR, RTOL, ATOL, ICNTRL_U, RCNTRL_U, ISTATUS_U, RSTATUS_U, IERR_U )
    INTEGER, INTENT(IN) :: NCELLS, N
    DOUBLE PRECISION, INTENT(INOUT) :: VAR(N,NCELLS)
    DOUBLE PRECISION, INTENT(IN) :: RTOL(N), ATOL(N), TIN, TOUT
    INTEGER, INTENT(IN), TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT) :: ISTATUS_U(20,NCELLS), IERR_U(NCELLS)
    DOUBLE PRECISION, INTENT(OUT) :: RSTATUS_U(20,NCELLS)
    INTEGER(C_INT) :: rc
    rc = fatode_ros_integrate_batch(current_model, INT(NCELLS, C_INT64_T), INT(N, C_INT), VAR, TIN, TOUT, ATOL, RTOL, &
                                    C_LOC(ICNTRL_U), C_LOC(RCNTRL_U), ISTATUS_U, RSTATUS_U, IERR_U, C_NULL_PTR, &
                                    FATODE_MEM_HOST, C_NULL_PTR)
    IF (rc /= 0) IERR_U(:) = rc
  END SUBROUTINE INTEGRATE_BATCH

  SUBROUTINE INTEGRATE_ADJ_BATCH( NCELLS, NVAR, NP, NADJ, Y, Lambda, Mu, TIN, TOUT, ATOL, RTOL, ICNTRL_U, RCNTRL_U, &
                                  ISTATUS_U, RSTATUS_U, IERR_U, Mu_sum )
    INTEGER, INTENT(IN) :: NCELLS, NVAR, NP, NADJ
    DOUBLE PRECISION, INTENT(INOUT) :: Y(NVAR,NCELLS), Lambda(NVAR,NADJ,NCELLS)
    DOUBLE PRECISION, INTENT(INOUT), TARGET :: Mu(NP,NADJ,NCELLS)
    DOUBLE PRECISION, INTENT(IN) :: TIN, TOUT, ATOL(NVAR), RTOL(NVAR)
    INTEGER, INTENT(IN), TARGET :: ICNTRL_U(20)
    DOUBLE PRECISION, INTENT(IN), TARGET :: RCNTRL_U(20)
    INTEGER, INTENT(OUT) :: ISTATUS_U(20,NCELLS), IERR_U(NCELLS)
    DOUBLE PRECISION, INTENT(OUT) :: RSTATUS_U(20,NCELLS)
    DOUBLE PRECISION, INTENT(OUT), OPTIONAL, TARGET :: Mu_sum(NP,NADJ)
    INTEGER(C_INT) :: rc
    TYPE(C_PTR) :: psum
    psum = C_NULL_PTR
    IF (PRESENT(Mu_sum)) psum = C_LOC(Mu_sum)
    rc = fatode_ros_integrate_adj_batch(current_model, INT(NCELLS, C_INT64_T), INT(NVAR, C_INT), INT(NP, C_INT), &
                                        INT(NADJ, C_INT), Y, Lambda, C_LOC(Mu), TIN, TOUT, C_NULL_PTR, C_NULL_PTR, ATOL, RTOL, &
                                        C_LOC(ICNTRL_U), C_LOC(RCNTRL_U), ISTATUS_U, RSTATUS_U, IERR_U, psum, C_NULL_PTR, &
                                        FATODE_MEM_HOST, C_NULL_PTR)
    IF (rc /= 0) IERR_U(:) = rc
  END SUBROUTINE INTEGRATE_ADJ_BATCH

END MODULE fatode_b200_mod
