// Host-side option handling of the Rosenbrock drivers: the batched engine keeps the reference's
// ICNTRL/RCNTRL contract verbatim.
//   FWD/ROS/ROS_f90_Integrator.F90      INTEGRATE :56-72 (override when > 0),  Rosenbrock :249-375
//   ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90  INTEGRATE_ADJ :75-96 (override when >= 0), RosenbrockADJ1 :383-535
//   TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90  INTEGRATE_TLM :61-91, RosenbrockTLM :270-400
// Family-specific literals: FWD/ROS spells its defaults and tableau in REAL(4) (0.2, 0.1, 0.9,
// 1.0E-5, 1.0E-6, 0.1*H), the ADJ and TLM modules in DOUBLE PRECISION.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>
#include "fatode_common.cuh"
#include "gen/ros_tableau_gen.h"

namespace fatode {

// merges user arrays into defaults; returns the reference's IERR (1 = fine, negative = error code)
inline int ros_parse_options(int family, int N, const int* icntrl_u, const double* rcntrl_u, double tin, double tout,
                             const double* atol, const double* rtol, RosParams& rp, int icntrl_out[20]) {
  int ic[20];
  double rc[20];
  std::memset(ic, 0, sizeof ic);
  std::memset(rc, 0, sizeof rc);
  // ROS_TLM presets (TLM :70-74): ICNTRL(2) = 1 (scalar tolerances!), (3) = 5, (12) = 1; RCNTRL(3) = STEPMIN is a SAVEd
  // variable carrying the last step size to the NEXT call — every system of a batch is a first call (0)
  if (family == FAM_TLM) { ic[1] = 1; ic[2] = 5; ic[11] = 1; }
  for (int i = 0; i < 20; ++i) {
    if (icntrl_u) { if (family == FAM_FWD ? icntrl_u[i] > 0 : icntrl_u[i] >= 0) ic[i] = icntrl_u[i]; }
    if (rcntrl_u) { if (family == FAM_FWD ? rcntrl_u[i] > 0 : rcntrl_u[i] >= 0) rc[i] = rcntrl_u[i]; }
  }
  if (icntrl_out) std::memcpy(icntrl_out, ic, sizeof ic);
  const bool sp = (family == FAM_FWD);
  std::memset(&rp, 0, sizeof rp);
  rp.family = family;
  rp.autonomous = (ic[0] != 0);
  rp.vector_tol = (ic[1] == 0);
  int method;
  switch (ic[2]) {
    case 1: case 2: case 3: case 4: case 5: method = ic[2]; break;
    case 0: method = (family == FAM_FWD) ? 3 : (family == FAM_ADJ ? 4 : 5); break;   // Ros4 / Rodas3 / Rodas4
    default: return -2;
  }
  const RosTableauData& t = sp ? ROS_TABLEAU_E[method - 1] : ROS_TABLEAU_D[method - 1];
  rp.S = t.S;
  std::memcpy(rp.A, t.A, sizeof rp.A);
  std::memcpy(rp.C, t.C, sizeof rp.C);
  std::memcpy(rp.M, t.M, sizeof rp.M);
  std::memcpy(rp.E, t.E, sizeof rp.E);
  std::memcpy(rp.Alpha, t.Alpha, sizeof rp.Alpha);
  std::memcpy(rp.Gamma, t.Gamma, sizeof rp.Gamma);
  rp.ELO = t.ELO;
  for (int i = 0; i < 6; ++i) rp.NewF[i] = t.NewF[i];
  if (ic[3] == 0) rp.max_no_steps = (family == FAM_ADJ) ? 200000 - 1 : 200000;   // bufsize-1 (ADJ :416) / 200000 (FWD :285)
  else if (ic[3] > 0) rp.max_no_steps = ic[3];   // ADJ :417 tests an uninitialised variable; intent restated
  else return -1;
  if (family == FAM_ADJ) {
    if (!(ic[5] == 0 || (ic[5] >= 1 && ic[5] <= 5))) return -2;
    if (!(ic[6] == 0 || (ic[6] >= 1 && ic[6] <= 4))) return -9;
  }
  rp.roundoff = DBL_EPSILON * 0.5;               // DLAMCH('E')
  if (rc[0] == 0.0) rp.hmin = 0.0; else if (rc[0] > 0.0) rp.hmin = rc[0]; else return -3;
  const double span = std::fabs(tout - tin);
  if (rc[1] == 0.0) rp.hmax = span; else if (rc[1] > 0.0) rp.hmax = std::fmin(std::fabs(rc[1]), span); else return -3;
  rp.deltamin_start = sp ? (double)1.0E-5f : 1.0e-5;
  rp.deltamin_fd = sp ? (double)1.0E-6f : 1.0e-6;
  rp.tenth = sp ? (double)0.1f : 0.1;
  if (rc[2] == 0.0) rp.hstart = std::fmax(rp.hmin, rp.deltamin_start);
  else if (rc[2] > 0.0) rp.hstart = std::fmin(std::fabs(rc[2]), span);
  else return -3;
  if (rc[3] == 0.0) rp.facmin = sp ? (double)0.2f : 0.2; else if (rc[3] > 0.0) rp.facmin = rc[3]; else return -4;
  if (rc[4] == 0.0) rp.facmax = 6.0; else if (rc[4] > 0.0) rp.facmax = rc[4]; else return -4;
  if (rc[5] == 0.0) rp.facrej = sp ? (double)0.1f : 0.1; else if (rc[5] > 0.0) rp.facrej = rc[5]; else return -4;
  if (rc[6] == 0.0) rp.facsafe = sp ? (double)0.9f : 0.9; else if (rc[6] > 0.0) rp.facsafe = rc[6]; else return -4;
  const int uplim = rp.vector_tol ? N : 1;
  for (int i = 0; i < uplim; ++i)
    if (atol[i] <= 0.0 || rtol[i] <= 10.0 * rp.roundoff || rtol[i] >= 1.0) return -5;
  rp.tstart = tin;
  rp.tend = tout;
  return 1;
}

}  // namespace fatode
