// Explicit Runge-Kutta discrete adjoint on the shallow-water ensemble: ADJ/ERK_ADJ INTEGRATE_ADJ (no parameters: ERKADJ2).
//
// Replaces ERK_FwdInt (ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:1600-1740: ERK_Integrator plus the checkpoint of every accepted
// step, ERK_Push :757-776), ADJINIT of the driver (EXAMPLES/swe/SWE_ERK_ADJ/swe2D_erk_adj_dr.F90:31-43: Lambda(k,k) = 1) and
// ERK_DadjInt (Lambda-only twin of :608-700): per popped step, stages i = S..1
//   U_i = JAC(T + c_i h, Z_i)^T ( h sum_{j>i} a_ji U_j + h b_i Lambda ),   then Lambda += sum_i U_i
// with the JAC callback of the example (swe2D_erk_adj_dr.F90:2-13 -> compute_JacF, swe2D_upwind.F90:483-2023), which the
// reference evaluates as a dense 4800 x 4800 matrix and applies with DGEMV('T').
//
// Mapping: one ensemble member per CTA does BOTH sweeps (320 threads; members from an atomic queue).  Forward: the kernel of
// erk_swe.cuh (state and stage vector in shared memory, stage derivatives in an L2-resident per-CTA scratch) plus the stage
// states Z_1..Z_S of the running attempt written to the CTA's tape slot; an accepted step keeps them.  Backward, per stage:
//   phase 1  one thread per matrix ROW: the 27 structural entries of the row from the generated formulas (gen/swe_jac_gen.h, the
//            reference's statements in the reference's order; eigenvalue signs recomputed as compute_F does) -> JV[entry][row],
//            once per stage as upstream (one JAC call), shared by all adjoint columns;
//   phase 2  per adjoint column: the operand vector in shared memory, then one thread per matrix COLUMN gathers its 27 products in
//            a fixed order (rows ascending in the interior, as DGEMV('T') sums them) -- no atomics, bit-reproducible.
// The dense matrix is never formed: 1 MB of entries per stage instead of 184 MB.  The forward sweep is bit-identical to the oracle
// (state, ISTATUS, RSTATUS); the backward sweep performs the reference's operations with the additions of a product ordered by
// row except at the periodic seams, and pow(x, 1.5) from CUDA's libm: Lambda agrees to rounding (1e-12 observed, contract 1e-9).
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/fatode_b200.h"
#include "units_internal.h"
#include "erk_swe.cuh"

namespace fatode {

constexpr int IERR_ERK_TAPE = -30;           // internal: more accepted steps than the CTA's tape slot holds (the host retries)

struct ErkAdjArgs {
  long nmembers;
  const int* ids;              // members of this launch (nullptr: 0..nmembers-1)
  double* y;                   // [member][N] in/out
  double* lambda;              // [member][nadj][N]
  const double* atol; const double* rtol;
  int* istatus; double* rstatus; int* ierr;
  unsigned long long* queue;
  double* scratch;             // per CTA: see the offsets in the kernel
  size_t scratch_stride;       // doubles per CTA
  int nadj, cap;               // adjoint columns; tape records per CTA
};

__global__ void __launch_bounds__(ERK_THREADS, 2)
k_erk_adj_swe(ErkParams p, ErkAdjArgs a) {
  extern __shared__ __align__(16) double erk_smem[];
  double* const Ys = erk_smem;
  double* const TMP = erk_smem + SWE_N;
  double* const red = erk_smem + 2 * SWE_N;
  const int tid = threadIdx.x;
  const int S = p.S, nadj = a.nadj, cap = a.cap;
  double* const base = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
  double* const K = base;                                         // [S][N]
  double* const JV = K + (size_t)S * SWE_N;                       // [27][N]
  double* const LAM = JV + (size_t)SWE_NNZR * SWE_N;              // [nadj][N]
  double* const U = LAM + (size_t)nadj * SWE_N;                   // [nadj][S][N]
  double* const HDR = U + (size_t)nadj * S * SWE_N;               // [cap][2]  (T, H)
  double* const TZ = HDR + (size_t)2 * cap;                       // [cap][S][N]
  const double Roundoff = p.roundoff;
  const double Tdirection = (p.tout - p.tin) >= 0.0 ? 1.0 : -1.0;
  const double dx = p.c6x / 6.0;
  for (;;) {
    __syncthreads();
    if (tid == 0) reinterpret_cast<unsigned long long*>(red)[1] = atomicAdd(a.queue, 1ull);
    __syncthreads();
    const unsigned long long slot = reinterpret_cast<unsigned long long*>(red)[1];
    if (slot >= (unsigned long long)a.nmembers) break;
    const long member = a.ids ? (long)a.ids[slot] : (long)slot;
    double* const yio = a.y + (size_t)member * SWE_N;
    for (int v = tid; v < SWE_N; v += ERK_THREADS) Ys[v] = yio[v];
    int nfun = 0, nstp = 0, nacc = 0, nrej = 0, ierr = 1;
    double T = p.tin, texit = 0.0, hexit = 0.0, hnew_out = 0.0;
    double H = fmax(fabs(p.hmin), fabs(p.hstart));
    if (fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = fmin(fabs(H), p.hmax);
    H = copysign(H, Tdirection);
    bool Reject = false, FirstStep = true;
    __syncthreads();
    // ---------------- forward sweep (ERK_FwdInt)
    while ((p.tout - T) * Tdirection - Roundoff > 0.0) {
      if (nstp > p.max_no_steps) { ierr = -6; break; }
      if ((T + 0.1 * H == T) || (fabs(H) <= Roundoff)) { ierr = -7; break; }
      if (nacc >= cap) { ierr = IERR_ERK_TAPE; break; }
      double* const rec = TZ + (size_t)nacc * S * SWE_N;
#pragma unroll 1
      for (int is = 0; is < S; ++is) {
        {
          double t[ERK_EPT];
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) t[q] = 0.0;
          for (int j = 0; j < is; ++j) {
            const double c = H * p.A[is][j];
            if (c != 0.0) {
              const double* Kj = K + (size_t)j * SWE_N + tid;
#pragma unroll
              for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Kj[q * ERK_THREADS];
            }
          }
          double* const Zi = rec + (size_t)is * SWE_N;
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) {
            const double z = t[q] + Ys[tid + q * ERK_THREADS];
            TMP[tid + q * ERK_THREADS] = z;
            Zi[tid + q * ERK_THREADS] = z;                          // the stage state Z_i of the checkpoint (:1651-1660, :1695)
          }
        }
        __syncthreads();
        swe_fun_cells<false>(TMP, p, K + (size_t)is * SWE_N, nullptr, 0.0);
        nfun++;
        __syncthreads();
      }
      nstp++;
      {
        double t[ERK_EPT];
#pragma unroll
        for (int q = 0; q < ERK_EPT; ++q) t[q] = 0.0;
        for (int i = 0; i < S; ++i)
          if (p.E[i] != 0.0) {
            const double c = H * p.E[i];
            if (c != 0.0) {
              const double* Ki = K + (size_t)i * SWE_N + tid;
#pragma unroll
              for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Ki[q * ERK_THREADS];
            }
          }
#pragma unroll
        for (int q = 0; q < ERK_EPT; ++q) {
          const int v = tid + q * ERK_THREADS;
          const int vt = p.vector_tol ? v : 0;
          const double scal = 1.0 / (a.atol[vt] + a.rtol[vt] * fabs(Ys[v]));
          const double w = t[q] * scal;
          TMP[v] = w * w;
        }
      }
      __syncthreads();
      if (tid == 0) {
        double e = 0.0;
#pragma unroll 8
        for (int v = 0; v < SWE_N; ++v) e = e + TMP[v];
        red[0] = fmax(sqrt(e / (double)SWE_N), 1.0e-10);
      }
      __syncthreads();
      const double Err = red[0];
      double Fac = p.facsafe * fpm::pow_neg_inv(Err, p.ELO);
      Fac = fmax(p.facmin, fmin(p.facmax, Fac));
      double Hnew = H * Fac;
      if (Err < 1.0) {
        FirstStep = false;
        nacc++;
        if (nacc > p.max_no_steps) { ierr = -6; break; }   // ERK_Push: "buffer overflow", the reference STOPs (:766-770)
        if (tid == 0) { HDR[2 * (nacc - 1)] = T; HDR[2 * (nacc - 1) + 1] = H; }
        T = T + H;
        {
          double t[ERK_EPT];
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) t[q] = Ys[tid + q * ERK_THREADS];
          for (int i = 0; i < S; ++i)
            if (p.B[i] != 0.0) {
              const double c = H * p.B[i];
              if (c != 0.0) {
                const double* Ki = K + (size_t)i * SWE_N + tid;
#pragma unroll
                for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Ki[q * ERK_THREADS];
              }
            }
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) Ys[tid + q * ERK_THREADS] = t[q];
        }
        Hnew = Tdirection * fmin(fabs(Hnew), p.hmax);
        texit = T; hexit = H; hnew_out = Hnew;
        if (Reject) Hnew = Tdirection * fmin(fabs(Hnew), fabs(H));
        Reject = false;
        if ((T + Hnew / p.qmin - p.tout) * Tdirection > 0.0) H = p.tout - T;
        else H = Hnew;
      } else {
        if (FirstStep || Reject) H = p.facrej * H; else H = Hnew;
        Reject = true;
        if (nacc >= 1) nrej++;
      }
      __syncthreads();
    }
    __syncthreads();
    if (ierr == IERR_ERK_TAPE) {   // the member keeps its initial state and is integrated again with a larger tape slot
      if (tid == 0) a.ierr[member] = IERR_ERK_TAPE;
      continue;
    }
    for (int v = tid; v < SWE_N; v += ERK_THREADS) yio[v] = Ys[v];
    const bool push_overflow = (ierr == -6 && nacc > p.max_no_steps);
    int njac = 0;
    if (!push_overflow) {
      // ---------------- ADJINIT and the backward sweep (ERK_DadjInt); a failed forward sweep still gets both (:1563-1576)
      for (int m = 0; m < nadj; ++m)
        for (int v = tid; v < SWE_N; v += ERK_THREADS) LAM[(size_t)m * SWE_N + v] = (v == m) ? 1.0 : 0.0;
      __syncthreads();
      for (int r = nacc - 1; r >= 0; --r) {
        const double Hs = HDR[2 * r + 1];
        const double* const rec = TZ + (size_t)r * S * SWE_N;
        nstp++;
#pragma unroll 1
        for (int is = S - 1; is >= 0; --is) {
          for (int v = tid; v < SWE_N; v += ERK_THREADS) TMP[v] = rec[(size_t)is * SWE_N + v];
          __syncthreads();
          swe_jac_rows(TMP, JV, dx);                                  // JAC(T + c_i H, Z_i)
          njac++;
          __syncthreads();
#pragma unroll 1
          for (int m = 0; m < nadj; ++m) {
            {   // x = sum_{j>i} (H a_ji) U_j + (H b_i) Lambda   (DAXPYs in this order; a zero factor returns at once)
              double t[ERK_EPT];
#pragma unroll
              for (int q = 0; q < ERK_EPT; ++q) t[q] = 0.0;
              for (int j = is + 1; j < S; ++j) {
                const double c = Hs * p.A[j][is];
                if (c != 0.0) {
                  const double* Uj = U + ((size_t)m * S + j) * SWE_N + tid;
#pragma unroll
                  for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Uj[q * ERK_THREADS];
                }
              }
              {
                const double c = Hs * p.B[is];
                if (c != 0.0) {
                  const double* Lm = LAM + (size_t)m * SWE_N + tid;
#pragma unroll
                  for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Lm[q * ERK_THREADS];
                }
              }
#pragma unroll
              for (int q = 0; q < ERK_EPT; ++q) Ys[tid + q * ERK_THREADS] = t[q];
            }
            __syncthreads();
            swe_jac_t_gather(JV, Ys, U + ((size_t)m * S + is) * SWE_N);   // U_i = JAC^T x  (DGEMV 'T')
            __syncthreads();
          }
        }
        for (int m = 0; m < nadj; ++m) {
          double t[ERK_EPT];
          double* Lm = LAM + (size_t)m * SWE_N + tid;
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) t[q] = Lm[q * ERK_THREADS];
          for (int i = 0; i < S; ++i) {
            const double* Ui = U + ((size_t)m * S + i) * SWE_N + tid;
#pragma unroll
            for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + Ui[q * ERK_THREADS];
          }
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) Lm[q * ERK_THREADS] = t[q];
        }
        __syncthreads();
      }
      double* lo = a.lambda + (size_t)member * nadj * SWE_N;
      for (int m = 0; m < nadj; ++m)
        for (int v = tid; v < SWE_N; v += ERK_THREADS) lo[(size_t)m * SWE_N + v] = LAM[(size_t)m * SWE_N + v];
      ierr = 1;                                                       // the return code is the backward sweep's
    }
    if (tid == 0) {
      a.ierr[member] = ierr;
      int* ist = a.istatus + member * 20;
      for (int i = 0; i < 20; ++i) ist[i] = 0;
      ist[0] = nfun; ist[1] = njac; ist[2] = nstp; ist[3] = nacc; ist[4] = nrej;
      double* rst = a.rstatus + member * 20;
      for (int i = 0; i < 20; ++i) rst[i] = 0.0;
      rst[0] = texit; rst[1] = hexit; rst[2] = hnew_out;
    }
  }
}

static size_t erk_adj_stride(int S, int nadj, int cap) {
  return (size_t)S * SWE_N + (size_t)SWE_NNZR * SWE_N + (size_t)nadj * SWE_N + (size_t)nadj * S * SWE_N + (size_t)2 * cap + (size_t)cap * S * SWE_N;
}

// members 0..nmembers-1 of the device arrays; budget = bytes the tape / scratch may take
int erk_adj_swe_device(const ErkParams& ep, long nmembers, int nadj, double* d_y, double* d_lambda, const double* d_atol,
                       const double* d_rtol, int* d_istatus, double* d_rstatus, int* d_ierr, int tape_slot, size_t budget,
                       cudaStream_t st, UnitTiming& tm, long& retried, std::string& err) {
  retried = 0;
  if (nmembers <= 0) return 0;
  int dev = 0, sms = 0, per_sm = 0;
  UNIT_TRY(cudaGetDevice(&dev));
  UNIT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  UNIT_TRY(cudaFuncSetAttribute(k_erk_adj_swe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ERK_SMEM));
  UNIT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_erk_adj_swe, ERK_THREADS, ERK_SMEM));
  unsigned long long* d_queue = nullptr;
  UNIT_TRY(cudaMalloc(&d_queue, sizeof(unsigned long long)));
  struct Free { void* p; ~Free() { if (p) cudaFree(p); } } gq{d_queue};
  auto launch = [&](long n, const int* d_ids, int cap) -> int {
    const size_t stride = erk_adj_stride(ep.S, nadj, cap);
    long blocks = std::min<long>(n, (long)sms * std::max(per_sm, 1));
    blocks = std::min<long>(blocks, (long)(budget / (stride * sizeof(double))));
    if (blocks < 1) { err = "ERK_ADJ: the tape of one member (Max_no_steps checkpoints) exceeds the tape budget"; return -104 /*FATODE_ENOMEM*/; }
    double* d_scratch = nullptr;
    UNIT_TRY(cudaMalloc(&d_scratch, stride * sizeof(double) * (size_t)blocks));
    Free gs{d_scratch};
    UNIT_TRY(cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), st));
    ErkAdjArgs a{n, d_ids, d_y, d_lambda, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr, d_queue, d_scratch, stride, nadj, cap};
    cudaEvent_t e0, e1;
    UNIT_TRY(cudaEventCreate(&e0));
    UNIT_TRY(cudaEventCreate(&e1));
    UNIT_TRY(cudaEventRecord(e0, st));
    k_erk_adj_swe<<<(unsigned)blocks, ERK_THREADS, ERK_SMEM, st>>>(ep, a);
    cudaError_t le = cudaGetLastError();
    cudaEventRecord(e1, st);
    cudaError_t se = cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    UNIT_TRY(le);
    UNIT_TRY(se);
    tm.ms += ms;
    tm.launches += 1;
    return 0;
  };
  // Max_no_steps + 1 records: the attempt that would be accepted beyond the reference's buffer still needs room for its stages
  const int cap_full = ep.max_no_steps + 1;
  const int cap1 = std::max(1, std::min(cap_full, tape_slot > 0 ? tape_slot : 128));
  if (int e = launch(nmembers, nullptr, cap1)) return e;
  if (cap1 < cap_full) {   // members with more accepted steps than the slot: again with room for Max_no_steps checkpoints
    std::vector<int> h_ierr((size_t)nmembers), again;
    UNIT_TRY(cudaMemcpyAsync(h_ierr.data(), d_ierr, sizeof(int) * nmembers, cudaMemcpyDeviceToHost, st));
    UNIT_TRY(cudaStreamSynchronize(st));
    for (long m = 0; m < nmembers; ++m)
      if (h_ierr[(size_t)m] == IERR_ERK_TAPE) again.push_back((int)m);
    if (!again.empty()) {
      int* d_ids = nullptr;
      UNIT_TRY(cudaMalloc(&d_ids, sizeof(int) * again.size()));
      Free gi{d_ids};
      UNIT_TRY(cudaMemcpyAsync(d_ids, again.data(), sizeof(int) * again.size(), cudaMemcpyHostToDevice, st));
      if (int e = launch((long)again.size(), d_ids, cap_full)) return e;
      retried = (long)again.size();
    }
  }
  return 0;
}

}  // namespace fatode
