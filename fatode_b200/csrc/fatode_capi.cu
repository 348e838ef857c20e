// fatode_b200 — C-ABI entry points (include/fatode_b200.h), kernels and the host-side wave scheduler.
//
// Execution model for INTEGRATE_ADJ over a batch:
//   cells are processed in waves sized to the trajectory-tape budget in HBM.  Per wave:
//     1. k_ros_fwd<TAPE=false>  forward sweep, one cell per warp -> y(Tend), ISTATUS, RSTATUS, IERR,
//                               accepted-step count per cell
//     2. k_scan_offsets         exclusive prefix sum of the counts -> exact tape offsets
//     3. k_ros_fwd<TAPE=true>   the same sweep again (bit-identical by construction), writing
//                               (T,H,Ystage,K) per accepted step at the cell's offset
//     4. k_ros_adj2             backward sweep, persistent producer/consumer warp pairs pulling cells from an atomic queue
//   then one deterministic two-stage reduction of Mu over cells (k_mu_reduce*).
// There is no CPU path: every failure to reach the GPU is reported as an error.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/fatode_b200.h"
#include "fatode_common.cuh"
#include "ros_options.h"
#include "model_cbm4.cuh"
#include "ros_warp_fwd.cuh"
#include "ros_warp_adj.cuh"
#include "ros_warp_adj2.cuh"
#include "ros_cell_fwd.cuh"
#include "ros_cell_fwd_tm.cuh"
#include "ros_cell_adj.cuh"
#include "ros_cell_adj_pair.cuh"
#include "model_small_cell.cuh"
#include "sdirk_cell_fwd.cuh"
#include "sdirk_cell_tlm.cuh"
#include "erk_swe.cuh"
#include "sdirk_cell_adj.cuh"
#include "rk_cell_fwd.cuh"
#include "model_cbm4_dense_cell.cuh"
#include "rk_cell_adj.cuh"
#include "sdirk_dense_sens.cuh"
#include "rk_dense_adj.cuh"
#include "rk_big_adj.cuh"
#include "units_internal.h"

using namespace fatode;

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local std::string g_err;
static thread_local fatode_stats g_stats;
static uint64_t g_tape_budget = 0;
static int64_t g_wave_cells = 0;
static int g_path = 0;                    // 0 = per-thread fast path with general fallback, 1 = general kernels only
static int g_pipeline = 0;                // 0 = phases back to back (default), 1/2 = forward of the next wave overlapped
static int g_debug_irregular_every = 0;   // test hook: every k-th cell is pushed through the fallback

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      int code_ = (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? FATODE_ENODEVICE : FATODE_ECUDA; \
      return fail(code_, std::string(#x) + ": " + cudaGetErrorString(e_));                      \
    }                                                                                            \
  } while (0)

extern "C" const char* fatode_last_error(void) { return g_err.c_str(); }
extern "C" const char* fatode_version(void) { return "fatode_b200 0.1 (sm_100a)"; }
extern "C" int fatode_set_tape_budget(uint64_t bytes) { g_tape_budget = bytes; return 0; }
extern "C" int fatode_set_wave_cells(int64_t cells) { g_wave_cells = cells; return 0; }
extern "C" int fatode_set_path(int mode) { g_path = mode; return 0; }
extern "C" int fatode_set_pipeline(int mode) { g_pipeline = mode; return 0; }
extern "C" int fatode_debug_irregular_every(int k) { g_debug_irregular_every = k; return 0; }
extern "C" int fatode_get_last_stats(fatode_stats* out) { if (!out) return FATODE_EINVAL; *out = g_stats; return 0; }

extern "C" void* fatode_alloc_pinned(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
  return p;
}
extern "C" void fatode_free_pinned(void* p) { if (p) cudaFreeHost(p); }

extern "C" int fatode_model_dims(int model, int* nvar, int* np) {
  switch (model) {
    case FATODE_MODEL_ROBERTS: *nvar = 3; *np = 3; return 0;
    case FATODE_MODEL_SMALL_STRATO: *nvar = 5; *np = 0; return 0;
    case FATODE_MODEL_CBM4: *nvar = 32; *np = 29; return 0;
    case FATODE_MODEL_SWE: *nvar = SWE_N; *np = 0; return 0;
  }
  return fail(FATODE_EINVAL, "unknown model id");
}

extern "C" int fatode_model_initial_state(int model, double* y) {
  if (!y) return fail(FATODE_EINVAL, "y is NULL");
  switch (model) {
    case FATODE_MODEL_ROBERTS: y[0] = 1; y[1] = 0; y[2] = 0; return 0;                    // roberts_ros_adj_dr.F90:194-196
    case FATODE_MODEL_SMALL_STRATO:                                                         // small_strato_dr.F90:50-54
      y[0] = (double)9.906E+01f; y[1] = (double)6.624E+08f; y[2] = (double)5.326E+11f;
      y[3] = (double)8.725E+08f; y[4] = (double)2.240E+08f; return 0;
    case FATODE_MODEL_CBM4: {                                                               // cbm4_ros_adj_dr.F90:233-264
      static const double v[32] = {
          3.652686375061191e-02, 3.470772014720235e+11, 2.558736854382598e+03, 3.353179955426548e-23,
          5.287698343141944e-20, 1.307511063007438e+07, 0.0, 1.985180650951722e+01,
          3.839565210778183e+08, 2.675425503948970e+08, 1.593461800730453e-24, 1.191924698309334e+12,
          4.894101892069533e-05, 5.462490511678749e-21, 1.943997577140665e-23, 2.329863421033919e+12,
          2.106517721289787e-30, 5.518957651363696e+08, 2.633932009302457e-21, 8.118206770655152e+03,
          3.000632019118241e+10, 0.0, 0.0, 2.687919546821573e+02,
          2.061185883675469e+12, 1.561906017211229e+10, 3.671987743784878e+07, 9.888485531199334e+08,
          1.248793414327174e+04, 3.357869003391008e+07, 2.727638922137502e+09, 6.534644475169904e+00};
      std::memcpy(y, v, sizeof v);
      return 0;
    }
  }
  if (model == FATODE_MODEL_SWE) return fatode_swe_initial_state(30.0, 2.0, 2.0, y);   // swe2D_erk_dr.F90:100-103
  return fail(FATODE_EINVAL, "unknown model id");
}

// initializeGaussianGrid(A) (EXAMPLES/swe/SWE_ERK/swe2D_upwind.F90:86-126) + constant velocities + grid2vec (:267-289):
// y[k * 1600 + (i - 3) * 40 + (j - 3)] = U(k, i, j), h = 100 + A exp(-x_j^2 - y_i^2)
extern "C" int fatode_swe_initial_state(double amplitude, double u0, double v0, double* y) {
  if (!y) return fail(FATODE_EINVAL, "y is NULL");
  const double dx = (3.0 - (-3.0)) / (SWE_MX - 1.0);
  for (int i = 3; i <= SWE_MX + 2; ++i)
    for (int j = 3; j <= SWE_MX + 2; ++j) {
      const double xj = -3.0 + (j - 2) * dx, yi = -3.0 + (i - 2) * dx;
      const int c = (i - 3) * SWE_MX + (j - 3);
      y[c] = amplitude * std::exp(-xj * xj - yi * yi) + 100.0;
      y[SWE_CELLS + c] = u0;
      y[2 * SWE_CELLS + c] = v0;
    }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernels

constexpr int FWD_WARPS = 4;

template <class Model>
__host__ __device__ constexpr int fwd_ws_doubles() { return FwdSmem<Model>::DOUBLES; }


// MODE 0: no tape, writes y/ISTATUS/RSTATUS/IERR/nacc.   MODE 1: replay with tape at exact offsets (tape_off), no
// status.   MODE 2: single pass, status + tape into a fixed slab of `slab_cap` entries per cell.
template <class Model, int MODE>
__global__ void __launch_bounds__(FWD_WARPS * 32, 3)
k_ros_fwd(RosParams rp, typename Model::Const mc, long ncells, const double* __restrict__ y_in, double* __restrict__ y_out,
          const double* __restrict__ atol, const double* __restrict__ rtol, int* __restrict__ istatus,
          double* __restrict__ rstatus, int* __restrict__ ierr, int* __restrict__ nacc,
          const long long* __restrict__ tape_off, double* __restrict__ tape, long slab_cap, int* __restrict__ n_overflow) {
  extern __shared__ __align__(16) double smem[];
  constexpr bool TAPE = (MODE != 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long cell = (long)blockIdx.x * FWD_WARPS + warp;
  if (cell >= ncells) return;
  if (MODE == 1 && ierr[cell] != 1) return;
  double* ws = smem + (size_t)warp * fwd_ws_doubles<Model>();
  double y = y_in[cell * 32 + lane];
  const int ti = rp.vector_tol ? lane : 0;
  FwdResult res;
  double* tb = nullptr;
  long cap = 0;
  const int NE = tape_entry_doubles(rp.S, 32);
  if (MODE == 1) {
    tb = tape + tape_off[cell] * NE;
    cap = (long)(tape_off[cell + 1] - tape_off[cell]);
  } else if (MODE == 2) {
    tb = tape + cell * slab_cap * NE;
    cap = slab_cap;
  }
  ros_forward_warp<Model, TAPE>(rp, mc, ws, y, atol[ti], rtol[ti], tb, cap, res, lane);
  if (MODE != 1) {
    const bool overflow = (MODE == 2) && res.ierr == -6 && res.nacc > slab_cap;   // ran out of slab, not of steps
    y_out[cell * 32 + lane] = y;
    if (lane < 20) {
      int v = 0;
      if (lane == Nfun) v = res.nfun; else if (lane == Njac) v = res.njac; else if (lane == Nstp) v = res.nstp;
      else if (lane == Nacc) v = res.nacc; else if (lane == Nrej) v = res.nrej; else if (lane == Ndec) v = res.ndec;
      else if (lane == Nsol) v = res.nsol; else if (lane == Nsng) v = res.nsng;
      istatus[cell * 20 + lane] = v;
    }
    if (lane < 20) {
      double r = 0.0;
      if (lane == Ntexit) r = res.texit; else if (lane == Nhexit) r = res.hexit; else if (lane == Nhnew) r = res.hnew;
      rstatus[cell * 20 + lane] = r;
    }
    if (lane == 0) {
      ierr[cell] = overflow ? IERR_TAPE_OVERFLOW : res.ierr;
      nacc[cell] = res.nacc;
      if (overflow) atomicAdd(n_overflow, 1);
    }
  }
}

// exclusive scan of nacc (only cells with ierr == 1 get tape space); single block
__global__ void k_scan_offsets(long ncells, const int* __restrict__ nacc, const int* __restrict__ ierr,
                               long long* __restrict__ off) {
  __shared__ long long part[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long base = 0; base < ncells; base += 1024) {
    long i = base + threadIdx.x;
    long long v = (i < ncells && ierr[i] == 1) ? nacc[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      long long t = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < ncells) off[i] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) off[ncells] = carry;
}

// Backward kernel: producer/consumer warp pairs, TMEM-resident running sums (ros_warp_adj2.cuh)
template <class Model>
__global__ void __launch_bounds__(ADJ2_THREADS, 1)
k_ros_adj2(RosParams rp, typename Model::Const mc, long ncells, int nadj, int with_mu, const double* __restrict__ tape,
           const long long* __restrict__ tape_off, const int* __restrict__ ierr, const double* __restrict__ y_final,
           double* __restrict__ lambda, double* __restrict__ mu, int* __restrict__ istatus,
           unsigned long long* __restrict__ queue, long slab_cap, const int* __restrict__ nacc) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint32_t tmem_base_s;
  typedef Adj2Smem<Model> L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp & 3;
  const bool producer = warp >= 4;
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)(slot * 32) << 16);
  double* sm = smem + (size_t)slot * L::DOUBLES;
  volatile long long* ctrl = reinterpret_cast<volatile long long*>(sm + L::OFF_CTRL);
  const int S = rp.S;
  const int NE = tape_entry_doubles(S, 32);
  if (producer) Model::init(sm + L::OFF_MODEL, mc, lane);
  for (;;) {
    if (producer && lane == 0) {
      long long c = (long long)atomicAdd(queue, 1ull);
      while (c < ncells && ierr[c] != 1) c = (long long)atomicAdd(queue, 1ull);
      ctrl[0] = (c < ncells) ? c : -1;
    }
    pair_barrier(slot);
    const long long c = ctrl[0];
    if (c < 0) break;
    const long long off = slab_cap > 0 ? c * slab_cap : tape_off[c];
    const int nsteps = slab_cap > 0 ? nacc[c] : (int)(tape_off[c + 1] - tape_off[c]);
    const double* ent0 = tape + off * NE;
    const int n_items = nsteps * S;
    if (!producer) {   // ADJINIT (model routine): lane = adjoint column m
      double lam[32], m0[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) lam[i] = Model::adjinit_lambda(i, lane, y_final[c * 32 + i]);
#pragma unroll
      for (int p = 0; p < 32; ++p) m0[p] = 0.0;
      tmem::st32(tm + TM_LAM, lam);
      tmem::st32(tm + TM_MU, m0);
      tmem::wait_st();
    } else if (n_items > 0) {
      adj2_produce<Model>(rp, sm, ent0 + (long)(nsteps - 1) * NE, S - 1, 0, 0, with_mu != 0, lane);
    }
    pair_barrier(slot);
    for (int t = 0; t < n_items; ++t) {
      if (!producer) {
        const int si = t / S;
        adj2_consume<Model>(rp, sm, S - 1 - (t - si * S), si & 1, t & 1, with_mu != 0, tm);
      } else if (t + 1 < n_items) {
        const int t1 = t + 1;
        const int si = t1 / S;
        adj2_produce<Model>(rp, sm, ent0 + (long)(nsteps - 1 - si) * NE, S - 1 - (t1 - si * S), si & 1, t1 & 1, with_mu != 0, lane);
      }
      pair_barrier(slot);
    }
    if (!producer) {   // results: column m of Lambda / Mu is this lane's private vector
      tmem::wait_st();
      {
        double lam[32];
        tmem::ld32(tm + TM_LAM, lam);   // .sync.aligned: executed by the whole warp, outside any lane-dependent branch
        if (lane < nadj) {
          double* dst = lambda + ((long)c * nadj + lane) * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(lam[i], lam[i + 1]);
        }
      }
      if (with_mu) {
        double m0[32];
        tmem::ld32(tm + TM_MU, m0);
        if (lane < nadj) {
          double* dst = mu + ((long)c * nadj + lane) * Model::NP;
#pragma unroll
          for (int p = 0; p < Model::NP; ++p) dst[p] = m0[p];
        }
      }
      if (lane == 0) {   // counters of the backward sweep (ADJ :1257-1267, :1303, :1312, :1381)
        int* is_ = istatus + c * 20;
        is_[Nstp] += nsteps;
        is_[Njac] += nsteps * (2 + S + (rp.autonomous ? 0 : 1));
        is_[Ndec] += nsteps;
        is_[Nsol] += nsteps * S * nadj;
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

// deterministic sum of mu over cells: stage 1 partial sums over fixed cell blocks, stage 2 over blocks
__global__ void k_mu_reduce1(long ncells, int len, const double* __restrict__ mu, const int* __restrict__ ierr,
                             double* __restrict__ partial, int cells_per_block) {
  const long c0 = (long)blockIdx.x * cells_per_block;
  const long c1 = min(ncells, c0 + cells_per_block);
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    double s = 0.0;
    for (long c = c0; c < c1; ++c)
      if (ierr[c] == 1) s += mu[c * len + j];
    partial[(long)blockIdx.x * len + j] = s;
  }
}
__global__ void k_mu_reduce2(int nblocks, int len, const double* __restrict__ partial, double* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(long)b * len + j];
  out[j] = s;
}

// FP64 peak probe: 8 independent DFMA chains per thread
__global__ void k_dfma_probe(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" double fatode_measure_fp64_peak(int iters) {
  if (iters <= 0) iters = 20000;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { g_err = "no CUDA device"; return -1.0; }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256;
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) { g_err = "cudaMalloc failed"; return -1.0; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_probe<<<blocks, threads>>>(d, iters / 10);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma_probe<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) { g_err = "dfma probe failed"; return -1.0; }
  const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  return flops / (best * 1e-3) / 1e12;
}

// ------------------------------------------------------------------------------------------------
// host side

static int cbm4_constants(const double* params, Cbm4Const& mc) {
  std::memset(&mc, 0, sizeof mc);
  mc.temp = params ? params[0] : 298.0;                    // cbm4_ros_adj_dr.F90:268
  mc.fix = params ? params[1] : (1.25e+8) * 2.55e10;       // :178,202
  mc.emis = params ? params[2] : 2.0e3;                    // :96-97
  for (int i = 0; i < 41; ++i)                             // ARR2 (cbm4_Parameters.F90:2389-2392)
    mc.rc_base[cbm4_dev::ARR_IDX[i]] = cbm4_dev::ARR_A0[i] * std::exp(cbm4_dev::ARR_B0[i] / mc.temp);
  return 0;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Grow-only device workspace cached across calls (cudaMalloc/cudaFree of a 100 GB tape costs ~100 ms).
struct CachedBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
static thread_local CachedBuf g_tape, g_y0, g_nacc, g_off, g_queue, g_part;
static thread_local CachedBuf g_ids, g_wnacc, g_wlast, g_prev, g_owner, g_order, g_wierr;   // per-thread path
static void release_pipeline_workspace();
static void release_rk_workspace();
static void release_sdirk_workspace();
static void release_erk_workspace();
extern "C" int fatode_release_workspace(void) {
  g_tape.release(); g_y0.release(); g_nacc.release(); g_off.release(); g_queue.release(); g_part.release();
  g_ids.release(); g_wnacc.release(); g_wlast.release(); g_prev.release(); g_owner.release(); g_order.release(); g_wierr.release();
  release_pipeline_workspace();
  release_rk_workspace();
  release_sdirk_workspace();
  release_erk_workspace();
  return 0;
}

struct PhaseTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t s;
  explicit PhaseTimer(cudaStream_t st) : s(st) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  ~PhaseTimer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  void start() { cudaEventRecord(e0, s); }
  double stop() { cudaEventRecord(e1, s); cudaEventSynchronize(e1); float ms = 0; cudaEventElapsedTime(&ms, e0, e1); return ms; }
};

extern "C" int fatode_release_workspace(void);
static thread_local int g_ws_device = -1;   // device the cached workspace of this host thread lives on
static int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(FATODE_ENODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != g_ws_device) {   // the caller switched GPUs: drop the other GPU's buffers
    if (g_ws_device >= 0) {
      cudaSetDevice(g_ws_device);
      fatode_release_workspace();
      cudaSetDevice(dev);
    }
    g_ws_device = dev;
  }
  return 0;
}

// GENERAL PATH: core of INTEGRATE / INTEGRATE_ADJ on device-resident arrays with the warp-per-cell kernels
// (full partial pivoting, singular-matrix retries).  Used for cells the per-thread path hands back and when
// fatode_set_path(1) asks for it.
template <class Model>
static int ros_device_general(const RosParams& rp, const typename Model::Const& mc, int64_t ncells, int nadj, bool with_mu,
                          bool do_adjoint, double* d_y, double* d_lambda, double* d_mu, const double* d_atol,
                          const double* d_rtol, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st) {
  const int NE = tape_entry_doubles(rp.S, 32);
  const size_t entry_bytes = sizeof(double) * NE;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t fwd_smem = sizeof(double) * fwd_ws_doubles<Model>() * FWD_WARPS;
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd<Model, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd<Model, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd<Model, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));

  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = g_tape_budget ? (size_t)g_tape_budget : (size_t)(0.6 * (double)free_b);
  budget = std::min(budget, (size_t)(0.9 * (double)free_b));
  if (!do_adjoint) budget = 0;
  if (do_adjoint && g_tape_budget == 0 && g_tape.cap >= entry_bytes * 4096) budget = g_tape.cap;   // reuse the cached tape
  const long long cap_entries = (long long)(budget / entry_bytes);

  const int64_t wave_max = std::min<int64_t>(g_wave_cells > 0 ? g_wave_cells : 32768, ncells);
  CachedBuf &b_y0 = g_y0, &b_nacc = g_nacc, &b_off = g_off, &b_tape = g_tape, &b_queue = g_queue;
  CUDA_TRY(b_y0.ensure(sizeof(double) * 32 * wave_max));
  CUDA_TRY(b_nacc.ensure(sizeof(int) * wave_max));
  CUDA_TRY(b_off.ensure(sizeof(long long) * (wave_max + 1)));
  CUDA_TRY(b_queue.ensure(2 * sizeof(unsigned long long)));
  if (do_adjoint) CUDA_TRY(b_tape.ensure(budget));
  unsigned long long* d_queue = b_queue.as<unsigned long long>();
  int* d_novf = reinterpret_cast<int*>(d_queue + 1);
  PhaseTimer tm(st);

  const size_t adj2_smem = sizeof(double) * (size_t)Adj2Smem<Model>::DOUBLES * ADJ2_SLOTS;
  CUDA_TRY(cudaFuncSetAttribute(k_ros_adj2<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj2_smem));
  auto launch_adjoint = [&](int64_t c0, int64_t n, long slab_cap) -> int {
    CUDA_TRY(cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), st));
    tm.start();
    const int ablocks = (int)std::min<int64_t>(sms, (n + ADJ2_SLOTS - 1) / ADJ2_SLOTS);
    k_ros_adj2<Model><<<ablocks, ADJ2_THREADS, adj2_smem, st>>>(
        rp, mc, n, nadj, with_mu ? 1 : 0, b_tape.as<double>(), b_off.as<long long>(), d_ierr + c0, d_y + c0 * 32,
        d_lambda + c0 * nadj * 32, with_mu ? d_mu + c0 * nadj * Model::NP : nullptr, d_istatus + c0 * 20, d_queue,
        slab_cap, b_nacc.as<int>());
    CUDA_TRY(cudaGetLastError());
    g_stats.ms_adjoint += tm.stop();
    g_stats.launches++;
    g_stats.launches_adjoint++;
    return 0;
  };

  // Exact (two-pass) wave: forward for the counts, prefix-sum offsets, replay with tape, adjoint.
  // Returns the number of leading cells that were completed (the rest did not fit the tape).
  auto exact_wave = [&](int64_t c0, int64_t n, int64_t& done, int& max_nacc) -> int {
    double* y = d_y + c0 * 32;
    int* ist = d_istatus + c0 * 20;
    double* rst = d_rstatus + c0 * 20;
    int* ier = d_ierr + c0;
    CUDA_TRY(cudaMemcpyAsync(b_y0.p, y, sizeof(double) * 32 * n, cudaMemcpyDeviceToDevice, st));
    const int fblocks = (int)((n + FWD_WARPS - 1) / FWD_WARPS);
    tm.start();
    k_ros_fwd<Model, 0><<<fblocks, FWD_WARPS * 32, fwd_smem, st>>>(rp, mc, n, b_y0.as<double>(), y, d_atol, d_rtol, ist, rst, ier,
                                                                     b_nacc.as<int>(), nullptr, nullptr, 0, nullptr);
    CUDA_TRY(cudaGetLastError());
    g_stats.ms_forward += tm.stop();
    g_stats.launches++;
    g_stats.launches_forward++;
    done = n;
    if (!do_adjoint) return 0;
    k_scan_offsets<<<1, 1024, 0, st>>>(n, b_nacc.as<int>(), ier, b_off.as<long long>());
    CUDA_TRY(cudaGetLastError());
    g_stats.launches++;
    std::vector<long long> h_off(n + 1);
    std::vector<int> h_nacc(n);
    CUDA_TRY(cudaMemcpyAsync(h_off.data(), b_off.p, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_nacc.data(), b_nacc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int64_t fit = n;
    if (h_off[n] > cap_entries) {   // keep the prefix of cells whose trajectories fit the tape
      fit = std::upper_bound(h_off.begin(), h_off.end(), cap_entries) - h_off.begin() - 1;
      if (fit <= 0) return fail(FATODE_ENOMEM, "trajectory tape of a single cell exceeds the tape budget");
    }
    for (int64_t i = 0; i < fit; ++i) max_nacc = std::max(max_nacc, h_nacc[i]);
    g_stats.tape_bytes_peak = std::max<uint64_t>(g_stats.tape_bytes_peak, (uint64_t)h_off[fit] * entry_bytes);
    g_stats.accepted_steps += h_off[fit];
    const int tblocks = (int)((fit + FWD_WARPS - 1) / FWD_WARPS);
    tm.start();
    k_ros_fwd<Model, 1><<<tblocks, FWD_WARPS * 32, fwd_smem, st>>>(rp, mc, fit, b_y0.as<double>(), nullptr, d_atol, d_rtol, ist, rst,
                                                                     ier, b_nacc.as<int>(), b_off.as<long long>(), b_tape.as<double>(),
                                                                     0, nullptr);
    CUDA_TRY(cudaGetLastError());
    g_stats.ms_forward_tape += tm.stop();
    g_stats.launches++;
    g_stats.launches_expand++;
    if (int e = launch_adjoint(c0, fit, 0)) return e;
    if (fit < n)   // cells beyond the prefix: restore their initial state, they run in a later wave
      CUDA_TRY(cudaMemcpyAsync(y + fit * 32, b_y0.as<double>() + fit * 32, sizeof(double) * 32 * (n - fit),
                               cudaMemcpyDeviceToDevice, st));
    done = fit;
    return 0;
  };

  int64_t c0 = 0;
  int max_nacc = 0;
  bool have_estimate = false;
  while (c0 < ncells) {
    int64_t remaining = ncells - c0;
    if (!do_adjoint || !have_estimate) {
      // forward-only call, or first wave: exact mode on a sample to learn the trajectory lengths
      int64_t n = std::min<int64_t>(do_adjoint ? std::min<int64_t>(wave_max, 2048) : wave_max, remaining);
      int64_t done = 0;
      if (int e = exact_wave(c0, n, done, max_nacc)) return e;
      c0 += done;
      g_stats.waves++;
      have_estimate = do_adjoint && max_nacc > 0;
      continue;
    }
    // single-pass waves: fixed slab per cell sized from the longest trajectory seen so far (+30%)
    const long slab = (long)std::min<long long>((long long)rp.max_no_steps + 1, (long long)(1.3 * max_nacc) + 32);
    int64_t n = std::min<int64_t>(std::min<int64_t>(wave_max, remaining), cap_entries / slab);
    if (n < 1) {   // not even one slab fits: fall back to the exact path (it reports ENOMEM if hopeless)
      int64_t done = 0;
      if (int e = exact_wave(c0, std::min<int64_t>(remaining, 64), done, max_nacc)) return e;
      c0 += done;
      g_stats.waves++;
      continue;
    }
    double* y = d_y + c0 * 32;
    CUDA_TRY(cudaMemcpyAsync(b_y0.p, y, sizeof(double) * 32 * n, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_novf, 0, sizeof(int), st));
    const int fblocks = (int)((n + FWD_WARPS - 1) / FWD_WARPS);
    tm.start();
    k_ros_fwd<Model, 2><<<fblocks, FWD_WARPS * 32, fwd_smem, st>>>(rp, mc, n, b_y0.as<double>(), y, d_atol, d_rtol,
                                                                     d_istatus + c0 * 20, d_rstatus + c0 * 20, d_ierr + c0,
                                                                     b_nacc.as<int>(), nullptr, b_tape.as<double>(), slab, d_novf);
    CUDA_TRY(cudaGetLastError());
    g_stats.ms_forward += tm.stop();
    g_stats.launches++;
    g_stats.launches_forward++;
    int h_novf = 0;
    std::vector<int> h_nacc(n);
    CUDA_TRY(cudaMemcpyAsync(&h_novf, d_novf, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_nacc.data(), b_nacc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_novf > 0) {   // some trajectory outgrew its slab: redo this wave exactly (rare), with a larger estimate after
      CUDA_TRY(cudaMemcpyAsync(y, b_y0.p, sizeof(double) * 32 * n, cudaMemcpyDeviceToDevice, st));
      int64_t done = 0;
      if (int e = exact_wave(c0, n, done, max_nacc)) return e;
      c0 += done;
      g_stats.waves++;
      continue;
    }
    long long steps = 0;
    for (int64_t i = 0; i < n; ++i) { steps += h_nacc[i]; max_nacc = std::max(max_nacc, h_nacc[i]); }
    g_stats.accepted_steps += steps;   // (cells with ierr != 1 contribute the steps they made)
    g_stats.tape_bytes_peak = std::max<uint64_t>(g_stats.tape_bytes_peak, (uint64_t)n * slab * entry_bytes);
    if (int e = launch_adjoint(c0, n, slab)) return e;
    c0 += n;
    g_stats.waves++;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// deterministic Mu sum over cells (fixed order -> bitwise reproducible)
static int mu_sum_device(int64_t ncells, int len, const double* d_mu, const int* d_ierr, double* d_mu_sum, cudaStream_t st) {
  const int cpb = 256;
  const int nb = (int)((ncells + cpb - 1) / cpb);
  CachedBuf& part = g_part;
  CUDA_TRY(part.ensure(sizeof(double) * (size_t)nb * len));
  PhaseTimer tm(st);
  tm.start();
  k_mu_reduce1<<<nb, 256, 0, st>>>(ncells, len, d_mu, d_ierr, part.as<double>(), cpb);
  k_mu_reduce2<<<(len + 255) / 256, 256, 0, st>>>(nb, len, part.as<double>(), d_mu_sum);
  CUDA_TRY(cudaGetLastError());
  g_stats.ms_other += tm.stop();
  g_stats.launches += 2;
  if (comm_active()) {   // the path's one collective: Mu summed over the cells of every rank (fatode_comm.cu)
    std::string cerr;
    if (int e = comm_allreduce_device(d_mu_sum, (size_t)len, st, cerr)) return fail(e, cerr);
    g_stats.launches += 1;
  }
  return 0;
}

// gather / scatter of per-cell rows for the cells handed to the general path
__global__ void k_gather_rows(int n, const int* __restrict__ ids, int row, const double* __restrict__ src, double* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * row) return;
  const long c = i / row, j = i - c * row;
  dst[i] = src[(long)ids[c] * row + j];
}
__global__ void k_scatter_rows(int n, const int* __restrict__ ids, int row, const double* __restrict__ src, double* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * row) return;
  const long c = i / row, j = i - c * row;
  dst[(long)ids[c] * row + j] = src[i];
}
__global__ void k_scatter_rows_i(int n, const int* __restrict__ ids, int row, const int* __restrict__ src, int* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * row) return;
  const long c = i / row, j = i - c * row;
  dst[(long)ids[c] * row + j] = src[i];
}

__global__ void k_gather_rows_i(int n, const int* __restrict__ ids, const int* __restrict__ src, int* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[ids[i]];
}


// FAST PATH (models with per-thread code).
//   forward-only calls: k_ros_fwd_cell<MODE 0> over all cells.
//   adjoint calls: a three-stream software pipeline over WAVES of cells and BATCHES inside a wave
//     sF: k_ros_fwd_cell<MODE 1> of wave k+1  -> step entries (T, H, L\U, Y, K) in a chunk pool (two pools, ping-pong)
//     sX: k_expand_tape of batch g+1          -> per-stage operands (J_i, M'_i, a_p) in a "fat" buffer (two buffers)
//     sA: k_ros_adj_cell of batch g           -> Lambda, Mu
//   The forward kernel occupies one warp and 153 KB of shared memory per SM, the backward kernel four warps, 62 KB and
//   the tensor memory, so both are resident on every SM at the same time and the latency-bound forward sweep of the
//   next wave hides behind the backward sweep of the current one.  Cells that need the general kernels are appended
//   to `general`; cells that found their pool exhausted are re-queued.
struct StreamSet {
  cudaStream_t sF = nullptr, sX = nullptr, sA = nullptr;
  cudaEvent_t comp_free[2] = {nullptr, nullptr}, fat_free[2] = {nullptr, nullptr}, expand_done[2] = {nullptr, nullptr};
  cudaEvent_t join = nullptr, doneF = nullptr, doneX = nullptr, doneA = nullptr, tables_ready = nullptr;
  void destroy();
  cudaError_t init() {
    if (sF && tables_ready) return cudaSuccess;
    destroy();   // a half-built set (an earlier init failed) starts over
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&sF, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&sX, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&sA, cudaStreamNonBlocking)) != cudaSuccess) return e;
    cudaEvent_t* all[] = {&comp_free[0], &comp_free[1], &fat_free[0], &fat_free[1], &expand_done[0], &expand_done[1],
                          &join, &doneF, &doneX, &doneA, &tables_ready};
    for (cudaEvent_t* ev : all)
      if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaSuccess;
  }
  // streams and events belong to the device they were created on: released with the workspace (fatode_release_workspace,
  // and check_device() when the host thread moves to another GPU), re-created by the next init()
  void destroy_impl() {
    cudaEvent_t* all[] = {&comp_free[0], &comp_free[1], &fat_free[0], &fat_free[1], &expand_done[0], &expand_done[1],
                          &join, &doneF, &doneX, &doneA, &tables_ready};
    for (cudaEvent_t* ev : all)
      if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
    cudaStream_t* str[] = {&sF, &sX, &sA};
    for (cudaStream_t* q : str)
      if (*q) { cudaStreamDestroy(*q); *q = nullptr; }
  }
};
inline void StreamSet::destroy() { destroy_impl(); }
static thread_local StreamSet g_ss;
static thread_local CachedBuf g_comp[2], g_fat[2], g_cprev[2], g_cowner[2], g_cseq[2], g_wids[2], g_wnacc2[2], g_wlast2[2],
    g_border[2], g_bfatoff[2], g_wbatch[2], g_cnext[2], g_wfirst[2], g_ctr;

static void release_pipeline_workspace() {
  for (int i = 0; i < 2; ++i) {
    g_comp[i].release(); g_fat[i].release(); g_cprev[i].release(); g_cowner[i].release(); g_cseq[i].release();
    g_wids[i].release(); g_wnacc2[i].release(); g_wlast2[i].release(); g_border[i].release(); g_bfatoff[i].release();
    g_wbatch[i].release(); g_cnext[i].release(); g_wfirst[i].release();
  }
  g_ctr.release();
  g_ss.destroy();
}

struct TimedLaunch { cudaEvent_t a, b; int phase; };
// the timing events of one call; whatever is still in the list when the call leaves (error exits included) is destroyed
struct TimedList : std::vector<TimedLaunch> {
  ~TimedList() {
    for (TimedLaunch& t : *this) {
      if (t.a) cudaEventDestroy(t.a);
      if (t.b) cudaEventDestroy(t.b);
    }
  }
};
// error exits of the pipeline: the caller's stream still waits for the side streams, so nothing of this call is in flight
// on them when the next call (or the caller's next kernel) starts
struct StreamJoin {
  cudaStream_t st, s[3];
  cudaEvent_t ev[3];
  bool armed = true;
  ~StreamJoin() {
    if (!armed) return;
    for (int i = 0; i < 3; ++i)
      if (s[i] && s[i] != st && cudaEventRecord(ev[i], s[i]) == cudaSuccess) cudaStreamWaitEvent(st, ev[i], 0);
  }
};

template <class Cell, class Warp>
static int ros_device_fast(const RosParams& rp, const typename Cell::Const& mc, int64_t ncells, int nadj, bool with_mu,
                           bool do_adjoint, double* d_y, double* d_lambda, double* d_mu, const double* d_atol,
                           const double* d_rtol, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st,
                           std::vector<int>& general, bool tlm_sweep = false) {
  // tlm_sweep: the second sweep is the tangent-linear one (forward order, d_lambda = Y_tlm, nadj = NTLM) instead of the
  // discrete adjoint; everything else (waves, batches, tape, expand) is shared
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  CUDA_TRY(g_ctr.ensure(16 * sizeof(unsigned long long)));
  unsigned long long* d_ctr = g_ctr.as<unsigned long long>();   // [0,1] forward queues, [2,3] pool heads, [4,5] backward queues
  CellTape tape{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0u, nullptr};
  // forward kernel: the tensor-memory variant (two warps = 58 cells per SM) unless FATODE_FWD_TM=0 (A/B runs) selects the
  // shared-memory one (one warp = 32 cells per SM); identical bits
  static const bool fwd_tm = !(getenv("FATODE_FWD_TM") && atoi(getenv("FATODE_FWD_TM")) == 0);
  if (!do_adjoint) {
    constexpr int LANES = CELL_LANES_SOLO;
    const size_t cell_smem = CellSmem<Cell, LANES>::BYTES;
    CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd_cell<Cell, 0, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cell_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd_cell_tm<Cell, 0, FWD_TM_LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)CellSmemTM<Cell, FWD_TM_LANES>::BYTES));
    CUDA_TRY(g_wnacc.ensure(sizeof(int) * ncells));
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, 16 * sizeof(unsigned long long), st));
    PhaseTimer tm(st);
    tm.start();
    if (fwd_tm) {
      constexpr int PER_CTA = FWD_TM_LANES * FWD_TM_WARPS;
      const unsigned blocks = (unsigned)std::min<int64_t>((ncells + PER_CTA - 1) / PER_CTA, (int64_t)sms);
      k_ros_fwd_cell_tm<Cell, 0, FWD_TM_LANES><<<blocks, 32 * FWD_TM_WARPS, CellSmemTM<Cell, FWD_TM_LANES>::BYTES, st>>>(
          rp, mc, ncells, nullptr, d_y, d_y, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr, g_wnacc.as<int>(), tape, 0,
          g_debug_irregular_every, d_ctr);
    } else {
      const unsigned blocks = (unsigned)std::min<int64_t>((ncells + LANES - 1) / LANES, (int64_t)sms);
      k_ros_fwd_cell<Cell, 0, LANES><<<blocks, 32, cell_smem, st>>>(rp, mc, ncells, nullptr, d_y, d_y, d_atol, d_rtol, d_istatus,
                                                                    d_rstatus, d_ierr, g_wnacc.as<int>(), tape, 0,
                                                                    g_debug_irregular_every, d_ctr);
    }
    CUDA_TRY(cudaGetLastError());
    g_stats.ms_forward += tm.stop();
    g_stats.launches++;
    g_stats.launches_forward++;
    g_stats.waves++;
    std::vector<int> h_ierr(ncells);
    CUDA_TRY(cudaMemcpyAsync(h_ierr.data(), d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t c = 0; c < ncells; ++c)
      if (h_ierr[c] == IERR_IRREGULAR) general.push_back((int)c);
    return 0;
  }

  // ---- adjoint mode
  // Pipeline mode 0 (default): forward, expand and backward kernels of a wave run back to back, each with the whole
  // GPU (forward: 32 cells per SM).  Modes 1/2 overlap the forward sweep of wave k+1 (26 cells per SM, so that a CTA of
  // the backward kernel fits beside it) with the backward sweep of wave k on separate streams and double-buffered
  // storage; measured on B200 the co-resident kernels slow each other down by more than the overlap gains (the
  // forward kernel's 190 KB of straight-line code evicts the backward kernel's from the instruction cache), so it is
  // kept as an option (fatode_set_pipeline) rather than the default.
  const int pmode = getenv("FATODE_PIPELINE") ? atoi(getenv("FATODE_PIPELINE")) : g_pipeline;
  // Mode 3 (default): forward sweep alone on the GPU (it needs the whole shared memory and tensor memory of an SM), then
  // the batches of the wave with k_expand_tape of batch g+1 (HBM bound, 75 % of the bandwidth when it runs alone) beside
  // the backward sweep of batch g (FP64-latency bound, 8 % of the bandwidth): two fat buffers, one pool.
  const int nbuf = (pmode == 1 || pmode == 2) ? 2 : 1;        // chunk pools
  const int nfat = (pmode >= 1) ? 2 : 1;                      // fat buffers
  const bool fwd_shared = (pmode == 1 || pmode == 2);          // forward sweep co-resident with the backward sweep
  const bool use_tm = fwd_tm && !fwd_shared;
  const int LANES = fwd_shared ? CELL_LANES_SHARED : (use_tm ? FWD_TM_LANES * FWD_TM_WARPS : CELL_LANES_SOLO);   // cells per SM
  const int S = rp.S;
  const size_t cell_smem = fwd_shared ? CellSmem<Cell, CELL_LANES_SHARED>::BYTES : CellSmem<Cell, CELL_LANES_SOLO>::BYTES;
  const size_t adj_smem = sizeof(double) * ((size_t)ADJ3_WARP_DOUBLES * ADJ3_WARPS + (size_t)ADJ3_PRIV_DOUBLES * (ADJ3_WARPS - 4));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd_cell<Cell, 1, CELL_LANES_SHARED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)CellSmem<Cell, CELL_LANES_SHARED>::BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd_cell<Cell, 1, CELL_LANES_SOLO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)CellSmem<Cell, CELL_LANES_SOLO>::BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd_cell_tm<Cell, 1, FWD_TM_LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)CellSmemTM<Cell, FWD_TM_LANES>::BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_adj_cell<Warp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem));
  CUDA_TRY(cudaFuncSetAttribute(k_ros_tlm_cell<Warp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem));
  // backward sweep: two warps per cell (solver + accumulator, ros_cell_adj_pair.cuh) unless FATODE_ADJ_PAIR=0 (A/B runs)
  static const bool adj_pair = !(getenv("FATODE_ADJ_PAIR") && atoi(getenv("FATODE_ADJ_PAIR")) == 0);
  const size_t adjp_smem = sizeof(double) * (size_t)ADJP_PAIR_DOUBLES * ADJP_PAIRS;
  CUDA_TRY(cudaFuncSetAttribute(k_ros_adj_pair<Warp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adjp_smem));
  CUDA_TRY(g_ss.init());
  StreamSet ss = g_ss;
  if (pmode <= 1) ss.sX = ss.sF;
  if (pmode <= 0) ss.sA = ss.sF;
  if (pmode == 3) ss.sA = ss.sF;       // forward and backward sweeps alternate on one stream, the expansion runs beside them
  {   // constant Hessian values of the mechanism for k_expand_tape
    CUDA_TRY(g_part.ensure(sizeof(double) * 512));
    k_cbm4_hess<<<1, 1, 0, st>>>(mc.fix, g_part.as<double>());
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_hess, g_part.p, sizeof(double) * 258, 0, cudaMemcpyDeviceToDevice, st));
  }
  // memory plan: chunk pool(s) of step entries and fat buffer(s): 22 % + 78 % of the budget, or 2 x (14 % + 36 %)
  // when the pipeline double-buffers them
  const size_t chunk_bytes = sizeof(double) * (size_t)FT_STEP * FT_CHUNK;
  const size_t fat_step_bytes = sizeof(double) * (size_t)fat_step_doubles(S);
  size_t budget;
  if (g_tape_budget) budget = (size_t)g_tape_budget;
  else if (g_comp[0].cap >= chunk_bytes * 64 && g_fat[0].cap >= fat_step_bytes * 64 && (nbuf == 1 || g_comp[1].cap == g_comp[0].cap) &&
           (nbuf == 2 || g_comp[1].cap == 0) && (nfat == 1 || g_fat[1].cap == g_fat[0].cap) && (nfat == 2 || g_fat[1].cap == 0))
    budget = 0;   // reuse the cached buffers
  else {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += g_comp[0].cap + g_comp[1].cap + g_fat[0].cap + g_fat[1].cap;
    budget = (size_t)(0.6 * (double)free_b);
  }
  // (the tensor-memory forward kernel holds 58 cells per SM: a wave of 148 x 58 cells wants a larger pool)
  const double comp_frac = nbuf == 2 ? 0.14 : (use_tm ? 0.36 : 0.22);
  const size_t comp_bytes = budget ? (size_t)(comp_frac * (double)budget) : g_comp[0].cap;
  const size_t fat_bytes = budget ? (size_t)((1.0 - nbuf * comp_frac) / nfat * (double)budget) : g_fat[0].cap;
  const unsigned pool_chunks = (unsigned)std::min<size_t>(comp_bytes / chunk_bytes, 0x7fffffffu);
  const long long fat_steps_cap = (long long)(fat_bytes / fat_step_bytes);
  if (pool_chunks < 1 || fat_steps_cap < 1) return fail(FATODE_ENOMEM, "tape budget smaller than one chunk of the trajectory tape");
  if (budget) {
    for (int i = 0; i < 2; ++i) { g_comp[i].release(); g_fat[i].release(); }   // re-plan (budget or mode changed)
  }
  const int64_t wave_max = std::min<int64_t>(g_wave_cells > 0 ? g_wave_cells : (int64_t)sms * LANES, ncells);
  for (int i = 0; i < nfat; ++i) CUDA_TRY(g_fat[i].ensure((size_t)fat_steps_cap * fat_step_bytes));
  for (int i = 0; i < nbuf; ++i) {
    CUDA_TRY(g_comp[i].ensure((size_t)pool_chunks * chunk_bytes));
    CUDA_TRY(g_cprev[i].ensure(sizeof(int) * (size_t)pool_chunks));
    CUDA_TRY(g_cowner[i].ensure(sizeof(int) * (size_t)pool_chunks));
    CUDA_TRY(g_cseq[i].ensure(sizeof(int) * (size_t)pool_chunks));
    CUDA_TRY(g_wids[i].ensure(sizeof(int) * wave_max));
    CUDA_TRY(g_wnacc2[i].ensure(sizeof(int) * wave_max));
    CUDA_TRY(g_wlast2[i].ensure(sizeof(int) * wave_max));
    CUDA_TRY(g_border[i].ensure(sizeof(int) * wave_max));
    CUDA_TRY(g_bfatoff[i].ensure(sizeof(long long) * wave_max));
    CUDA_TRY(g_wbatch[i].ensure(sizeof(int) * wave_max));
    if (tlm_sweep) {
      CUDA_TRY(g_cnext[i].ensure(sizeof(int) * (size_t)pool_chunks));
      CUDA_TRY(g_wfirst[i].ensure(sizeof(int) * wave_max));
    }
  }
  CUDA_TRY(g_wierr.ensure(sizeof(int) * wave_max));

  // the side streams start after whatever precedes this call on the caller's stream
  CUDA_TRY(cudaEventRecord(ss.join, st));
  CUDA_TRY(cudaStreamWaitEvent(ss.sF, ss.join, 0));
  CUDA_TRY(cudaStreamWaitEvent(ss.sX, ss.join, 0));
  CUDA_TRY(cudaStreamWaitEvent(ss.sA, ss.join, 0));

  TimedList timed;
  StreamJoin join_guard{st, {ss.sF, ss.sX, ss.sA}, {ss.doneF, ss.doneX, ss.doneA}};
  auto t_begin = [&](cudaStream_t s, int phase) -> int {
    TimedLaunch t{nullptr, nullptr, phase};
    CUDA_TRY(cudaEventCreate(&t.a));
    CUDA_TRY(cudaEventCreate(&t.b));
    CUDA_TRY(cudaEventRecord(t.a, s));
    timed.push_back(t);
    return 0;
  };
  auto t_end = [&](cudaStream_t s) -> int { CUDA_TRY(cudaEventRecord(timed.back().b, s)); return 0; };
  const bool dbg = getenv("FATODE_DEBUG_WAVES") != nullptr;

  std::vector<int> pending(ncells), next_pending, ids, h_nacc, h_ierr, order;
  std::vector<long long> h_fatoff;
  std::vector<int> h_batch;
  for (int64_t c = 0; c < ncells; ++c) pending[c] = (int)c;
  // running estimate of the tape chunks a cell needs (0 = unknown: the first wave is then a small probing one — if the pool
  // ran dry every cell still integrating would be deferred together).  A wave costs the same time whether its lanes are all
  // busy or not, so the estimate of the previous call is reused when the call describes the same problem.
  struct TapeHint { double tstart, tend, rtol0, atol0, hstart; int method, family; double chunks_per_cell; };
  static thread_local TapeHint hint = {0, 0, 0, 0, 0, -1, -1, 0.0};
  double h_tol[2] = {0.0, 0.0};
  CUDA_TRY(cudaMemcpyAsync(&h_tol[0], d_rtol, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&h_tol[1], d_atol, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const bool hint_ok = hint.chunks_per_cell > 0.0 && hint.tstart == rp.tstart && hint.tend == rp.tend && hint.rtol0 == h_tol[0] &&
                       hint.atol0 == h_tol[1] && hint.hstart == rp.hstart && hint.method == rp.S * 100 + (int)(rp.ELO * 10) &&
                       hint.family == rp.family;
  double chunks_per_cell = hint_ok ? hint.chunks_per_cell * 1.05 : 0.0, measured_cpc = 0.0;
  size_t pos = 0;
  int k = 0, g = 0;
  bool comp_used[2] = {false, false}, fat_used[2] = {false, false};
  int rc_fail = 0;
  while (pos < pending.size() || !next_pending.empty()) {
    if (pos >= pending.size()) { pending.swap(next_pending); next_pending.clear(); pos = 0; }
    const int p = (nbuf == 2) ? (k & 1) : 0;
    const int64_t remaining = (int64_t)(pending.size() - pos);
    // wave size: what the chunk pool holds; a batch = what one fat buffer holds; both rounded down to a multiple of the
    // backward kernel's cell slots (4 per SM) so that the last round of cells of every batch is a full one
    const int64_t slots = (int64_t)sms * ADJ3_WARPS;
    int64_t cap, batch_cells = 0;
    if (chunks_per_cell > 0.0) {
      cap = (int64_t)((double)pool_chunks / (chunks_per_cell * 1.02 + 0.5));
      batch_cells = (int64_t)((double)fat_steps_cap / (chunks_per_cell * FT_CHUNK * 1.02));
      if (batch_cells >= slots) batch_cells = batch_cells / slots * slots;
      batch_cells = std::max<int64_t>(batch_cells, 1);
      if (cap >= slots) cap = cap / slots * slots;
    } else {
      cap = std::max<int64_t>(1, std::min<int64_t>(2 * slots, pool_chunks / 64));
    }
    const int64_t n = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(remaining, wave_max), cap));
    ids.assign(pending.begin() + pos, pending.begin() + pos + n);
    pos += n;
    // ---- forward sweep of wave k into pool p (waits until the backward sweep of wave k-2 has released it)
    if (comp_used[p]) CUDA_TRY(cudaStreamWaitEvent(ss.sF, ss.comp_free[p], 0));
    CUDA_TRY(cudaMemcpyAsync(g_wids[p].p, ids.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ss.sF));
    CUDA_TRY(cudaMemsetAsync(d_ctr + p, 0, sizeof(unsigned long long), ss.sF));
    CUDA_TRY(cudaMemsetAsync(d_ctr + 2 + p, 0, sizeof(unsigned long long), ss.sF));
    tape.base = g_comp[p].as<double>();
    tape.chunk_prev = g_cprev[p].as<int>();
    tape.chunk_owner = g_cowner[p].as<int>();
    tape.chunk_seq = g_cseq[p].as<int>();
    tape.chunk_next = tlm_sweep ? g_cnext[p].as<int>() : nullptr;
    tape.first_chunk = tlm_sweep ? g_wfirst[p].as<int>() : nullptr;
    tape.pool_head = reinterpret_cast<unsigned int*>(d_ctr + 2 + p);
    tape.pool_chunks = pool_chunks;
    tape.last_chunk = g_wlast2[p].as<int>();
    if (int e = t_begin(ss.sF, 0)) return e;
    const unsigned fblocks = (unsigned)std::min<int64_t>((n + LANES - 1) / LANES, (int64_t)sms);
    if (use_tm)
      k_ros_fwd_cell_tm<Cell, 1, FWD_TM_LANES><<<fblocks, 32 * FWD_TM_WARPS, CellSmemTM<Cell, FWD_TM_LANES>::BYTES, ss.sF>>>(
          rp, mc, n, g_wids[p].as<int>(), d_y, d_y, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr, g_wnacc2[p].as<int>(), tape,
          with_mu ? 1 : 0, g_debug_irregular_every, d_ctr + p);
    else if (fwd_shared)
      k_ros_fwd_cell<Cell, 1, CELL_LANES_SHARED><<<fblocks, 32, cell_smem, ss.sF>>>(
          rp, mc, n, g_wids[p].as<int>(), d_y, d_y, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr, g_wnacc2[p].as<int>(), tape,
          with_mu ? 1 : 0, g_debug_irregular_every, d_ctr + p);
    else
      k_ros_fwd_cell<Cell, 1, CELL_LANES_SOLO><<<fblocks, 32, cell_smem, ss.sF>>>(
          rp, mc, n, g_wids[p].as<int>(), d_y, d_y, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr, g_wnacc2[p].as<int>(), tape,
          with_mu ? 1 : 0, g_debug_irregular_every, d_ctr + p);
    CUDA_TRY(cudaGetLastError());
    if (int e = t_end(ss.sF)) return e;
    g_stats.launches++;
    g_stats.launches_forward++;
    g_stats.waves++;
    // per-slot outcome (the host waits for this wave's forward sweep only; earlier waves keep the GPU busy meanwhile)
    h_nacc.resize(n);
    h_ierr.resize(n);
    k_gather_rows_i<<<(unsigned)((n + 255) / 256), 256, 0, ss.sF>>>((int)n, g_wids[p].as<int>(), d_ierr, g_wierr.as<int>());
    CUDA_TRY(cudaMemcpyAsync(h_ierr.data(), g_wierr.p, sizeof(int) * n, cudaMemcpyDeviceToHost, ss.sF));
    CUDA_TRY(cudaMemcpyAsync(h_nacc.data(), g_wnacc2[p].p, sizeof(int) * n, cudaMemcpyDeviceToHost, ss.sF));
    unsigned long long h_head = 0;
    CUDA_TRY(cudaMemcpyAsync(&h_head, d_ctr + 2 + p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ss.sF));
    CUDA_TRY(cudaStreamSynchronize(ss.sF));
    const unsigned chunks_used = (unsigned)std::min<unsigned long long>(h_head & 0xffffffffull, pool_chunks);
    int64_t n_ok = 0, n_def = 0;
    long long steps = 0;
    order.clear();
    for (int64_t w = 0; w < n; ++w) {
      if (h_ierr[w] == IERR_DEFERRED) { next_pending.push_back(ids[w]); ++n_def; }
      else if (h_ierr[w] == IERR_IRREGULAR) general.push_back(ids[w]);
      else if (h_ierr[w] == 1 || (tlm_sweep && h_ierr[w] < 0 && h_ierr[w] > -100)) { ++n_ok; steps += h_nacc[w]; order.push_back((int)w); }
    }
    g_stats.cells_deferred += n_def;
    if (dbg) fprintf(stderr, "[wave %d] cells %ld ok %ld deferred %ld chunks %u/%u\n", k, (long)n, (long)n_ok, (long)n_def, chunks_used, pool_chunks);
    if (n_ok == 0 && n_def == n) {
      if (n == 1) { rc_fail = fail(FATODE_ENOMEM, "trajectory tape of a single cell exceeds the tape budget"); break; }
      chunks_per_cell = std::max(2.0 * chunks_per_cell, (double)pool_chunks / std::max<int64_t>(1, n / 2));
    }
    if (n_ok > 0) {
      measured_cpc = std::max(measured_cpc, ((double)steps / (double)n_ok) / FT_CHUNK + 0.5);
      chunks_per_cell = std::max(chunks_per_cell, measured_cpc);
      g_stats.accepted_steps += steps;
      g_stats.tape_bytes_peak = std::max<uint64_t>(g_stats.tape_bytes_peak, (uint64_t)chunks_used * chunk_bytes + (uint64_t)std::min<long long>(steps, fat_steps_cap) * fat_step_bytes);
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_nacc[a] > h_nacc[b]; });   // longest first
      // ---- batches of this wave (as many cells as one fat buffer holds): per-wave tables first, uploaded on sF (idle
      // now, so the pageable copies do not stall the host), then per batch expand (sX) -> backward sweep (sA)
      struct Batch { size_t o0, o1; long long steps; };
      std::vector<Batch> batches;
      h_fatoff.assign(n, -1);
      h_batch.assign(n, -1);
      {
        size_t o = 0;
        while (o < order.size()) {
          Batch bt{o, o, 0};
          while (o < order.size() && bt.steps + h_nacc[order[o]] <= fat_steps_cap &&
                 (batch_cells <= 0 || (int64_t)(o - bt.o0) < batch_cells)) {
            h_fatoff[order[o]] = bt.steps;
            h_batch[order[o]] = (int)batches.size();
            bt.steps += h_nacc[order[o]];
            ++o;
          }
          bt.o1 = o;
          if (bt.o1 == bt.o0) { rc_fail = fail(FATODE_ENOMEM, "trajectory tape of a single cell exceeds the tape budget"); break; }
          batches.push_back(bt);
        }
      }
      if (rc_fail) break;
      CUDA_TRY(cudaMemcpyAsync(g_border[p].p, order.data(), sizeof(int) * n_ok, cudaMemcpyHostToDevice, ss.sF));
      CUDA_TRY(cudaMemcpyAsync(g_bfatoff[p].p, h_fatoff.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, ss.sF));
      CUDA_TRY(cudaMemcpyAsync(g_wbatch[p].p, h_batch.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ss.sF));
      CUDA_TRY(cudaEventRecord(ss.tables_ready, ss.sF));
      CUDA_TRY(cudaStreamWaitEvent(ss.sX, ss.tables_ready, 0));
      for (size_t bi = 0; bi < batches.size(); ++bi) {
        const Batch& bt = batches[bi];
        const int f = (nfat == 2) ? (g & 1) : 0;
        const int64_t nb = (int64_t)(bt.o1 - bt.o0);
        if (fat_used[f]) CUDA_TRY(cudaStreamWaitEvent(ss.sX, ss.fat_free[f], 0));
        if (int e = t_begin(ss.sX, 1)) return e;
        if (chunks_used > 0)   // no accepted step in the whole wave (TIN == TOUT): nothing to expand, Lambda = ADJINIT
          k_expand_tape<Cell><<<chunks_used, FT_CHUNK * S, 0, ss.sX>>>(
            rp, mc, chunks_used, g_cowner[p].as<int>(), g_cseq[p].as<int>(), g_wids[p].as<int>(), d_ierr, g_wnacc2[p].as<int>(),
            g_wlast2[p].as<int>(), g_comp[p].as<double>(), g_fat[f].as<double>(), g_bfatoff[p].as<long long>(),
            g_wbatch[p].as<int>(), (int)bi, tlm_sweep ? 2 : (with_mu ? 1 : 0));
        CUDA_TRY(cudaGetLastError());
        if (int e = t_end(ss.sX)) return e;
        CUDA_TRY(cudaEventRecord(ss.expand_done[f], ss.sX));
        CUDA_TRY(cudaStreamWaitEvent(ss.sA, ss.expand_done[f], 0));
        CUDA_TRY(cudaMemsetAsync(d_ctr + 4 + f, 0, sizeof(unsigned long long), ss.sA));
        if (int e = t_begin(ss.sA, 2)) return e;
        const int ablocks = (int)std::min<int64_t>(sms, (nb + ADJ3_WARPS - 1) / ADJ3_WARPS);
        if (tlm_sweep) {
          TlmArgs ta{nb, g_wids[p].as<int>(), g_wnacc2[p].as<int>(), g_wfirst[p].as<int>(), g_comp[p].as<double>(), g_cnext[p].as<int>(),
                     g_fat[f].as<double>(), g_bfatoff[p].as<long long>(), d_ierr, d_lambda, d_istatus, d_ctr + 4 + f,
                     g_border[p].as<int>() + bt.o0, nadj};
          k_ros_tlm_cell<Warp><<<ablocks, ADJ3_WARPS * 32, adj_smem, ss.sA>>>(rp, ta);
        } else {
          AdjArgs aa{nb, g_wids[p].as<int>(), g_wnacc2[p].as<int>(), g_wlast2[p].as<int>(), g_comp[p].as<double>(), g_cprev[p].as<int>(),
                     g_fat[f].as<double>(), g_bfatoff[p].as<long long>(), d_ierr, d_y, d_lambda, with_mu ? d_mu : nullptr, d_istatus,
                     d_ctr + 4 + f, g_border[p].as<int>() + bt.o0, nadj, with_mu ? 1 : 0};
          if (adj_pair) k_ros_adj_pair<Warp><<<ablocks, ADJP_PAIRS * 64, adjp_smem, ss.sA>>>(rp, aa);
          else k_ros_adj_cell<Warp><<<ablocks, ADJ3_WARPS * 32, adj_smem, ss.sA>>>(rp, aa);
        }
        CUDA_TRY(cudaGetLastError());
        if (int e = t_end(ss.sA)) return e;
        CUDA_TRY(cudaEventRecord(ss.fat_free[f], ss.sA));
        fat_used[f] = true;
        g_stats.launches += 2;
        g_stats.launches_expand++;
        g_stats.launches_adjoint++;
        if (dbg) fprintf(stderr, "[wave %d] batch %d: %ld cells, %lld steps\n", k, g, (long)nb, bt.steps);
        ++g;
      }
      if (rc_fail) break;
    }
    CUDA_TRY(cudaEventRecord(ss.comp_free[p], ss.sA));
    comp_used[p] = true;
    ++k;
  }
  if (measured_cpc > 0.0 && !rc_fail)
    hint = TapeHint{rp.tstart, rp.tend, h_tol[0], h_tol[1], rp.hstart, rp.S * 100 + (int)(rp.ELO * 10), rp.family, measured_cpc};
  // join: the caller's stream continues after all three side streams
  CUDA_TRY(cudaEventRecord(ss.doneF, ss.sF));
  CUDA_TRY(cudaEventRecord(ss.doneX, ss.sX));
  CUDA_TRY(cudaEventRecord(ss.doneA, ss.sA));
  CUDA_TRY(cudaStreamWaitEvent(st, ss.doneF, 0));
  CUDA_TRY(cudaStreamWaitEvent(st, ss.doneX, 0));
  CUDA_TRY(cudaStreamWaitEvent(st, ss.doneA, 0));
  join_guard.armed = false;
  CUDA_TRY(cudaStreamSynchronize(st));
  for (TimedLaunch& t : timed) {   // device time per phase (the phases overlap; their sum exceeds the elapsed time)
    float ms = 0.f;
    if (t.b && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
      if (t.phase == 0) g_stats.ms_forward += ms; else if (t.phase == 1) g_stats.ms_forward_tape += ms; else g_stats.ms_adjoint += ms;
      if (dbg) fprintf(stderr, "[timing] phase %d %.2f ms\n", t.phase, ms);
    }
  }
  return rc_fail;
}

// Small models (roberts, small_strato): forward sweep only, one system per thread, everything inside one kernel
// (dense DGETF2 with full partial pivoting and the singular-matrix retry per thread; no hand-back needed).
template <class Cell>
static int ros_device_small(const RosParams& rp, const typename Cell::Const& mc, int64_t ncells, double* d_y,
                            const double* d_atol, const double* d_rtol, int* d_istatus, double* d_rstatus, int* d_ierr,
                            cudaStream_t st) {
  std::memset(&g_stats, 0, sizeof g_stats);
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  constexpr int LANES = 32;
  const size_t smem = CellSmem<Cell, LANES>::BYTES;
  CUDA_TRY(cudaFuncSetAttribute(k_ros_fwd_cell<Cell, 0, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(g_ctr.ensure(16 * sizeof(unsigned long long)));
  CUDA_TRY(g_wnacc.ensure(sizeof(int) * ncells));
  CUDA_TRY(cudaMemsetAsync(g_ctr.p, 0, 16 * sizeof(unsigned long long), st));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ros_fwd_cell<Cell, 0, LANES>, 32, smem));
  const unsigned blocks = (unsigned)std::min<int64_t>((ncells + LANES - 1) / LANES, (int64_t)sms * std::max(per_sm, 1));
  CellTape tape{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0u, nullptr};
  PhaseTimer tm(st);
  tm.start();
  k_ros_fwd_cell<Cell, 0, LANES><<<blocks, 32, smem, st>>>(rp, mc, ncells, nullptr, d_y, d_y, d_atol, d_rtol, d_istatus, d_rstatus,
                                                            d_ierr, g_wnacc.as<int>(), tape, 0, 0, g_ctr.as<unsigned long long>());
  CUDA_TRY(cudaGetLastError());
  g_stats.ms_forward += tm.stop();
  g_stats.launches++;
  g_stats.launches_forward++;
  g_stats.waves++;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// INTEGRATE_TLM.  Forward error control (ICNTRL(12) = 0): state sweep + tangent-linear sweep on the per-thread path; systems the
// static-pattern LU cannot take are integrated by the lock-step kernel (general partial pivoting, ros_tlm_lock.cu).  TLM
// truncation-error control (ICNTRL(12) /= 0, the preset of INTEGRATE_TLM): every system goes through the lock-step kernel.
template <class Cell, class Warp>
static int ros_tlm_device(const RosParams& rp, const typename Warp::Const& mc, int64_t ncells, int ntlm, bool trunc, double* d_y,
                          double* d_ytlm, const double* d_atol, const double* d_rtol, const double* d_atol_tlm,
                          const double* d_rtol_tlm, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st) {
  std::memset(&g_stats, 0, sizeof g_stats);
  UnitTiming ut;
  std::string uerr;
  if (trunc || g_path == 1) {
    if (int e = ros_tlm_lock_device(rp, mc, ncells, nullptr, ntlm, trunc, d_y, d_ytlm, d_atol, d_rtol, d_atol_tlm, d_rtol_tlm, d_istatus,
                                    d_rstatus, d_ierr, st, ut, uerr)) return fail(e, uerr);
    g_stats.ms_forward += ut.ms;
    g_stats.launches += ut.launches;
    g_stats.launches_forward += 1;
    g_stats.waves = 1;
    g_stats.cells_general = ncells;
    return 0;
  }
  std::vector<int> general;
  if (int e = ros_device_fast<Cell, Warp>(rp, mc, ncells, ntlm, false, true, d_y, d_ytlm, nullptr, d_atol, d_rtol, d_istatus,
                                          d_rstatus, d_ierr, st, general, /*tlm_sweep=*/true)) return e;
  g_stats.cells_general = (int64_t)general.size();
  if (!general.empty()) {   // these systems kept Y / Y_tlm untouched: integrate them in lock-step with general pivoting
    std::sort(general.begin(), general.end());
    DevBuf b_ids;
    CUDA_TRY(b_ids.alloc(sizeof(int) * general.size()));
    CUDA_TRY(cudaMemcpyAsync(b_ids.p, general.data(), sizeof(int) * general.size(), cudaMemcpyHostToDevice, st));
    if (int e = ros_tlm_lock_device(rp, mc, (long)general.size(), b_ids.as<int>(), ntlm, false, d_y, d_ytlm, d_atol, d_rtol, nullptr,
                                    nullptr, d_istatus, d_rstatus, d_ierr, st, ut, uerr)) return fail(e, uerr);
    g_stats.ms_forward += ut.ms;
    g_stats.launches += ut.launches;
    g_stats.launches_forward += 1;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// dispatcher: fast path where the model has one, then the general path for what is left
template <class Cell, class Warp>
static int ros_device(const RosParams& rp, const typename Warp::Const& mc, int64_t ncells, int nadj, bool with_mu,
                      bool do_adjoint, double* d_y, double* d_lambda, double* d_mu, const double* d_atol, const double* d_rtol,
                      int* d_istatus, double* d_rstatus, int* d_ierr, double* d_mu_sum, cudaStream_t st) {
  std::memset(&g_stats, 0, sizeof g_stats);
  if (g_path == 1) {
    if (int e = ros_device_general<Warp>(rp, mc, ncells, nadj, with_mu, do_adjoint, d_y, d_lambda, d_mu, d_atol, d_rtol,
                                         d_istatus, d_rstatus, d_ierr, st)) return e;
    g_stats.cells_general = ncells;
  } else {
    std::vector<int> general;
    if (int e = ros_device_fast<Cell, Warp>(rp, mc, ncells, nadj, with_mu, do_adjoint, d_y, d_lambda, d_mu, d_atol, d_rtol,
                                            d_istatus, d_rstatus, d_ierr, st, general)) return e;
    const int m = (int)general.size();
    g_stats.cells_general = m;
    if (m > 0) {   // hand-back: compact copies of the rows, general kernels, scatter
      std::sort(general.begin(), general.end());
      const int N = Warp::N, NP = Warp::NP;
      DevBuf b_ids, b_y, b_lam, b_mu, b_ist, b_rst, b_ierr;
      CUDA_TRY(b_ids.alloc(sizeof(int) * m));
      CUDA_TRY(b_y.alloc(sizeof(double) * N * m));
      CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * m));
      CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * m));
      CUDA_TRY(b_ierr.alloc(sizeof(int) * m));
      CUDA_TRY(cudaMemcpyAsync(b_ids.p, general.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st));
      auto blocks = [](long n) { return (unsigned)((n + 255) / 256); };
      k_gather_rows<<<blocks((long)m * N), 256, 0, st>>>(m, b_ids.as<int>(), N, d_y, b_y.as<double>());
      if (do_adjoint) {
        CUDA_TRY(b_lam.alloc(sizeof(double) * (size_t)N * nadj * m));
        k_gather_rows<<<blocks((long)m * N * nadj), 256, 0, st>>>(m, b_ids.as<int>(), N * nadj, d_lambda, b_lam.as<double>());
        if (with_mu) {
          CUDA_TRY(b_mu.alloc(sizeof(double) * (size_t)NP * nadj * m));
          k_gather_rows<<<blocks((long)m * NP * nadj), 256, 0, st>>>(m, b_ids.as<int>(), NP * nadj, d_mu, b_mu.as<double>());
        }
      }
      CUDA_TRY(cudaGetLastError());
      if (int e = ros_device_general<Warp>(rp, mc, m, nadj, with_mu, do_adjoint, b_y.as<double>(), b_lam.as<double>(),
                                           b_mu.as<double>(), d_atol, d_rtol, b_ist.as<int>(), b_rst.as<double>(),
                                           b_ierr.as<int>(), st)) return e;
      k_scatter_rows<<<blocks((long)m * N), 256, 0, st>>>(m, b_ids.as<int>(), N, b_y.as<double>(), d_y);
      k_scatter_rows<<<blocks((long)m * 20), 256, 0, st>>>(m, b_ids.as<int>(), 20, b_rst.as<double>(), d_rstatus);
      k_scatter_rows_i<<<blocks((long)m * 20), 256, 0, st>>>(m, b_ids.as<int>(), 20, b_ist.as<int>(), d_istatus);
      k_scatter_rows_i<<<blocks((long)m), 256, 0, st>>>(m, b_ids.as<int>(), 1, b_ierr.as<int>(), d_ierr);
      if (do_adjoint) {
        k_scatter_rows<<<blocks((long)m * N * nadj), 256, 0, st>>>(m, b_ids.as<int>(), N * nadj, b_lam.as<double>(), d_lambda);
        if (with_mu)
          k_scatter_rows<<<blocks((long)m * NP * nadj), 256, 0, st>>>(m, b_ids.as<int>(), NP * nadj, b_mu.as<double>(), d_mu);
      }
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(st));
    }
  }
  if (do_adjoint && with_mu && d_mu_sum)
    if (int e = mu_sum_device(ncells, nadj * Warp::NP, d_mu, d_ierr, d_mu_sum, st)) return e;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

static int check_model_dims(int model, int nvar, int np, bool explicit_family = false) {
  int n = 0, p = 0;
  if (fatode_model_dims(model, &n, &p) != 0) return FATODE_EINVAL;
  if ((model == FATODE_MODEL_SWE) != explicit_family)
    return fail(FATODE_EUNSUPPORTED, explicit_family ? "the explicit Runge-Kutta entry points are built for the swe model only"
                                                     : "the swe model is available through the explicit Runge-Kutta entry points only");
  if (n != nvar) return fail(FATODE_EINVAL, "nvar does not match the registered model");
  if (np != 0 && np != p) return fail(FATODE_EINVAL, "np does not match the registered model");
  return 0;
}

// fills per-cell ierr with a parse error (the reference returns before touching Y)
static int broadcast_ierr(int64_t ncells, int code, int* istatus, double* rstatus, int* ierr, int mem, cudaStream_t st) {
  if (mem == FATODE_MEM_HOST) {
    for (int64_t c = 0; c < ncells; ++c) ierr[c] = code;
    std::memset(istatus, 0, sizeof(int) * 20 * ncells);
    std::memset(rstatus, 0, sizeof(double) * 20 * ncells);
    return 0;
  }
  std::vector<int> h(ncells, code);
  CUDA_TRY(cudaMemcpyAsync(ierr, h.data(), sizeof(int) * ncells, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(istatus, 0, sizeof(int) * 20 * ncells, st));
  CUDA_TRY(cudaMemsetAsync(rstatus, 0, sizeof(double) * 20 * ncells, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// Continuous Rosenbrock adjoint (ICNTRL(7) = 3) for cbm4, on top of a forward-only sweep: the extra FUN / JAC (/ time derivative)
// of the last checkpoint (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:1143-1157), ADJINIT of the driver (Lambda = I, Mu = 0), and the
// IERR -7 (-6) with which ros_CadjInt leaves its time loop before the first step (:1511-1519; H is negative there).
__global__ void k_ros_cadj_epilogue_cbm4(int64_t ncells, int nadj, int np, int autonomous, int max_no_steps, const double* y, double* lambda,
                                         double* mu, int* istatus, double* rstatus, int* ierr) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncells || ierr[cell] != 1) return;
  for (int m = 0; m < nadj; ++m)
    for (int i = 0; i < 32; ++i) lambda[((size_t)cell * nadj + m) * 32 + i] = Cbm4Warp::adjinit_lambda(i, m, y[cell * 32 + i]);
  if (mu && np > 0)
    for (int q = 0; q < nadj * np; ++q) mu[(size_t)cell * nadj * np + q] = 0.0;
  int* is_ = istatus + cell * 20;
  is_[Nfun] += autonomous ? 1 : 2;
  is_[Njac] += 1;
  rstatus[cell * 20 + Nhexit] = 0.0;
  ierr[cell] = (is_[Nstp] > max_no_steps) ? -6 : -7;
}

static int ros_batch_impl(int family, int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda,
                          double* mu, double tin, double tout, const double* atol, const double* rtol, const int* icntrl_u,
                          const double* rcntrl_u, int* istatus, double* rstatus, int* ierr, double* mu_sum,
                          const double* model_params, int mem, void* stream, const double* atol_tlm = nullptr,
                          const double* rtol_tlm = nullptr) {
  g_err.clear();
  if (ncells < 0 || !y || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (int e = check_model_dims(model, nvar, np)) return e;
  if (family == FAM_ADJ && (!lambda || nadj < 1)) return fail(FATODE_EINVAL, "lambda is NULL or nadj < 1");
  if (family == FAM_TLM && (!lambda || nadj < 1)) return fail(FATODE_EINVAL, "y_tlm is NULL or ntlm < 1");
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RosParams rp;
  int ic[20];
  const int perr = ros_parse_options(family, nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, rp, ic);
  bool do_adjoint = false, cadj3 = false;
  if (family == FAM_ADJ) {
    const int adjtype = ic[6] == 0 ? 2 : ic[6];
    // ICNTRL(7) = 3 (continuous adjoint): reproduced as the reference behaves — forward sweep, ADJINIT, then IERR -7 at the first
    // statement of ros_CadjInt's time loop (negative H against `H <= Roundoff`, ADJ/ROS_ADJ :1517), see cadj_epilogue below.
    // ICNTRL(7) = 4 (simplified continuous adjoint): ros_SimpleCadjInt multiplies its first stage by the Jacobian instead of its
    // transpose (LSS_Mul_Jac :1741 vs LSS_Mul_Jactr1 :1781) and overflows within tens of steps (oracle/ros.hpp restates it,
    // tests/test_oracle_cadj.py shows it): nothing a caller could want bit for bit.
    if (perr == 1 && adjtype == 4)
      return fail(FATODE_EUNSUPPORTED, "the simplified continuous adjoint (ICNTRL(7)=4) is not built: the reference routine "
                                       "multiplies its first stage by J instead of J^T and overflows (see DESIGN.md section 8)");
    cadj3 = (perr == 1 && adjtype == 3);
    if (perr == 1 && ic[7] != 0) return fail(FATODE_EUNSUPPORTED, "ICNTRL(8) (SaveLU) is not implemented by the reference either");
    do_adjoint = (adjtype == 2);
    if (do_adjoint && nadj > 32) return fail(FATODE_EUNSUPPORTED, "nadj > 32 per call is not supported yet");
  }
  const bool do_tlm = (family == FAM_TLM);
  if (do_tlm) {
    if (perr == 1 && ic[11] != 0 && (!atol_tlm || !rtol_tlm))
      return fail(FATODE_EINVAL, "ROS_TLM with TLM truncation-error control (ICNTRL(12) /= 0, the preset of INTEGRATE_TLM) needs "
                                 "ATOL_TLM and RTOL_TLM");
    if (nadj > 32) return fail(FATODE_EUNSUPPORTED, "ntlm > 32 per call is not supported yet");
    if (model != FATODE_MODEL_CBM4) return fail(FATODE_EUNSUPPORTED, "INTEGRATE_TLM is available for the cbm4 model only in this build");
  }
  if (perr != 1) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  const bool with_mu = (family == FAM_ADJ) && np > 0 && mu != nullptr;
  if (model != FATODE_MODEL_CBM4 && do_adjoint && nadj > 8)
    return fail(FATODE_EUNSUPPORTED, "the small models take at most 8 adjoint columns per call");
  Cbm4Const mc;
  cbm4_constants(model_params, mc);
  SmallConst sc;
  sc.fix[0] = model_params ? model_params[0] : (double)8.120E+16f * (double)1.0f;   // small_strato_dr.F90:55-56 (REAL(4) literals)
  sc.fix[1] = model_params ? model_params[1] : (double)1.697E+16f * (double)1.0f;

  // stage arrays
  DevBuf b_atol, b_rtol, b_y, b_lam, b_mu, b_ist, b_rst, b_ierr, b_musum;
  CUDA_TRY(b_atol.alloc(sizeof(double) * nvar));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * nvar));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_lam = lambda, *d_mu = mu, *d_rst = rstatus, *d_musum = mu_sum;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * nvar * ncells, nl = sizeof(double) * (size_t)nvar * nadj * ncells,
               nm = sizeof(double) * (size_t)np * nadj * ncells;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
    if (family == FAM_ADJ || do_tlm) {
      // Lambda and Mu are INTENT(INOUT) in the reference: cells whose forward sweep fails keep the caller's values
      CUDA_TRY(b_lam.alloc(nl));
      d_lam = b_lam.as<double>();
      CUDA_TRY(cudaMemcpyAsync(d_lam, lambda, nl, cudaMemcpyHostToDevice, st));
      if (with_mu) {
        CUDA_TRY(b_mu.alloc(nm));
        d_mu = b_mu.as<double>();
        CUDA_TRY(cudaMemcpyAsync(d_mu, mu, nm, cudaMemcpyHostToDevice, st));
      }
      if (with_mu && mu_sum) { CUDA_TRY(b_musum.alloc(sizeof(double) * np * nadj)); d_musum = b_musum.as<double>(); }
    }
  }
  int rc;
  if (do_tlm) {
    const bool trunc = ic[11] != 0;
    DevBuf b_atl, b_rtl;          // ATOL_tlm / RTOL_tlm (N, NTLM), shared by the batch like ATOL / RTOL
    if (trunc) {
      const size_t nb = sizeof(double) * (size_t)nvar * nadj;
      CUDA_TRY(b_atl.alloc(nb));
      CUDA_TRY(b_rtl.alloc(nb));
      CUDA_TRY(cudaMemcpyAsync(b_atl.p, atol_tlm, nb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(b_rtl.p, rtol_tlm, nb, cudaMemcpyHostToDevice, st));
    }
    rc = ros_tlm_device<Cbm4Cell, Cbm4Warp>(rp, mc, ncells, nadj, trunc, d_y, d_lam, b_atol.as<double>(), b_rtol.as<double>(),
                                            b_atl.as<double>(), b_rtl.as<double>(), d_ist, d_rst, d_ierr, st);
  }
  else if (model == FATODE_MODEL_CBM4)
    rc = ros_device<Cbm4Cell, Cbm4Warp>(rp, mc, ncells, nadj, with_mu, do_adjoint, d_y, d_lam, d_mu, b_atol.as<double>(),
                                        b_rtol.as<double>(), d_ist, d_rst, d_ierr, d_musum, st);
  else if (do_adjoint) {   // small_strato / roberts: forward sweep, tape and discrete adjoint of a system in one thread
    std::memset(&g_stats, 0, sizeof g_stats);
    UnitTiming ut;
    std::string uerr;
    rc = ros_small_adj_device(model, rp, sc.fix, (long)ncells, nadj, with_mu, d_y, d_lam, d_mu, b_atol.as<double>(),
                              b_rtol.as<double>(), d_ist, d_rst, d_ierr, (size_t)g_tape_budget, st, ut, uerr);
    if (rc != 0) return fail(rc, uerr);
    g_stats.ms_forward = ut.ms;
    g_stats.launches = ut.launches;
    if (with_mu && d_musum)
      if (int e = mu_sum_device(ncells, nadj * np, d_mu, d_ierr, d_musum, st)) return e;
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  else if (model == FATODE_MODEL_SMALL_STRATO)
    rc = ros_device_small<SmallStratoCell>(rp, sc, ncells, d_y, b_atol.as<double>(), b_rtol.as<double>(), d_ist, d_rst, d_ierr, st);
  else
    rc = ros_device_small<RobertsCell>(rp, sc, ncells, d_y, b_atol.as<double>(), b_rtol.as<double>(), d_ist, d_rst, d_ierr, st);
  if (rc != 0) return rc;
  if (cadj3) {   // continuous adjoint: ADJINIT and the immediate IERR -7 of ros_CadjInt on top of the forward-only sweep
    if (model == FATODE_MODEL_CBM4) {
      k_ros_cadj_epilogue_cbm4<<<(unsigned)((ncells + 127) / 128), 128, 0, st>>>(ncells, nadj, with_mu ? np : 0, rp.autonomous ? 1 : 0,
                                                                                rp.max_no_steps, d_y, d_lam, d_mu, d_ist, d_rst, d_ierr);
      CUDA_TRY(cudaGetLastError());
    } else {
      if (int e = ros_small_cadj_epilogue_device(model, rp, sc.fix, (long)ncells, nadj, with_mu, d_y, d_lam, d_mu, d_ist, d_rst, d_ierr, st))
        return fail(e, "continuous-adjoint epilogue of the small models failed (nadj must be 1..8)");
    }
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    if (do_tlm) CUDA_TRY(cudaMemcpyAsync(lambda, d_lam, nl, cudaMemcpyDeviceToHost, st));
    if (family == FAM_ADJ && (do_adjoint || cadj3)) {
      CUDA_TRY(cudaMemcpyAsync(lambda, d_lam, nl, cudaMemcpyDeviceToHost, st));
      if (with_mu) CUDA_TRY(cudaMemcpyAsync(mu, d_mu, nm, cudaMemcpyDeviceToHost, st));
      if (with_mu && mu_sum) CUDA_TRY(cudaMemcpyAsync(mu_sum, d_musum, sizeof(double) * np * nadj, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int fatode_ros_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                                          const double* atol, const double* rtol, const int* icntrl_u,
                                          const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                          const double* model_params, int mem, void* stream) {
  return ros_batch_impl(FAM_FWD, model, ncells, nvar, 0, 0, y, nullptr, nullptr, tin, tout, atol, rtol, icntrl_u, rcntrl_u,
                        istatus, rstatus, ierr, nullptr, model_params, mem, stream);
}

extern "C" int fatode_ros_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y,
                                              double* lambda, double* mu, double tin, double tout, const double* atol_adj,
                                              const double* rtol_adj, const double* atol, const double* rtol,
                                              const int* icntrl_u, const double* rcntrl_u, int* istatus, double* rstatus,
                                              int* ierr, double* mu_sum, const double* model_params, int mem, void* stream) {
  (void)atol_adj; (void)rtol_adj;
  return ros_batch_impl(FAM_ADJ, model, ncells, nvar, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl_u, rcntrl_u,
                        istatus, rstatus, ierr, mu_sum, model_params, mem, stream);
}

extern "C" int fatode_ros_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm, double tin,
                                              double tout, const double* atol_tlm, const double* rtol_tlm, const double* atol,
                                              const double* rtol, const int* icntrl_u, const double* rcntrl_u, int* istatus,
                                              double* rstatus, int* ierr, const double* model_params, int mem, void* stream) {
  // ATOL_tlm / RTOL_tlm are read only when ICNTRL(12) /= 0 (TLM truncation-error control, the preset)
  return ros_batch_impl(FAM_TLM, model, ncells, nvar, 0, ntlm, y, y_tlm, nullptr, tin, tout, atol, rtol, icntrl_u, rcntrl_u,
                        istatus, rstatus, ierr, nullptr, model_params, mem, stream, atol_tlm, rtol_tlm);
}

// ------------------------------------------------------------------------------------------------
// FWD/SDIRK and FWD/RK INTEGRATE: one system per thread for all three models
template <class Params, class Const, class... Extra>
using CellFwdKernel = void (*)(Params, Const, long, const double*, double*, const double*, const double*, int*, double*, int*,
                               unsigned long long*, const int*, Extra...);

template <class Params, class Const, class... Extra>
static int cellfwd_launch(CellFwdKernel<Params, Const, Extra...> kern, size_t smem, int lanes, const Params& prm, const Const& mc,
                          int64_t ncells, double* d_y, const double* d_atol, const double* d_rtol, int* d_istatus,
                          double* d_rstatus, int* d_ierr, cudaStream_t st, const int* d_cell_ids, Extra... extra) {
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(g_ctr.ensure(16 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(g_ctr.p, 0, 16 * sizeof(unsigned long long), st));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  const unsigned blocks = (unsigned)std::min<int64_t>((ncells + lanes - 1) / lanes, (int64_t)sms * std::max(per_sm, 1));
  PhaseTimer tm(st);
  tm.start();
  kern<<<blocks, 32, smem, st>>>(prm, mc, (long)ncells, d_y, d_y, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr,
                                 g_ctr.as<unsigned long long>(), d_cell_ids, extra...);
  CUDA_TRY(cudaGetLastError());
  g_stats.ms_forward += tm.stop();
  g_stats.launches++;
  g_stats.launches_forward++;
  g_stats.waves++;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

enum { IMPL_SDIRK = 0, IMPL_RK = 1 };
// Active lanes per one-warp block of the per-thread forward kernels: the shared-memory footprint per system (6 KB SDIRK, 13.4 KB
// RK: three factor planes + 17 vectors) fixes the number of systems per SM (about 32 / 16).  Measured: spreading them over four
// blocks of 8 / 4 lanes (four schedulers) is SLOWER (SDIRK 8.0 -> 7.2 k systems/s, RK 8.3 -> 3.2 k): the sweeps are bound by the
// issue rate of their one warp, and narrower warps issue the same instruction stream four times.
constexpr int SDIRK_LANES_CBM4 = 32;
constexpr int RK_LANES_CBM4 = 16;
constexpr int RK_LANES_DENSE = 7, SDIRK_LANES_DENSE = 16;   // general-pivoting second pass (dense 32 x 32 planes)

static int implicit_integrate_batch(int family, int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                                    const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                    int* istatus, double* rstatus, int* ierr, const double* model_params, int mem, void* stream) {
  g_err.clear();
  if (ncells < 0 || !y || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (int e = check_model_dims(model, nvar, 0)) return e;
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SdirkParams sp;
  RkParams rp;
  const int perr = (family == IMPL_SDIRK) ? sdirk_parse_options(nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, sp)
                                          : rk_parse_options(nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, rp);
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  Cbm4Const mc;
  cbm4_constants(model_params, mc);
  SmallConst sc;
  sc.fix[0] = model_params ? model_params[0] : (double)8.120E+16f * (double)1.0f;
  sc.fix[1] = model_params ? model_params[1] : (double)1.697E+16f * (double)1.0f;
  DevBuf b_atol, b_rtol, b_y, b_ist, b_rst, b_ierr;
  CUDA_TRY(b_atol.alloc(sizeof(double) * nvar));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * nvar));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_rst = rstatus;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * nvar * ncells;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
  }
  if (model == FATODE_MODEL_CBM4) CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  const double *da = b_atol.as<double>(), *dr = b_rtol.as<double>();
  std::memset(&g_stats, 0, sizeof g_stats);
  int rc;
  if (family == IMPL_SDIRK) {
    if (model == FATODE_MODEL_CBM4)
      rc = cellfwd_launch<SdirkParams, Cbm4Const, double*, int, int>(k_sdirk_fwd_cell<Cbm4Cell, SDIRK_LANES_CBM4>, SdirkSmem<Cbm4Cell, SDIRK_LANES_CBM4>::BYTES, SDIRK_LANES_CBM4, sp, mc, ncells, d_y, da, dr, d_ist, d_rst, d_ierr, st, nullptr, (double*)nullptr, 0, 0);
    else if (model == FATODE_MODEL_SMALL_STRATO)
      rc = cellfwd_launch<SdirkParams, SmallConst, double*, int, int>(k_sdirk_fwd_cell<SmallStratoCell, 32>, SdirkSmem<SmallStratoCell, 32>::BYTES, 32, sp, sc, ncells, d_y, da, dr, d_ist, d_rst, d_ierr, st, nullptr, (double*)nullptr, 0, 0);
    else
      rc = cellfwd_launch<SdirkParams, SmallConst, double*, int, int>(k_sdirk_fwd_cell<RobertsCell, 32>, SdirkSmem<RobertsCell, 32>::BYTES, 32, sp, sc, ncells, d_y, da, dr, d_ist, d_rst, d_ierr, st, nullptr, (double*)nullptr, 0, 0);
  } else {
    if (model == FATODE_MODEL_CBM4)
      rc = cellfwd_launch<RkParams, Cbm4Const, double*, int>(k_rk_fwd_cell<Cbm4Cell, RK_LANES_CBM4>, RkSmem<Cbm4Cell, RK_LANES_CBM4>::BYTES, RK_LANES_CBM4, rp, mc, ncells, d_y, da, dr, d_ist, d_rst, d_ierr, st, nullptr, (double*)nullptr, 0);
    else if (model == FATODE_MODEL_SMALL_STRATO)
      rc = cellfwd_launch<RkParams, SmallConst, double*, int>(k_rk_fwd_cell<SmallStratoCell, 32>, RkSmem<SmallStratoCell, 32>::BYTES, 32, rp, sc, ncells, d_y, da, dr, d_ist, d_rst, d_ierr, st, nullptr, (double*)nullptr, 0);
    else
      rc = cellfwd_launch<RkParams, SmallConst, double*, int>(k_rk_fwd_cell<RobertsCell, 32>, RkSmem<RobertsCell, 32>::BYTES, 32, rp, sc, ncells, d_y, da, dr, d_ist, d_rst, d_ierr, st, nullptr, (double*)nullptr, 0);
  }
  if (rc != 0) return rc;
  if (model == FATODE_MODEL_CBM4) {
    // second pass: systems that left the static pivot pattern kept their initial state; integrate them again with
    // general partial pivoting (same arithmetic as the reference for every pivot sequence, singular retries included)
    std::vector<int> h_ierr((size_t)ncells), ids;
    CUDA_TRY(cudaMemcpyAsync(h_ierr.data(), d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t c = 0; c < ncells; ++c)
      if (h_ierr[(size_t)c] == RK_IERR_IRREGULAR) ids.push_back((int)c);
    static_assert(RK_IERR_IRREGULAR == SDIRK_IERR_IRREGULAR, "one hand-back code for both implicit families");
    if (!ids.empty()) {
      DevBuf b_ids;
      CUDA_TRY(b_ids.alloc(sizeof(int) * ids.size()));
      CUDA_TRY(cudaMemcpyAsync(b_ids.p, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice, st));
      if (family == IMPL_SDIRK)
        rc = cellfwd_launch<SdirkParams, Cbm4Const, double*, int, int>(k_sdirk_fwd_cell<Cbm4DenseCell, SDIRK_LANES_DENSE>, SdirkSmem<Cbm4DenseCell, SDIRK_LANES_DENSE>::BYTES,
                                                   SDIRK_LANES_DENSE, sp, mc, (int64_t)ids.size(), d_y, da, dr, d_ist, d_rst, d_ierr, st, b_ids.as<int>(), (double*)nullptr, 0, 0);
      else
        rc = cellfwd_launch<RkParams, Cbm4Const, double*, int>(k_rk_fwd_cell<Cbm4DenseCell, RK_LANES_DENSE>, RkSmem<Cbm4DenseCell, RK_LANES_DENSE>::BYTES,
                                                RK_LANES_DENSE, rp, mc, (int64_t)ids.size(), d_y, da, dr, d_ist, d_rst, d_ierr, st, b_ids.as<int>(), (double*)nullptr, 0);
      if (rc != 0) return rc;
      g_stats.cells_general = (int64_t)ids.size();
    }
  }
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int fatode_sdirk_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                                            const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                            int* istatus, double* rstatus, int* ierr, const double* model_params, int mem,
                                            void* stream) {
  return implicit_integrate_batch(IMPL_SDIRK, model, ncells, nvar, y, tin, tout, atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus,
                                  ierr, model_params, mem, stream);
}

extern "C" int fatode_rk_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout,
                                         const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                         int* istatus, double* rstatus, int* ierr, const double* model_params, int mem,
                                         void* stream) {
  return implicit_integrate_batch(IMPL_RK, model, ncells, nvar, y, tin, tout, atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus,
                                  ierr, model_params, mem, stream);
}

// ------------------------------------------------------------------------------------------------
// ADJ/RK_ADJ INTEGRATE_ADJ for CBM-IV: taping forward sweep (one system per thread), per-step preparation (one thread per
// accepted step), backward sweep (one warp per system, lane = adjoint column).  rk_cell_adj.cuh.
static thread_local CachedBuf g_rk_tape, g_rk_fat, g_rk_idx, g_rk_stat;
static int g_rk_tape_slot = 0;   // accepted steps per system in the first pass (0 = 1024)
extern "C" int fatode_set_rk_tape_slot(int steps) { g_rk_tape_slot = steps; return 0; }
static void release_rk_workspace() { g_rk_tape.release(); g_rk_fat.release(); g_rk_idx.release(); g_rk_stat.release(); }

__global__ void k_rk_gather_status(int n, const int* __restrict__ ids, const int* __restrict__ ierr, const int* __restrict__ istatus,
                                   int* __restrict__ out) {   // out[2 i] = ierr, out[2 i + 1] = Nacc
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = ids[i];
  out[2 * i] = ierr[c];
  out[2 * i + 1] = istatus[(size_t)c * 20 + Nacc];
}

// One group of systems (ids) through forward + backward.  Systems that overflow their tape slot are appended to `overflow`.
static thread_local bool g_rk_tlm = false;          // the current call is INTEGRATE_TLM of TLM/RK_TLM: forward-order big-solve column pass
static thread_local bool g_rk_adj_direct = false;   // ICNTRL(7) = 2 of the current call: one 3N x 3N factorisation per step (rk_big_adj.cuh)
static int rk_adj_group(const RkParams& rp, const Cbm4Const& mc, const std::vector<int>& ids, int cap, int nadj, bool with_mu,
                        int np, bool do_adjoint, size_t fat_budget, double* d_y, double* d_lam, double* d_mu, const double* d_atol,
                        const double* d_rtol, int* d_ist, double* d_rst, int* d_ierr, cudaStream_t st, int sms,
                        std::vector<int>& overflow, bool dense = false) {
  const int n = (int)ids.size();
  if (n == 0) return 0;
  // index buffer: [ids n][slot n][cell n] ints, then [rec_off n + 1] long long
  const size_t idx_bytes = sizeof(int) * (size_t)n * 3 + 16 + sizeof(long long) * (size_t)(n + 1);
  CUDA_TRY(g_rk_idx.ensure(idx_bytes));
  int* d_ids = g_rk_idx.as<int>();
  int* d_slot = d_ids + n;
  int* d_cell = d_slot + n;
  long long* d_off = reinterpret_cast<long long*>(reinterpret_cast<char*>(g_rk_idx.p) + ((sizeof(int) * (size_t)n * 3 + 15) / 16) * 16);
  CUDA_TRY(cudaMemcpyAsync(d_ids, ids.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  if (do_adjoint) CUDA_TRY(g_rk_tape.ensure(sizeof(double) * (size_t)n * cap * RK_REC));
  if (dense) {   // general partial pivoting (systems the static-pattern pass handed back)
    if (int e = cellfwd_launch<RkParams, Cbm4Const, double*, int>(k_rk_fwd_cell<Cbm4DenseCell, RK_LANES_DENSE, true>, RkSmem<Cbm4DenseCell, RK_LANES_DENSE>::BYTES,
                                                                 RK_LANES_DENSE, rp, mc, n, d_y, d_atol, d_rtol, d_ist, d_rst, d_ierr, st, d_ids,
                                                                 do_adjoint ? g_rk_tape.as<double>() : (double*)nullptr, cap))
      return e;
  } else {
    if (int e = cellfwd_launch<RkParams, Cbm4Const, double*, int>(k_rk_fwd_cell<Cbm4Cell, RK_LANES_CBM4, true>, RkSmem<Cbm4Cell, RK_LANES_CBM4>::BYTES,
                                                                 RK_LANES_CBM4, rp, mc, n, d_y, d_atol, d_rtol, d_ist, d_rst, d_ierr, st, d_ids,
                                                                 do_adjoint ? g_rk_tape.as<double>() : (double*)nullptr, cap))
      return e;
  }
  if (!do_adjoint) return 0;
  CUDA_TRY(g_rk_stat.ensure(sizeof(int) * 2 * (size_t)n));
  k_rk_gather_status<<<(n + 255) / 256, 256, 0, st>>>(n, d_ids, d_ierr, d_ist, g_rk_stat.as<int>());
  std::vector<int> h_stat(2 * (size_t)n);
  CUDA_TRY(cudaMemcpyAsync(h_stat.data(), g_rk_stat.p, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  // backward sweep in batches that fit the record budget
  const size_t rec_bytes = sizeof(double) * (dense ? DR_REC : RA_REC);
  std::vector<int> b_slot, b_cell;
  std::vector<long long> b_off;
  auto flush = [&]() -> int {
    const int nb = (int)b_slot.size();
    if (nb == 0) return 0;
    const long long nrec = b_off.back();
    CUDA_TRY(g_rk_fat.ensure((size_t)std::max<long long>(nrec, 1) * rec_bytes));
    CUDA_TRY(cudaMemcpyAsync(d_slot, b_slot.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_cell, b_cell.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_off, b_off.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(g_ctr.p, 0, 16 * sizeof(unsigned long long), st));
    PhaseTimer tm(st);
    if (nrec > 0) {
      RkAdjPrepArgs pa{(long)nrec, nb, d_off, d_slot, g_rk_tape.as<double>(), cap, g_rk_fat.as<double>()};
      auto pk = dense ? k_rk_adj_prep_dense : k_rk_adj_prep;
      const int plan = dense ? DR_PREP_LANES : RA_PREP_LANES;
      const size_t psm = sizeof(double) * (size_t)(dense ? 2 * DR_LUN + 1024 + 2 * 32 : 3 * cbm4_cell::NLU + 3 * 32) * plan;
      CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
      const unsigned pblocks = (unsigned)std::min<long long>((nrec + plan - 1) / plan, (long long)sms);
      tm.start();
      pk<<<pblocks, 32, psm, st>>>(rp, mc, pa);
      CUDA_TRY(cudaGetLastError());
      g_stats.ms_forward_tape += tm.stop();
      g_stats.launches++; g_stats.launches_expand++;
    }
    RkAdjArgs aa{nb, d_off, d_cell, g_rk_fat.as<double>(), d_lam, d_mu, d_ist, g_ctr.as<unsigned long long>(), nadj, with_mu ? 1 : 0, np, d_ierr};
    // static path: tensor-memory variant (one CTA of four sweep warps per SM; FATODE_RK_ADJ_SMEM=1 selects the shared-memory variant
    // with two one-warp CTAs per SM); general-pivoting pass: shared-memory variant with dense factors, one warp per SM
    static const bool use_tm = !(getenv("FATODE_RK_ADJ_SMEM") && atoi(getenv("FATODE_RK_ADJ_SMEM")) != 0);
    tm.start();
    if (g_rk_tlm) {
      CUDA_TRY(cudaFuncSetAttribute(k_rk_tlm_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RK_ADJ_BIG_SMEM));
      const RkBigRecord rl = dense ? RkBigRecord{DR_REC, DR_J} : RkBigRecord{RA_REC, RA_J};
      const unsigned ablocks = (unsigned)std::min<int>(nb, 2 * sms);
      k_rk_tlm_big<<<ablocks, RB_THREADS, RK_ADJ_BIG_SMEM, st>>>(rp, aa, rl);
    } else if (g_rk_adj_direct) {
      CUDA_TRY(cudaFuncSetAttribute(k_rk_adj_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RK_ADJ_BIG_SMEM));
      const RkBigRecord rl = dense ? RkBigRecord{DR_REC, DR_J} : RkBigRecord{RA_REC, RA_J};
      const unsigned ablocks = (unsigned)std::min<int>(nb, 2 * sms);
      k_rk_adj_big<<<ablocks, RB_THREADS, RK_ADJ_BIG_SMEM, st>>>(rp, aa, rl);
    } else if (!dense && use_tm) {
      CUDA_TRY(cudaFuncSetAttribute(k_rk_adj_cell_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RK_ADJ_TM_SMEM));
      const unsigned ablocks = (unsigned)std::min<int>((nb + RA_TM_WARPS - 1) / RA_TM_WARPS, sms);
      k_rk_adj_cell_tm<<<ablocks, RA_TM_WARPS * 32, RK_ADJ_TM_SMEM, st>>>(rp, aa);
    } else {
      auto ck = dense ? k_rk_adj_cell_dense : k_rk_adj_cell;
      const size_t csm = dense ? RK_ADJ_DENSE_SMEM : RK_ADJ_SMEM;
      CUDA_TRY(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
      const unsigned ablocks = (unsigned)std::min<int>(nb, (dense ? 1 : 2) * sms);
      ck<<<ablocks, 32, csm, st>>>(rp, aa);
    }
    CUDA_TRY(cudaGetLastError());
    g_stats.ms_adjoint += tm.stop();
    g_stats.launches++; g_stats.launches_adjoint++;
    CUDA_TRY(cudaStreamSynchronize(st));
    b_slot.clear(); b_cell.clear(); b_off.clear();
    return 0;
  };
  for (int i = 0; i < n; ++i) {
    const int code = h_stat[2 * (size_t)i], nacc = h_stat[2 * (size_t)i + 1];
    if (code == IERR_TAPE_OVERFLOW) { overflow.push_back(ids[i]); continue; }
    // failed forward sweep: no adjoint (the reference returns); the tangent-linear columns were advanced inside every accepted step
    if (code != 1 && !(g_rk_tlm && code < 0 && code >= -13)) continue;   // -1..-13: the reference's codes (not the internal hand-backs)
    if (b_off.empty()) b_off.push_back(0);
    if (!b_slot.empty() && (size_t)(b_off.back() + nacc) * rec_bytes > fat_budget)
      if (int e = flush()) return e;
    if (b_off.empty()) b_off.push_back(0);
    b_slot.push_back(i);
    b_cell.push_back(ids[i]);
    b_off.push_back(b_off.back() + nacc);
    g_stats.accepted_steps += nacc;
  }
  return flush();
}

static int rk_adj_device(const RkParams& rp, const Cbm4Const& mc, int64_t ncells, int nadj, bool with_mu, int np, bool do_adjoint,
                         double* d_y, double* d_lam, double* d_mu, const double* d_atol, const double* d_rtol, int* d_ist,
                         double* d_rst, int* d_ierr, double* d_mu_sum, cudaStream_t st) {
  std::memset(&g_stats, 0, sizeof g_stats);
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  free_b += g_rk_tape.cap + g_rk_fat.cap;
  const size_t budget = g_tape_budget ? (size_t)g_tape_budget : (size_t)(0.6 * (double)free_b);
  const int cap = std::min(rp.max_no_steps + 1, g_rk_tape_slot > 0 ? g_rk_tape_slot : 1024);   // Nstp > Max_no_steps stops: one more may be accepted
  int64_t wave = g_wave_cells > 0 ? g_wave_cells : (int64_t)sms * 16 * 2;   // two resident rounds of 16 systems per SM
  wave = std::min<int64_t>(wave, std::max<int64_t>(1, (int64_t)(0.2 * (double)budget / ((double)cap * RK_REC * sizeof(double)))));
  const size_t fat_budget = (size_t)(0.75 * (double)budget);
  auto run_all = [&](const std::vector<int>& all, bool dense) -> int {
    std::vector<int> ids, overflow;
    for (size_t w0 = 0; w0 < all.size(); w0 += (size_t)wave) {
      const size_t w1 = std::min(all.size(), w0 + (size_t)wave);
      ids.assign(all.begin() + (long)w0, all.begin() + (long)w1);
      if (int e = rk_adj_group(rp, mc, ids, cap, nadj, with_mu, np, do_adjoint, fat_budget, d_y, d_lam, d_mu, d_atol, d_rtol, d_ist, d_rst,
                               d_ierr, st, sms, overflow, dense))
        return e;
      g_stats.waves++;
    }
    // systems with more accepted steps than the default slot: again, a few at a time, with room for Max_no_steps
    while (!overflow.empty()) {
      const int capx = rp.max_no_steps + 1;
      const size_t per = (size_t)capx * RK_REC * sizeof(double);
      const size_t take = std::max<size_t>(1, std::min<size_t>(overflow.size(), (size_t)(0.2 * (double)budget) / per));
      std::vector<int> part(overflow.end() - (long)take, overflow.end()), again;
      overflow.resize(overflow.size() - take);
      if (int e = rk_adj_group(rp, mc, part, capx, nadj, with_mu, np, do_adjoint, fat_budget, d_y, d_lam, d_mu, d_atol, d_rtol, d_ist,
                               d_rst, d_ierr, st, sms, again, dense))
        return e;
      if (!again.empty()) return fail(FATODE_ENOMEM, "RK_ADJ tape overflow with a Max_no_steps slot");
      g_stats.waves++;
    }
    return 0;
  };
  std::vector<int> all((size_t)ncells);
  for (int64_t c = 0; c < ncells; ++c) all[(size_t)c] = (int)c;
  if (int e = run_all(all, false)) return e;
  {   // second pass: systems that left the static pivot pattern kept their initial state; general partial pivoting throughout
    std::vector<int> h_ierr((size_t)ncells), irregular;
    CUDA_TRY(cudaMemcpyAsync(h_ierr.data(), d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t c = 0; c < ncells; ++c)
      if (h_ierr[(size_t)c] == RK_IERR_IRREGULAR) irregular.push_back((int)c);
    if (!irregular.empty()) {
      if (int e = run_all(irregular, true)) return e;
      g_stats.cells_general = (int64_t)irregular.size();
    }
  }
  if (with_mu && d_mu_sum && do_adjoint)
    if (int e = mu_sum_device(ncells, nadj * np, d_mu, d_ierr, d_mu_sum, st)) return e;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int fatode_rk_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda,
                                             double* mu, double tin, double tout, const double* atol_adj, const double* rtol_adj,
                                             const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                             int* istatus, double* rstatus, int* ierr, double* mu_sum, const double* model_params,
                                             int mem, void* stream) {
  (void)atol_adj; (void)rtol_adj;   // only used by the adaptive adjoint solve (ICNTRL(7) = 3), which is not built
  g_err.clear();
  if (ncells < 0 || !y || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (int e = check_model_dims(model, nvar, np)) return e;
  if (!lambda || nadj < 1) return fail(FATODE_EINVAL, "lambda is NULL or nadj < 1");
  if (model != FATODE_MODEL_CBM4) return fail(FATODE_EUNSUPPORTED, "the RK discrete adjoint is available for the cbm4 model only in this build");
  if (nadj > 32) return fail(FATODE_EUNSUPPORTED, "nadj > 32 per call is not supported yet");
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RkParams rp;
  RkAdjOptions ao;
  const int perr = rk_parse_options(nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, rp, &ao);
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  const bool do_adjoint = (ao.type == 0 || ao.type == 2);
  if (do_adjoint && ao.solve >= 3)
    return fail(FATODE_EUNSUPPORTED, "RK_ADJ adaptive adjoint solve (ICNTRL(7) = 3) is not built; use the preset fixed-sweep solve or "
                                     "the direct solve (ICNTRL(7) = 2)");
  g_rk_adj_direct = do_adjoint && ao.solve == 2;
  g_rk_tlm = false;
  const bool with_mu = np > 0 && mu != nullptr;
  Cbm4Const mc;
  cbm4_constants(model_params, mc);
  DevBuf b_atol, b_rtol, b_y, b_lam, b_mu, b_ist, b_rst, b_ierr, b_musum;
  CUDA_TRY(b_atol.alloc(sizeof(double) * nvar));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * nvar));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_lam = lambda, *d_mu = mu, *d_rst = rstatus, *d_musum = mu_sum;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * nvar * ncells, nl = sizeof(double) * (size_t)nvar * nadj * ncells,
               nm = sizeof(double) * (size_t)np * nadj * ncells;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
    // Lambda and Mu are INTENT(INOUT) upstream: systems whose forward sweep fails keep the caller's values
    CUDA_TRY(b_lam.alloc(nl));
    d_lam = b_lam.as<double>();
    CUDA_TRY(cudaMemcpyAsync(d_lam, lambda, nl, cudaMemcpyHostToDevice, st));
    if (with_mu) {
      CUDA_TRY(b_mu.alloc(nm));
      d_mu = b_mu.as<double>();
      CUDA_TRY(cudaMemcpyAsync(d_mu, mu, nm, cudaMemcpyHostToDevice, st));
      if (mu_sum) { CUDA_TRY(b_musum.alloc(sizeof(double) * np * nadj)); d_musum = b_musum.as<double>(); }
    }
  }
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  if (int e = rk_adj_device(rp, mc, ncells, nadj, with_mu, np, do_adjoint, d_y, d_lam, d_mu, b_atol.as<double>(), b_rtol.as<double>(),
                            d_ist, d_rst, d_ierr, with_mu ? d_musum : nullptr, st))
    return e;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    if (do_adjoint) {
      CUDA_TRY(cudaMemcpyAsync(lambda, d_lam, nl, cudaMemcpyDeviceToHost, st));
      if (with_mu) CUDA_TRY(cudaMemcpyAsync(mu, d_mu, nm, cudaMemcpyDeviceToHost, st));
      if (with_mu && mu_sum) CUDA_TRY(cudaMemcpyAsync(mu_sum, d_musum, sizeof(double) * np * nadj, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// TLM/RK_TLM INTEGRATE_TLM for CBM-IV in the reference's preset mode (direct TLM solve, ICNTRL(7) = 1; state-only error control):
// the RK_IntegratorTLM flavour of the taping forward sweep, the preparation records of the RK adjoint, then k_rk_tlm_big
// (rk_big_adj.cuh) over the records in forward order.
extern "C" int fatode_rk_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm, double tin,
                                             double tout, const double* atol_tlm, const double* rtol_tlm, const double* atol,
                                             const double* rtol, const int* icntrl_u, const double* rcntrl_u, int* istatus,
                                             double* rstatus, int* ierr, const double* model_params, int mem, void* stream) {
  g_err.clear();
  if (ncells < 0 || !y || !y_tlm || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (int e = check_model_dims(model, nvar, 0)) return e;
  if (ntlm < 1) return fail(FATODE_EINVAL, "ntlm < 1");
  if (model != FATODE_MODEL_CBM4) return fail(FATODE_EUNSUPPORTED, "the Runge-Kutta tangent linear model is available for the cbm4 model only in this build");
  if (ntlm > 32) return fail(FATODE_EUNSUPPORTED, "ntlm > 32 per call is not supported yet");
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RkParams rp;
  RkTlmOptions to;
  const int perr = rk_parse_options(nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, rp, nullptr, &to);
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  if (to.trunc_err)
    return fail(FATODE_EUNSUPPORTED, "RK_TLM with the columns in the truncation error control (ICNTRL(12)) is not built");
  // ICNTRL(7) = 0 (modified-Newton TLM) cannot be selected through INTEGRATE_TLM: the preset 1 survives the `> 0` override rule;
  // ICNTRL(9) only matters in that mode
  Cbm4Const mc;
  cbm4_constants(model_params, mc);
  DevBuf b_atol, b_rtol, b_y, b_yt, b_ist, b_rst, b_ierr;
  CUDA_TRY(b_atol.alloc(sizeof(double) * nvar));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * nvar));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_yt = y_tlm, *d_rst = rstatus;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * nvar * ncells, nl = sizeof(double) * (size_t)nvar * ntlm * ncells;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_yt.alloc(nl));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b_yt.p, y_tlm, nl, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_yt = b_yt.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
  }
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  g_rk_tlm = true;
  g_rk_adj_direct = false;
  const int e = rk_adj_device(rp, mc, ncells, ntlm, /*with_mu=*/false, 0, /*do_adjoint=*/true, d_y, d_yt, nullptr, b_atol.as<double>(),
                              b_rtol.as<double>(), d_ist, d_rst, d_ierr, nullptr, st);
  g_rk_tlm = false;
  if (e) return e;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(y_tlm, d_yt, nl, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// TLM/SDIRK_TLM INTEGRATE_TLM for CBM-IV: taping state sweep (one system per thread), per-step preparation (one thread per
// accepted step), column sweep (one warp per system, lane = TLM column).  sdirk_cell_tlm.cuh.
static thread_local CachedBuf g_sd_tape, g_sd_fat, g_sd_idx, g_sd_stat;
static int g_sdirk_tape_slot = 0;   // accepted steps per system in the first pass (0 = 4608)
extern "C" int fatode_set_sdirk_tape_slot(int steps) { g_sdirk_tape_slot = steps; return 0; }
static void release_sdirk_workspace() { g_sd_tape.release(); g_sd_fat.release(); g_sd_idx.release(); g_sd_stat.release(); }

// One group of systems (ids) through the state sweep and the column sweep.  Systems that overflow their tape slot are
// appended to `overflow`.  A system whose state sweep fails keeps Y_tlm as of its last accepted step, like the reference.
// adj != nullptr: the same organisation for ADJ/SDIRK_ADJ (forward sweep in SD_MODE_ADJ, sdirk_cell_adj.cuh backward); ntlm is then
// NADJ and d_ytlm is Lambda.  A system whose forward sweep fails still gets its backward sweep and IERR = 1, as the reference.
struct SdirkAdjHost { SdirkAdjCoef co; double* d_mu; int with_mu, np; int newton = 0; const double* d_atol_adj = nullptr; const double* d_rtol_adj = nullptr; };
static int sdirk_tlm_group(const SdirkParams& sp, const Cbm4Const& mc, const std::vector<int>& ids, int cap, int ntlm, size_t fat_budget,
                           double* d_y, double* d_ytlm, const double* d_atol, const double* d_rtol, int* d_ist, double* d_rst,
                           int* d_ierr, cudaStream_t st, int sms, std::vector<int>& overflow, const SdirkAdjHost* adj = nullptr,
                           bool dense = false) {
  const int n = (int)ids.size();
  if (n == 0) return 0;
  const size_t idx_bytes = sizeof(int) * (size_t)n * 3 + 16 + sizeof(long long) * (size_t)(n + 1);
  CUDA_TRY(g_sd_idx.ensure(idx_bytes));
  int* d_ids = g_sd_idx.as<int>();
  int* d_slot = d_ids + n;
  int* d_cell = d_slot + n;
  long long* d_off = reinterpret_cast<long long*>(reinterpret_cast<char*>(g_sd_idx.p) + ((sizeof(int) * (size_t)n * 3 + 15) / 16) * 16);
  CUDA_TRY(cudaMemcpyAsync(d_ids, ids.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(g_sd_tape.ensure(sizeof(double) * (size_t)n * cap * SD_REC));
  {
    typedef CellFwdKernel<SdirkParams, Cbm4Const, double*, int, int> Kern;
    Kern kern;
    size_t smem;
    int lanes;
    if (dense) {   // general partial pivoting (systems the static-pattern pass handed back)
      kern = adj ? (Kern)k_sdirk_fwd_cell<Cbm4DenseCell, SDIRK_LANES_DENSE, SD_MODE_ADJ> : (Kern)k_sdirk_fwd_cell<Cbm4DenseCell, SDIRK_LANES_DENSE, SD_MODE_TLM>;
      smem = SdirkSmem<Cbm4DenseCell, SDIRK_LANES_DENSE>::BYTES;
      lanes = SDIRK_LANES_DENSE;
    } else {
      kern = adj ? (Kern)k_sdirk_fwd_cell<Cbm4Cell, SDIRK_LANES_CBM4, SD_MODE_ADJ> : (Kern)k_sdirk_fwd_cell<Cbm4Cell, SDIRK_LANES_CBM4, SD_MODE_TLM>;
      smem = SdirkSmem<Cbm4Cell, SDIRK_LANES_CBM4>::BYTES;
      lanes = SDIRK_LANES_CBM4;
    }
    if (int e = cellfwd_launch<SdirkParams, Cbm4Const, double*, int, int>(kern, smem, lanes, sp, mc, n, d_y, d_atol, d_rtol, d_ist, d_rst, d_ierr,
                                                                         st, d_ids, g_sd_tape.as<double>(), cap, ntlm))
      return e;
  }
  CUDA_TRY(g_sd_stat.ensure(sizeof(int) * 2 * (size_t)n));
  k_rk_gather_status<<<(n + 255) / 256, 256, 0, st>>>(n, d_ids, d_ierr, d_ist, g_sd_stat.as<int>());
  std::vector<int> h_stat(2 * (size_t)n);
  CUDA_TRY(cudaMemcpyAsync(h_stat.data(), g_sd_stat.p, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const size_t rec_bytes = sizeof(double) * (adj ? (dense ? DA_REC : SA_REC) : (dense ? DT_REC : ST_REC));
  std::vector<int> b_slot, b_cell;
  std::vector<long long> b_off;
  auto flush = [&]() -> int {
    const int nb = (int)b_slot.size();
    if (nb == 0) return 0;
    const long long nrec = b_off.back();
    if (adj && adj->newton) {   // simplified-Newton adjoint stages: one warp per system straight from the tape (sdirk_adj_newton.cu)
      CUDA_TRY(cudaMemcpyAsync(d_slot, b_slot.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_cell, b_cell.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_off, b_off.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, st));
      UnitTiming ut;
      std::string uerr;
      if (int e = sdirk_adj_newton_device(sp, mc, adj->co, nb, d_off, d_slot, d_cell, g_sd_tape.as<double>(), cap, ntlm, adj->with_mu,
                                          adj->np, d_ytlm, adj->d_mu, adj->d_atol_adj, adj->d_rtol_adj, d_ist, d_ierr,
                                          g_ctr.as<unsigned long long>(), st, ut, uerr))
        return fail(e, uerr);
      g_stats.ms_adjoint += ut.ms;
      g_stats.launches += ut.launches; g_stats.launches_adjoint++;
      b_slot.clear(); b_cell.clear(); b_off.clear();
      return 0;
    }
    if (adj) {
      CUDA_TRY(g_sd_fat.ensure((size_t)std::max<long long>(nrec, 1) * rec_bytes));
      CUDA_TRY(cudaMemcpyAsync(d_slot, b_slot.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_cell, b_cell.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_off, b_off.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemsetAsync(g_ctr.p, 0, 16 * sizeof(unsigned long long), st));
      PhaseTimer tm(st);
      if (nrec > 0) {
        SdirkAdjPrepArgs pa{(long)nrec, nb, d_off, d_slot, g_sd_tape.as<double>(), cap, g_sd_fat.as<double>()};
        auto pk = dense ? k_sdirk_adj_prep_dense : k_sdirk_adj_prep;
        const int plan = dense ? DA_PREP_LANES : SA_PREP_LANES;
        const size_t psm = sizeof(double) * (size_t)((dense ? DN_LU + 32 : cbm4_cell::NLU + 2 * 32)) * plan;
        CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        int per_sm = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk, 32, psm));
        const long long nwork = nrec * sp.S;
        const unsigned pblocks = (unsigned)std::min<long long>((nwork + plan - 1) / plan, (long long)sms * std::max(per_sm, 1));
        tm.start();
        pk<<<pblocks, 32, psm, st>>>(sp, mc, pa);
        CUDA_TRY(cudaGetLastError());
        g_stats.ms_forward_tape += tm.stop();
        g_stats.launches++; g_stats.launches_expand++;
      }
      SdirkAdjArgs aa{nb, d_off, d_cell, g_sd_fat.as<double>(), d_ytlm, adj->d_mu, d_ist, d_ierr, g_ctr.as<unsigned long long>(), ntlm,
                      adj->with_mu, adj->np};
      // static path: tensor-memory variant (one CTA of four sweep warps per SM; FATODE_SDIRK_ADJ_SMEM=1 selects the shared-memory
      // variant with three one-warp CTAs per SM); general-pivoting pass: shared-memory variant with dense factors
      static const bool use_tm = !(getenv("FATODE_SDIRK_ADJ_SMEM") && atoi(getenv("FATODE_SDIRK_ADJ_SMEM")) != 0);
      tm.start();
      if (!dense && use_tm) {
        CUDA_TRY(cudaFuncSetAttribute(k_sdirk_adj_cell_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDIRK_ADJ_TM_SMEM));
        const unsigned ablocks = (unsigned)std::min<int>((nb + SA_TM_WARPS - 1) / SA_TM_WARPS, sms);
        k_sdirk_adj_cell_tm<<<ablocks, SA_TM_WARPS * 32, SDIRK_ADJ_TM_SMEM, st>>>(sp, adj->co, aa);
      } else {
        auto ck = dense ? k_sdirk_adj_cell_dense : k_sdirk_adj_cell;
        const size_t csm = dense ? SDIRK_ADJ_DENSE_SMEM : SDIRK_ADJ_SMEM;
        CUDA_TRY(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        const unsigned ablocks = (unsigned)std::min<int>(nb, 3 * sms);
        ck<<<ablocks, 32, csm, st>>>(sp, adj->co, aa);
      }
      CUDA_TRY(cudaGetLastError());
      g_stats.ms_adjoint += tm.stop();
      g_stats.launches++; g_stats.launches_adjoint++;
      CUDA_TRY(cudaStreamSynchronize(st));
      b_slot.clear(); b_cell.clear(); b_off.clear();
      return 0;
    }
    CUDA_TRY(g_sd_fat.ensure((size_t)std::max<long long>(nrec, 1) * rec_bytes));
    CUDA_TRY(cudaMemcpyAsync(d_slot, b_slot.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_cell, b_cell.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_off, b_off.data(), sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(g_ctr.p, 0, 16 * sizeof(unsigned long long), st));
    PhaseTimer tm(st);
    if (nrec > 0) {
      SdirkTlmPrepArgs pa{(long)nrec, nb, d_off, d_slot, g_sd_tape.as<double>(), cap, g_sd_fat.as<double>()};
      auto pk = dense ? k_sdirk_tlm_prep_dense : k_sdirk_tlm_prep;
      const int plan = dense ? DT_PREP_LANES : ST_PREP_LANES;
      const size_t psm = sizeof(double) * (size_t)((dense ? DN_LU + 32 : cbm4_cell::NLU + 2 * 32)) * plan;
      CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
      int per_sm = 1;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk, 32, psm));
      const long long nwork = nrec * (sp.S + 1);
      const unsigned pblocks = (unsigned)std::min<long long>((nwork + plan - 1) / plan, (long long)sms * std::max(per_sm, 1));
      tm.start();
      pk<<<pblocks, 32, psm, st>>>(sp, mc, pa);
      CUDA_TRY(cudaGetLastError());
      g_stats.ms_forward_tape += tm.stop();
      g_stats.launches++; g_stats.launches_expand++;
      SdirkTlmArgs aa{nb, d_off, d_cell, g_sd_fat.as<double>(), d_ytlm, g_ctr.as<unsigned long long>(), ntlm};
      // static path: tensor-memory variant (one CTA of four sweep warps per SM; FATODE_SDIRK_TLM_SMEM=1 selects the shared-memory
      // variant with three one-warp CTAs per SM); general-pivoting pass: shared-memory variant with dense factors
      static const bool use_tm = !(getenv("FATODE_SDIRK_TLM_SMEM") && atoi(getenv("FATODE_SDIRK_TLM_SMEM")) != 0);
      tm.start();
      if (!dense && use_tm) {
        CUDA_TRY(cudaFuncSetAttribute(k_sdirk_tlm_cell_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDIRK_TLM_TM_SMEM));
        const unsigned ablocks = (unsigned)std::min<int>((nb + ST_TM_WARPS - 1) / ST_TM_WARPS, sms);
        k_sdirk_tlm_cell_tm<<<ablocks, ST_TM_WARPS * 32, SDIRK_TLM_TM_SMEM, st>>>(sp, aa);
      } else {
        auto ck = dense ? k_sdirk_tlm_cell_dense : k_sdirk_tlm_cell;
        const size_t csm = dense ? SDIRK_TLM_DENSE_SMEM : SDIRK_TLM_SMEM;
        CUDA_TRY(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        const unsigned ablocks = (unsigned)std::min<int>(nb, (dense ? 2 : 3) * sms);   // 71.6 KB (dense: 85 KB) of shared memory per warp
        ck<<<ablocks, 32, csm, st>>>(sp, aa);
      }
      CUDA_TRY(cudaGetLastError());
      g_stats.ms_adjoint += tm.stop();
      g_stats.launches++; g_stats.launches_adjoint++;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    b_slot.clear(); b_cell.clear(); b_off.clear();
    return 0;
  };
  for (int i = 0; i < n; ++i) {
    const int code = h_stat[2 * (size_t)i], nacc = h_stat[2 * (size_t)i + 1];
    if (code == IERR_TAPE_OVERFLOW) { overflow.push_back(ids[i]); continue; }
    if (code == SDIRK_IERR_IRREGULAR || (!adj && nacc <= 0)) continue;   // handed back / nothing accepted: the columns stay as given
    if (b_off.empty()) b_off.push_back(0);
    if (!b_slot.empty() && (size_t)(b_off.back() + nacc) * rec_bytes > fat_budget)
      if (int e = flush()) return e;
    if (b_off.empty()) b_off.push_back(0);
    b_slot.push_back(i);
    b_cell.push_back(ids[i]);
    b_off.push_back(b_off.back() + nacc);
    g_stats.accepted_steps += nacc;
  }
  return flush();
}

static int sdirk_tlm_device(const SdirkParams& sp, const Cbm4Const& mc, int64_t ncells, int ntlm, double* d_y, double* d_ytlm,
                            const double* d_atol, const double* d_rtol, int* d_ist, double* d_rst, int* d_ierr, cudaStream_t st,
                            const SdirkAdjHost* adj = nullptr) {
  std::memset(&g_stats, 0, sizeof g_stats);
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  free_b += g_sd_tape.cap + g_sd_fat.cap;
  const size_t budget = g_tape_budget ? (size_t)g_tape_budget : (size_t)(0.6 * (double)free_b);
  const int cap = std::min(sp.max_no_steps + 1, g_sdirk_tape_slot > 0 ? g_sdirk_tape_slot : 4608);   // Nstp > Max_no_steps stops: one more may be accepted
  int64_t wave = g_wave_cells > 0 ? g_wave_cells : (int64_t)sms * 32 * 2;
  wave = std::min<int64_t>(wave, std::max<int64_t>(1, (int64_t)(0.35 * (double)budget / ((double)cap * SD_REC * sizeof(double)))));
  const size_t fat_budget = (size_t)(0.6 * (double)budget);
  auto run_all = [&](const std::vector<int>& all, bool dense) -> int {
    std::vector<int> ids, overflow;
    for (size_t w0 = 0; w0 < all.size(); w0 += (size_t)wave) {
      const size_t w1 = std::min(all.size(), w0 + (size_t)wave);
      ids.assign(all.begin() + (long)w0, all.begin() + (long)w1);
      if (int e = sdirk_tlm_group(sp, mc, ids, cap, ntlm, fat_budget, d_y, d_ytlm, d_atol, d_rtol, d_ist, d_rst, d_ierr, st, sms, overflow, adj, dense))
        return e;
      g_stats.waves++;
    }
    while (!overflow.empty()) {   // more accepted steps than the default slot: again with room for Max_no_steps
      const int capx = sp.max_no_steps + 1;
      const size_t per = (size_t)capx * SD_REC * sizeof(double);
      const size_t take = std::max<size_t>(1, std::min<size_t>(overflow.size(), (size_t)(0.35 * (double)budget) / per));
      std::vector<int> part(overflow.end() - (long)take, overflow.end()), again;
      overflow.resize(overflow.size() - take);
      if (int e = sdirk_tlm_group(sp, mc, part, capx, ntlm, fat_budget, d_y, d_ytlm, d_atol, d_rtol, d_ist, d_rst, d_ierr, st, sms, again, adj, dense))
        return e;
      if (!again.empty()) return fail(FATODE_ENOMEM, "SDIRK tape overflow with a Max_no_steps slot");
      g_stats.waves++;
    }
    return 0;
  };
  std::vector<int> all((size_t)ncells);
  for (int64_t c = 0; c < ncells; ++c) all[(size_t)c] = (int)c;
  if (int e = run_all(all, false)) return e;
  {   // second pass: systems that left the static pivot pattern kept their initial state; general partial pivoting throughout
    std::vector<int> h_ierr((size_t)ncells), irregular;
    CUDA_TRY(cudaMemcpyAsync(h_ierr.data(), d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t c = 0; c < ncells; ++c)
      if (h_ierr[(size_t)c] == SDIRK_IERR_IRREGULAR) irregular.push_back((int)c);
    if (!irregular.empty()) {
      if (int e = run_all(irregular, true)) return e;
      g_stats.cells_general = (int64_t)irregular.size();
    }
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int fatode_sdirk_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm, double tin,
                                                double tout, const double* atol_tlm, const double* rtol_tlm, const double* atol,
                                                const double* rtol, const int* icntrl_u, const double* rcntrl_u, int* istatus,
                                                double* rstatus, int* ierr, const double* model_params, int mem, void* stream) {
  g_err.clear();
  if (ncells < 0 || !y || !y_tlm || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (int e = check_model_dims(model, nvar, 0)) return e;
  if (ntlm < 1) return fail(FATODE_EINVAL, "ntlm < 1");
  if (model != FATODE_MODEL_CBM4) return fail(FATODE_EUNSUPPORTED, "the SDIRK tangent linear model is available for the cbm4 model only in this build");
  if (ntlm > 32) return fail(FATODE_EUNSUPPORTED, "ntlm > 32 per call is not supported yet");
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // INTEGRATE_TLM presets ICNTRL(5) = 8 before the `> 0` override (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:60-70)
  int ic[20];
  std::memset(ic, 0, sizeof ic);
  ic[4] = 8;
  if (icntrl_u) for (int i = 0; i < 20; ++i) if (icntrl_u[i] > 0) ic[i] = icntrl_u[i];
  SdirkParams sp;
  int tlm_opt[3] = {0, 0, 0};
  const int perr = sdirk_parse_options(nvar, ic, rcntrl_u, tin, tout, atol, rtol, sp, tlm_opt);
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  // the modes in which the columns feed back into the step sequence run in the lock-step kernel (sdirk_tlm_lock.cu)
  const bool direct = (tlm_opt[0] == 1), newton_est = (tlm_opt[1] != 0), trunc = (tlm_opt[2] != 0);
  const bool lock = direct || newton_est || trunc;
  if ((newton_est || trunc) && (!atol_tlm || !rtol_tlm))
    return fail(FATODE_EINVAL, "ICNTRL(9) / ICNTRL(12) need atol_tlm and rtol_tlm");
  Cbm4Const mc;
  cbm4_constants(model_params, mc);
  DevBuf b_atol, b_rtol, b_y, b_yt, b_ist, b_rst, b_ierr;
  CUDA_TRY(b_atol.alloc(sizeof(double) * nvar));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * nvar));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_yt = y_tlm, *d_rst = rstatus;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * nvar * ncells, nt = sizeof(double) * (size_t)nvar * ntlm * ncells;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_yt.alloc(nt));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b_yt.p, y_tlm, nt, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_yt = b_yt.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
  }
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  if (lock) {
    DevBuf b_at, b_rt;
    if (newton_est || trunc) {   // (N, NTLM) column-major, one set per call
      CUDA_TRY(b_at.alloc(sizeof(double) * nvar * ntlm));
      CUDA_TRY(b_rt.alloc(sizeof(double) * nvar * ntlm));
      CUDA_TRY(cudaMemcpyAsync(b_at.p, atol_tlm, sizeof(double) * nvar * ntlm, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(b_rt.p, rtol_tlm, sizeof(double) * nvar * ntlm, cudaMemcpyHostToDevice, st));
    }
    std::memset(&g_stats, 0, sizeof g_stats);
    UnitTiming ut;
    std::string uerr;
    if (int e = sdirk_tlm_lock_device(sp, mc, (long)ncells, nullptr, ntlm, direct, newton_est, trunc, d_y, d_yt, b_atol.as<double>(),
                                      b_rtol.as<double>(), b_at.as<double>(), b_rt.as<double>(), d_ist, d_rst, d_ierr, st, ut, uerr))
      return fail(e, uerr);
    g_stats.ms_forward += ut.ms;
    g_stats.launches += ut.launches;
    g_stats.launches_forward += 1;
    g_stats.waves = 1;
    g_stats.cells_general = ncells;
  } else if (int e = sdirk_tlm_device(sp, mc, ncells, ntlm, d_y, d_yt, b_atol.as<double>(), b_rtol.as<double>(), d_ist, d_rst, d_ierr, st)) {
    return e;
  }
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(y_tlm, d_yt, nt, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

// INTEGRATE_ADJ of ADJ/SDIRK_ADJ for CBM-IV (direct adjoint solve): sdirk_cell_adj.cuh
extern "C" int fatode_sdirk_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda,
                                                double* mu, double tin, double tout, const double* atol_adj, const double* rtol_adj,
                                                const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u,
                                                int* istatus, double* rstatus, int* ierr, double* mu_sum, const double* model_params,
                                                int mem, void* stream) {
  g_err.clear();
  if (ncells < 0 || !y || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (int e = check_model_dims(model, nvar, np)) return e;
  if (!lambda || nadj < 1) return fail(FATODE_EINVAL, "lambda is NULL or nadj < 1");
  if (model != FATODE_MODEL_CBM4) return fail(FATODE_EUNSUPPORTED, "the SDIRK discrete adjoint is available for the cbm4 model only in this build");
  if (nadj > 32) return fail(FATODE_EUNSUPPORTED, "nadj > 32 per call is not supported yet");
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // INTEGRATE_ADJ presets ICNTRL(5) = 8, ICNTRL(7) = 1, ICNTRL(8) = 0 before the `> 0` override (:70-83)
  int ic[20];
  std::memset(ic, 0, sizeof ic);
  ic[4] = 8; ic[6] = 1;
  if (icntrl_u) for (int i = 0; i < 20; ++i) if (icntrl_u[i] > 0) ic[i] = icntrl_u[i];
  const bool direct = (ic[6] == 1);
  int icp[20];
  std::memcpy(icp, ic, sizeof ic);
  icp[5] = 0;                                      // starting values are always interpolated in this family (:623)
  SdirkParams sp;
  const int perr = sdirk_parse_options(nvar, icp, rcntrl_u, tin, tout, atol, rtol, sp);
  if (ic[3] == 0) sp.max_no_steps = 10000;        // :375
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  if (!direct && (!atol_adj || !rtol_adj))
    return fail(FATODE_EINVAL, "the simplified-Newton adjoint stages (ICNTRL(7) /= 1) need atol_adj and rtol_adj");
  const bool with_mu = np > 0 && mu != nullptr;
  SdirkAdjHost ah;
  const int method = (ic[2] == 1 || ic[2] == 2 || ic[2] == 3 || ic[2] == 5) ? ic[2] : 4;
  std::memcpy(ah.co.A, SDIRK_ADJ_AB[method - 1].A, sizeof ah.co.A);
  std::memcpy(ah.co.B, SDIRK_ADJ_AB[method - 1].B, sizeof ah.co.B);
  ah.with_mu = with_mu ? 1 : 0;
  ah.np = np;
  Cbm4Const mc;
  cbm4_constants(model_params, mc);
  DevBuf b_atol, b_rtol, b_y, b_lam, b_mu, b_ist, b_rst, b_ierr, b_musum;
  CUDA_TRY(b_atol.alloc(sizeof(double) * nvar));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * nvar));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_lam = lambda, *d_mu = mu, *d_rst = rstatus, *d_musum = mu_sum;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * nvar * ncells, nl = sizeof(double) * (size_t)nvar * nadj * ncells,
               nm = sizeof(double) * (size_t)np * nadj * ncells;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
    CUDA_TRY(b_lam.alloc(nl));
    d_lam = b_lam.as<double>();
    CUDA_TRY(cudaMemcpyAsync(d_lam, lambda, nl, cudaMemcpyHostToDevice, st));
    if (with_mu) {
      CUDA_TRY(b_mu.alloc(nm));
      d_mu = b_mu.as<double>();
      CUDA_TRY(cudaMemcpyAsync(d_mu, mu, nm, cudaMemcpyHostToDevice, st));
      if (mu_sum) { CUDA_TRY(b_musum.alloc(sizeof(double) * np * nadj)); d_musum = b_musum.as<double>(); }
    }
  }
  ah.d_mu = d_mu;
  DevBuf b_ata, b_rta;
  if (!direct) {   // (N, NADJ) column-major, one set per call
    CUDA_TRY(b_ata.alloc(sizeof(double) * nvar * nadj));
    CUDA_TRY(b_rta.alloc(sizeof(double) * nvar * nadj));
    CUDA_TRY(cudaMemcpyAsync(b_ata.p, atol_adj, sizeof(double) * nvar * nadj, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b_rta.p, rtol_adj, sizeof(double) * nvar * nadj, cudaMemcpyHostToDevice, st));
    ah.newton = 1;
    ah.d_atol_adj = b_ata.as<double>();
    ah.d_rtol_adj = b_rta.as<double>();
  }
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_cbm4_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  if (int e = sdirk_tlm_device(sp, mc, ncells, nadj, d_y, d_lam, b_atol.as<double>(), b_rtol.as<double>(), d_ist, d_rst, d_ierr, st, &ah))
    return e;
  if (with_mu && d_musum)
    if (int e = mu_sum_device(ncells, nadj * np, d_mu, d_ierr, d_musum, st)) return e;
  CUDA_TRY(cudaStreamSynchronize(st));
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(lambda, d_lam, nl, cudaMemcpyDeviceToHost, st));
    if (with_mu) CUDA_TRY(cudaMemcpyAsync(mu, d_mu, nm, cudaMemcpyDeviceToHost, st));
    if (with_mu && mu_sum) CUDA_TRY(cudaMemcpyAsync(mu_sum, d_musum, sizeof(double) * np * nadj, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// FWD/ERK INTEGRATE and TLM/ERK_TLM INTEGRATE_TLM on the shallow-water ensemble: one member per CTA (erk_swe.cuh)
static thread_local CachedBuf g_erk_scratch;
static void release_erk_workspace() { g_erk_scratch.release(); }

static int erk_batch_impl(bool tlm, int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm, double tin, double tout,
                          const double* atol_tlm, const double* rtol_tlm, const double* atol, const double* rtol, const int* icntrl_u, const double* rcntrl_u, int* istatus,
                          double* rstatus, int* ierr, int mem, void* stream) {
  g_err.clear();
  if (ncells < 0 || !y || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (tlm && (!y_tlm || ntlm < 1)) return fail(FATODE_EINVAL, "y_tlm is NULL or ntlm < 1");
  if (int e = check_model_dims(model, nvar, 0, true)) return e;
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ErkParams ep;
  int tlm_opt[2] = {0, 0};
  const int perr = erk_parse_options(nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, ep, tlm ? tlm_opt : nullptr);
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  const bool dense = tlm && tlm_opt[0] != 0;   // ICNTRL(5) /= 0: JAC + MATMUL per stage, the example's JAC callback applied matrix-free
  const bool trunc = tlm && tlm_opt[1] != 0;
  if (trunc && (!atol_tlm || !rtol_tlm)) return fail(FATODE_EINVAL, "ERK_TLM with ICNTRL(12) /= 0 needs ATOL_TLM and RTOL_TLM");
  const int ntl = tlm ? ntlm : 0;
  const int ntol = ep.vector_tol ? nvar : 1;
  DevBuf b_atol, b_rtol, b_y, b_yt, b_ist, b_rst, b_ierr, b_atl, b_rtl;
  CUDA_TRY(b_atol.alloc(sizeof(double) * ntol));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * ntol));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * ntol, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * ntol, cudaMemcpyHostToDevice, st));
  if (trunc) {   // ATOL_TLM / RTOL_TLM (NVAR, NTLM): host arrays like ATOL / RTOL, shared by all members
    const size_t nb = sizeof(double) * (size_t)nvar * ntl;
    CUDA_TRY(b_atl.alloc(nb));
    CUDA_TRY(b_rtl.alloc(nb));
    CUDA_TRY(cudaMemcpyAsync(b_atl.p, atol_tlm, nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b_rtl.p, rtol_tlm, nb, cudaMemcpyHostToDevice, st));
  }
  double *d_y = y, *d_yt = y_tlm, *d_rst = rstatus;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * (size_t)nvar * ncells, nt = ny * (size_t)ntl;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    d_y = b_y.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
    if (tlm) {
      CUDA_TRY(b_yt.alloc(nt));
      CUDA_TRY(cudaMemcpyAsync(b_yt.p, y_tlm, nt, cudaMemcpyHostToDevice, st));
      d_yt = b_yt.as<double>();
    }
  }
  std::memset(&g_stats, 0, sizeof g_stats);
  int dev = 0, sms = 0, per_sm = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  auto kern = tlm ? k_erk_swe<true> : k_erk_swe<false>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ERK_SMEM));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ERK_THREADS, ERK_SMEM));
  const unsigned blocks = (unsigned)std::min<int64_t>(ncells, (int64_t)sms * std::max(per_sm, 1));
  CUDA_TRY(g_erk_scratch.ensure(sizeof(double) * (size_t)blocks * ((size_t)ep.S * SWE_N * (1 + ntl) + (dense ? (size_t)SWE_NNZR * SWE_N : 0))));
  CUDA_TRY(g_ctr.ensure(16 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(g_ctr.p, 0, 16 * sizeof(unsigned long long), st));
  ErkArgs ea{(long)ncells, d_y, d_y, d_yt, b_atol.as<double>(), b_rtol.as<double>(), d_ist, d_rst, d_ierr,
             g_ctr.as<unsigned long long>(), g_erk_scratch.as<double>(), ntl, trunc ? 1 : 0, b_atl.as<double>(), b_rtl.as<double>(), dense ? 1 : 0};
  PhaseTimer tm(st);
  tm.start();
  kern<<<blocks, ERK_THREADS, ERK_SMEM, st>>>(ep, ea);
  CUDA_TRY(cudaGetLastError());
  g_stats.ms_forward += tm.stop();
  g_stats.launches++;
  g_stats.launches_forward++;
  g_stats.waves++;
  CUDA_TRY(cudaStreamSynchronize(st));
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    if (tlm) CUDA_TRY(cudaMemcpyAsync(y_tlm, d_yt, nt, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int fatode_erk_integrate_batch(int model, int64_t ncells, int nvar, double* y, double tin, double tout, const double* atol,
                                          const double* rtol, const int* icntrl_u, const double* rcntrl_u, int* istatus,
                                          double* rstatus, int* ierr, const double* model_params, int mem, void* stream) {
  (void)model_params;
  return erk_batch_impl(false, model, ncells, nvar, 0, y, nullptr, tin, tout, nullptr, nullptr, atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus, ierr,
                        mem, stream);
}

extern "C" int fatode_erk_integrate_tlm_batch(int model, int64_t ncells, int nvar, int ntlm, double* y, double* y_tlm, double tin,
                                              double tout, const double* atol_tlm, const double* rtol_tlm, const double* atol,
                                              const double* rtol, const int* icntrl_u, const double* rcntrl_u, int* istatus,
                                              double* rstatus, int* ierr, const double* model_params, int mem, void* stream) {
  (void)model_params;   // the TLM tolerances (host arrays [ntlm][nvar], shared by all members) are only read with ICNTRL(12) /= 0
  return erk_batch_impl(true, model, ncells, nvar, ntlm, y, y_tlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, icntrl_u, rcntrl_u, istatus, rstatus, ierr,
                        mem, stream);
}

// INTEGRATE_ADJ of ADJ/ERK_ADJ on the shallow-water ensemble (erk_adj_swe.cu)
static int g_erk_tape_slot = 0;
extern "C" int fatode_set_erk_tape_slot(int checkpoints) { g_erk_tape_slot = checkpoints > 0 ? checkpoints : 0; return 0; }
extern "C" int fatode_erk_integrate_adj_batch(int model, int64_t ncells, int nvar, int np, int nadj, double* y, double* lambda, double* mu,
                                              double tin, double tout, const double* atol, const double* rtol, const int* icntrl_u,
                                              const double* rcntrl_u, int* istatus, double* rstatus, int* ierr,
                                              const double* model_params, int mem, void* stream) {
  (void)model_params;
  g_err.clear();
  if (ncells < 0 || !y || !lambda || !atol || !rtol || !istatus || !rstatus || !ierr) return fail(FATODE_EINVAL, "NULL required argument");
  if (nadj < 1) return fail(FATODE_EINVAL, "nadj < 1");
  if (int e = check_model_dims(model, nvar, 0, true)) return e;
  if (np > 0 && mu) return fail(FATODE_EUNSUPPORTED, "ERK_ADJ with parameter sensitivities (NP > 0, Mu, JACP): the shallow-water example has no parameters");
  if (nadj > nvar) return fail(FATODE_EINVAL, "nadj > nvar (ADJINIT sets Lambda(k,k) = 1)");
  if (int e = check_device()) return e;
  if (ncells == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ErkParams ep;
  const int perr = erk_parse_options(nvar, icntrl_u, rcntrl_u, tin, tout, atol, rtol, ep, nullptr);
  if (perr < 0) return broadcast_ierr(ncells, perr, istatus, rstatus, ierr, mem, st);
  const int ntol = ep.vector_tol ? nvar : 1;
  DevBuf b_atol, b_rtol, b_y, b_lam, b_ist, b_rst, b_ierr;
  CUDA_TRY(b_atol.alloc(sizeof(double) * ntol));
  CUDA_TRY(b_rtol.alloc(sizeof(double) * ntol));
  CUDA_TRY(cudaMemcpyAsync(b_atol.p, atol, sizeof(double) * ntol, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b_rtol.p, rtol, sizeof(double) * ntol, cudaMemcpyHostToDevice, st));
  double *d_y = y, *d_lam = lambda, *d_rst = rstatus;
  int *d_ist = istatus, *d_ierr = ierr;
  const size_t ny = sizeof(double) * (size_t)nvar * ncells, nl = ny * (size_t)nadj;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(b_y.alloc(ny));
    CUDA_TRY(b_lam.alloc(nl));
    CUDA_TRY(b_ist.alloc(sizeof(int) * 20 * ncells));
    CUDA_TRY(b_rst.alloc(sizeof(double) * 20 * ncells));
    CUDA_TRY(b_ierr.alloc(sizeof(int) * ncells));
    CUDA_TRY(cudaMemcpyAsync(b_y.p, y, ny, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b_lam.p, lambda, nl, cudaMemcpyHostToDevice, st));   // members that stop in ERK_Push keep the caller's Lambda
    d_y = b_y.as<double>(); d_lam = b_lam.as<double>(); d_ist = b_ist.as<int>(); d_rst = b_rst.as<double>(); d_ierr = b_ierr.as<int>();
  }
  std::memset(&g_stats, 0, sizeof g_stats);
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const size_t budget = g_tape_budget ? (size_t)g_tape_budget : (size_t)(0.6 * (double)free_b);
  UnitTiming ut;
  std::string uerr;
  long retried = 0;
  if (int e = erk_adj_swe_device(ep, (long)ncells, nadj, d_y, d_lam, b_atol.as<double>(), b_rtol.as<double>(), d_ist, d_rst, d_ierr,
                                 g_erk_tape_slot, budget, st, ut, retried, uerr))
    return fail(e, uerr);
  g_stats.ms_forward += ut.ms;
  g_stats.launches += ut.launches;
  g_stats.launches_forward += ut.launches;
  g_stats.waves = (int)ut.launches;
  g_stats.cells_deferred = retried;
  if (mem == FATODE_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(y, d_y, ny, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(lambda, d_lam, nl, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(istatus, d_ist, sizeof(int) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(rstatus, d_rst, sizeof(double) * 20 * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ierr, d_ierr, sizeof(int) * ncells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Unit-level probes (used by tests/): evaluate the device model / warp LU on explicit inputs.
template <class Model>
__global__ void k_probe_model(typename Model::Const mc, double t, const double* y, double* f, double* jac_dense,
                              double* hess) {
  extern __shared__ __align__(16) double smem[];
  typedef FwdSmem<Model> L;
  const int lane = threadIdx.x;
  double* ws = smem + L::OFF_MODEL;
  double* jv = smem + L::OFF_JV;
  double* tile = smem + L::OFF_A;
  Model::init(ws, mc, lane);
  Model::set_time(ws, t, lane);
  Model::set_y(ws, y[lane], lane);
  f[lane] = Model::fun(ws, mc, lane);
  Model::jac(ws, jv, lane);
  Model::template scatter_E<LDA>(jv, 0.0, tile, lane);     // tile = -J (column-major)
  for (int j = 0; j < 32; ++j) jac_dense[j * 32 + lane] = -tile[j * LDA + lane];
  __syncwarp();
  double* hs = tile;   // the tile is free again
  Model::hess(ws, hs, lane);
  for (int h = lane; h < Model::NHESS; h += 32) hess[h] = hs[h];
}

__global__ void k_probe_lu(const double* A /*col-major 32x32*/, const double* b, double* lu_out /*row-major logical*/,
                           int* ipiv_who, double* x, int* info) {
  __shared__ double tile[32 * LDA];
  __shared__ int who[32];
  const int lane = threadIdx.x;
  for (int j = 0; j < 32; ++j) tile[j * LDA + lane] = A[j * 32 + lane];
  __syncwarp();
  int pos;
  LUMasks mk;
  int inf = warp_lu_factor(tile, who, pos, lane, &mk);
  for (int j = 0; j < 32; ++j) lu_out[pos * 32 + j] = tile[j * LDA + lane];
  ipiv_who[lane] = who[lane];
  if (lane == 0) *info = inf;
  x[lane] = warp_lu_solve(tile, who, pos, mk, b[lane], lane);
}

extern "C" int fatode_probe_cbm4(double t, const double* y, double* f, double* jac_dense, double* hess) {
  if (int e = check_device()) return e;
  Cbm4Const mc;
  cbm4_constants(nullptr, mc);
  DevBuf by, bf, bj, bh;
  CUDA_TRY(by.alloc(256)); CUDA_TRY(bf.alloc(256)); CUDA_TRY(bj.alloc(8192)); CUDA_TRY(bh.alloc(8 * 264));
  CUDA_TRY(cudaMemcpy(by.p, y, 256, cudaMemcpyHostToDevice));
  const size_t sm = sizeof(double) * fwd_ws_doubles<Cbm4Warp>();
  k_probe_model<Cbm4Warp><<<1, 32, sm>>>(mc, t, by.as<double>(), bf.as<double>(), bj.as<double>(), bh.as<double>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(f, bf.p, 256, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(jac_dense, bj.p, 8192, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(hess, bh.p, 8 * 258, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int fatode_probe_lu32(const double* A, const double* b, double* lu_out, int* who, double* x, int* info) {
  if (int e = check_device()) return e;
  DevBuf bA, bb, bl, bw, bx, bi;
  CUDA_TRY(bA.alloc(8192)); CUDA_TRY(bb.alloc(256)); CUDA_TRY(bl.alloc(8192)); CUDA_TRY(bw.alloc(128)); CUDA_TRY(bx.alloc(256));
  CUDA_TRY(bi.alloc(4));
  CUDA_TRY(cudaMemcpy(bA.p, A, 8192, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(bb.p, b, 256, cudaMemcpyHostToDevice));
  k_probe_lu<<<1, 32>>>(bA.as<double>(), bb.as<double>(), bl.as<double>(), bw.as<int>(), bx.as<double>(), bi.as<int>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(lu_out, bl.p, 8192, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(who, bw.p, 128, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(x, bx.p, 256, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(info, bi.p, 4, cudaMemcpyDeviceToHost));
  return 0;
}
