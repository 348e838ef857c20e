// TMEM as a per-lane FP64 scratchpad: correctness + rough cost of tcgen05.st / tcgen05.ld (32x32b.x2).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st2(uint32_t taddr, double v) {
  uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tmem_ld2(uint32_t taddr) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(128, 1) k_tmem(int iters, int* err, long long* cyc_out, double* sink) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = base_s;
  const uint32_t wbase = base + ((uint32_t)(warp * 32) << 16);
  // write 256 doubles per lane, read back
  for (int c = 0; c < 256; ++c) tmem_st2(wbase + 2 * c, (double)(blockIdx.x * 1000003 + threadIdx.x * 257 + c) * 1.25);
  tmem_wait_st();
  int bad = 0;
  for (int c = 255; c >= 0; --c) {
    double v = tmem_ld2(wbase + 2 * c);
    if (v != (double)(blockIdx.x * 1000003 + threadIdx.x * 257 + c) * 1.25) bad++;
  }
  if (bad) atomicAdd(err, bad);
  // timing: read-modify-write sweeps over 160 doubles (the W arrays)
  __syncthreads();
  long long t0 = clock64();
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < 160; ++c) {
      double v = tmem_ld2(wbase + 2 * c);
      v = fma(v, 1.0000001, 1e-9);
      tmem_st2(wbase + 2 * c, v);
      acc += v;
    }
    tmem_wait_st();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc_out = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "n"(512));
}

// same RMW pattern on shared memory for comparison
__global__ void __launch_bounds__(128, 1) k_smem(int iters, long long* cyc_out, double* sink) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* w = sm + warp * 160 * 32;
  for (int c = 0; c < 160; ++c) w[c * 32 + lane] = c;
  __syncthreads();
  long long t0 = clock64();
  double acc = 0.0;
  for (int it = 0; it < iters; ++it)
    for (int c = 0; c < 160; ++c) {
      double v = w[c * 32 + lane];
      v = fma(v, 1.0000001, 1e-9);
      w[c * 32 + lane] = v;
      acc += v;
    }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc_out = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int* err; long long* cyc; double* sink;
  cudaMalloc(&err, 4); cudaMalloc(&cyc, 8); cudaMalloc(&sink, 8 * 148 * 128);
  cudaMemset(err, 0, 4);
  const int iters = 200;
  k_tmem<<<148, 128>>>(iters, err, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = -1; long long hc = 0;
  cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("tmem: cuda=%s mismatches=%d cycles/RMW(per warp, 4 warps/SM)=%.1f\n", cudaGetErrorString(e), herr, (double)hc / (iters * 160.0));
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 160 * 32 * 8);
  k_smem<<<148, 128, 4 * 160 * 32 * 8>>>(iters, cyc, sink);
  e = cudaDeviceSynchronize();
  cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("smem: cuda=%s cycles/RMW(per warp, 4 warps/SM)=%.1f\n", cudaGetErrorString(e), (double)hc / (iters * 160.0));
  return 0;
}
